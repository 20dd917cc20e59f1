# A/B of the late PDL trigger in the conv kernel (SKB_PDL_LATE=1) against the shipped implicit trigger
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
SKB_PDL_LATE=1 timeout 300 python -m pytest tests/test_extraction_gpu.py -m gpu -x -q 2>&1 | tail -1
SKB_PDL_LATE=1 $B > gpurun_out/ab_late_on1.json 2>/dev/null
$B > gpurun_out/ab_late_off1.json 2>/dev/null
SKB_PDL_LATE=1 $B > gpurun_out/ab_late_on2.json 2>/dev/null
$B > gpurun_out/ab_late_off2.json 2>/dev/null
python tools/ab_print.py gpurun_out/ab_late_*.json
