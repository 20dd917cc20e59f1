timeout 600 python -m pytest tests/test_extraction_gpu.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pool2.csv python tools/profile_step.py hr34 2 > /dev/null 2>&1
python tools/parse_launches.py gpurun_out/launches_pool2.csv | grep -E "launches in pass|softmax_pool|meanstd|gather_frames"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pool2.json 2>/dev/null; python tools/ab_print.py gpurun_out/bench_pool2.json
