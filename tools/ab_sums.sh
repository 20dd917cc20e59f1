timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/ab_sums_new1.json 2>/dev/null
SKB_NO_FUSED_SUMS=1 $B > gpurun_out/ab_sums_off1.json 2>/dev/null
$B > gpurun_out/ab_sums_new2.json 2>/dev/null
SKB_NO_WARP_BORDER=1 $B > gpurun_out/ab_sums_cta_border.json 2>/dev/null
python tools/ab_print.py gpurun_out/ab_sums_*.json
python - <<PY
import json
for f in ("new1","off1","new2","cta_border"):
    d=json.load(open("gpurun_out/ab_sums_%s.json"%f)); print(f, {k: round(v/20,3) for k,v in d["roofline"]["device_ms_by_category"].items()}, round(d["roofline"]["frac"],4))
PY
