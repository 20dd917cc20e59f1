# per-kernel durations (ncu, serialised) of one HalfResNet34 step for two settings of an A/B knob
for v in 1 0; do
  SKB_FUSED_SUMS_64=$v ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sums64_$v.csv python tools/profile_step.py hr34 2 > /dev/null 2>&1
  echo "SKB_FUSED_SUMS_64=$v"; python tools/parse_launches.py gpurun_out/launches_sums64_$v.csv | head -12
done
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
SKB_FUSED_SUMS_64=0 $B > gpurun_out/ab_s64_off1.json 2>/dev/null
$B > gpurun_out/ab_s64_on1.json 2>/dev/null
SKB_FUSED_SUMS_64=0 $B > gpurun_out/ab_s64_off2.json 2>/dev/null
$B > gpurun_out/ab_s64_on2.json 2>/dev/null
python tools/ab_print.py gpurun_out/ab_s64_*.json
