# Round-1 verification pass of the final conv kernel (contiguous tile ranges, fused channel totals): GPU tests, smoke,
# bench (N=1 + reference arm), launch list of one HalfResNet34 step, ncu --set full of the layer-1 conv1 launch.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r01h.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_r01h.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_r01h.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/bench_r01h.json 2> $O/bench_r01h.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r01h_reference.json 2> $O/bench_r01h_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_hr34_r01h.csv python tools/profile_step.py hr34 2 > $O/ncu_hr34.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:"stem_kernel|resample_kernel" -c 1 -o $O/rh_conv python tools/profile_step.py hr34 1 > $O/ncu_conv.log 2>&1
python tools/ncu_summary.py $O/rh_conv.ncu-rep 0 > $O/sumh_stem.txt 2>&1
rm -f $O/*.ncu-rep
tail -3 $O/pytest_gpu_r01h.log
cat $O/bench_r01h.json | head -c 1500
