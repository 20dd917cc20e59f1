# Final round-1 profiling pass: bench (N=1), launch lists and ncu --set full summaries of the top kernels.
set -x
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err || exit 1
NCU="ncu --set full --clock-control none --import-source on -f"
summ() { for i in $(seq 0 $(($2 - 1))); do python tools/ncu_summary.py $O/$1.ncu-rep $i > $O/sumf_$3_$i.txt 2>&1; done; }
# the same command for the launch list and the full captures (one pass of the HalfResNet34 step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_hr34_final.csv python tools/profile_step.py hr34 2 > $O/ncu_hr34.log 2>&1
for skip in 3 10 19 32; do
  $NCU -k regex:conv_umma --launch-skip $skip --launch-count 2 -o $O/rf_conv_$skip python tools/profile_step.py hr34 1 > $O/ncu_conv.log 2>&1
  summ rf_conv_$skip 2 conv$skip
  if [ $skip != 19 ]; then rm -f $O/rf_conv_$skip.ncu-rep; fi
done
$NCU -k regex:'plane_sum|se_border|se_mean_partial|se_fc' -c 4 -o $O/rf_se python tools/profile_step.py hr34 1 > $O/ncu_se.log 2>&1
summ rf_se 4 se; rm -f $O/rf_se.ncu-rep
$NCU -k regex:'frontend_kernel|stem_kernel' -c 2 -o $O/rf_fe python tools/profile_step.py hr34 1 > $O/ncu_fe.log 2>&1
summ rf_fe 2 fe; rm -f $O/rf_fe.ncu-rep
$NCU -k regex:score_gemm -c 1 -o $O/rf_plda python tools/profile_step.py plda 1 > $O/ncu_plda.log 2>&1
summ rf_plda 1 plda; rm -f $O/rf_plda.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_tdnn_final.csv python tools/profile_step.py tdnn 2 > $O/ncu_tdnn.log 2>&1
du -sh $O
