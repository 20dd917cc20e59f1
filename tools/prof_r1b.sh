# Round-1 (second half) profiling pass on the GPU box: tests, bench, write-bandwidth ceiling, ncu captures.
# Reports are summarised to text on the box (gpurun copies back at most 64 MiB) and only two .ncu-rep files are kept.
set -x
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench5.json 2> $O/bench5.err || exit 1
./tools/wr_bench > $O/wr_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
summ() {  # rep n_launches tag
  for i in $(seq 0 $(($2 - 1))); do python tools/ncu_summary.py $O/$1.ncu-rep $i > $O/sum_$3_$i.txt 2>&1; done
}
$NCU -k regex:score_gemm -c 1 -o $O/r1b_plda python tools/profile_step.py plda 1 > $O/ncu_plda.log 2>&1
summ r1b_plda 1 plda
for skip in 3 10 19 32; do
  $NCU -k regex:conv_umma --launch-skip $skip --launch-count 2 -o $O/r1b_conv_$skip python tools/profile_step.py hr34 1 > $O/ncu_conv.log 2>&1
  summ r1b_conv_$skip 2 conv$skip
  if [ $skip != 3 ]; then rm -f $O/r1b_conv_$skip.ncu-rep; fi
done
$NCU -k regex:'plane_sum|se_mean_partial|se_border|se_fc' -c 4 -o $O/r1b_se python tools/profile_step.py hr34 1 > $O/ncu_se.log 2>&1
summ r1b_se 4 se; rm -f $O/r1b_se.ncu-rep
$NCU -k regex:'frontend_kernel|cmvn|stem_kernel' -c 4 -o $O/r1b_fe python tools/profile_step.py hr34 1 > $O/ncu_fe.log 2>&1
summ r1b_fe 4 fe; rm -f $O/r1b_fe.ncu-rep
$NCU -k regex:'softmax_pool|meanstd|gather_frames|pack_split' -c 6 -o $O/r1b_pool python tools/profile_step.py hr34 1 > $O/ncu_pool.log 2>&1
summ r1b_pool 6 pool; rm -f $O/r1b_pool.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_hr34_v17.csv python tools/profile_step.py hr34 2 > $O/ncu_hr34.log 2>&1
du -sh $O
