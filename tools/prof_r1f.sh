# Round-1 closing pass: bench (N=1), launch lists of one HalfResNet34 / TDNN step, ncu --set full summaries of the kernels
# that changed after the r01_final captures (front-ends, stem, TDNN pack kernel).
set -x
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 3 > $O/bench_r01f.json 2> $O/bench_r01f.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r01f_reference.json 2> $O/bench_r01f_reference.err
NCU="ncu --set full --clock-control none --import-source on -f"
summ() { for i in $(seq 0 $(($2 - 1))); do python tools/ncu_summary.py $O/$1.ncu-rep $i > $O/sumf_$3_$i.txt 2>&1; done; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_hr34_r01f.csv python tools/profile_step.py hr34 2 > $O/ncu_hr34.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_tdnn_r01f.csv python tools/profile_step.py tdnn 2 > $O/ncu_tdnn.log 2>&1
$NCU -k regex:'frontend_kernel|stem_kernel' -c 2 -o $O/rf_fe python tools/profile_step.py hr34 1 > $O/ncu_fe.log 2>&1
summ rf_fe 2 fe; rm -f $O/rf_fe.ncu-rep
$NCU -k regex:'frontend_kernel|pack_frames' -c 2 -o $O/rf_fet python tools/profile_step.py tdnn 1 > $O/ncu_fet.log 2>&1
summ rf_fet 2 fet; rm -f $O/rf_fet.ncu-rep
rm -f $O/*.ncu-rep
du -sh $O
