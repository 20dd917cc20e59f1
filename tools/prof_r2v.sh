# Round 2 final evidence for the shipped bench configuration (192 utterances / 2087 audio-s per step):
# A/B of the step size, launch list of one step (duration only), ncu --set full of the 36 conv launches of one step.
O=gpurun_out
python bench.py --utts 96 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02v_bench_utts96.json 2> $O/r02v_bench_utts96.err; echo "bench96 rc=$?"
python bench.py --steps 20 --warmup 3 > $O/r02v_bench_n1.json 2> $O/r02v_bench_n1.err; echo "bench192 rc=$?"
python tools/ab_print.py $O/r02v_bench_utts96.json $O/r02v_bench_n1.json
python tools/step_for_ncu.py hr34 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02v_launches_hr34.csv python tools/step_for_ncu.py hr34 1 > $O/r02v_ncu_launches.log 2>&1
python tools/step_summary.py $O/r02v_launches_hr34.csv > $O/r02v_launches_hr34_step_summary.txt 2>&1; head -12 $O/r02v_launches_hr34_step_summary.txt
ncu --set full --clock-control none --import-source on -k regex:conv_umma --launch-skip 72 --launch-count 36 -o /tmp/r02v_conv -f python tools/step_for_ncu.py hr34 1 > $O/r02v_ncu_full.log 2>&1
python tools/ncu_table.py /tmp/r02v_conv.ncu-rep > $O/r02v_ncu_conv_step_table.txt 2>&1; head -5 $O/r02v_ncu_conv_step_table.txt
