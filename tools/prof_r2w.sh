# Round 2 closing evidence on ONE B200 (HEAD): GPU tests, smoke, both bench arms, launch lists of one HalfResNet34 / TDNN step and
# of every score-GEMM mode (duration-only ncu passes; the step's full conv table is profiles/r02s_ncu_conv_step_table.txt -- the
# convolution kernel has not changed since).
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02w_pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/r02w_pytest_gpu.txt
tail -2 $O/r02w_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a $O/r02w_pytest_gpu.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02w_bench_reference_arm.json 2> $O/r02w_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 3 > $O/r02w_bench_n1.json 2> $O/r02w_bench_n1.err; echo "bench rc=$?"
python tools/ab_print.py $O/r02w_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02w_launches_hr34.csv python tools/step_for_ncu.py hr34 1 > /dev/null 2>&1
python tools/step_summary.py $O/r02w_launches_hr34.csv > $O/r02w_launches_hr34_step_summary.txt 2>&1; head -4 $O/r02w_launches_hr34_step_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02w_launches_tdnn.csv python tools/step_for_ncu.py tdnn 1 > /dev/null 2>&1
python tools/step_summary.py $O/r02w_launches_tdnn.csv > $O/r02w_launches_tdnn_step_summary.txt 2>&1; head -4 $O/r02w_launches_tdnn_step_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02w_score_modes_launches.csv python tools/score_modes_probe.py > /dev/null 2>&1
grep -c skb $O/r02w_score_modes_launches.csv
