# Closing launch lists of the round (final kernels): one HalfResNet34 step and one TDNN step, ncu --metrics gpu__time_duration.sum
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_hr34_r01i.csv python tools/profile_step.py hr34 2 > $O/ncu_hr34.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_tdnn_r01i.csv python tools/profile_step.py tdnn 2 > $O/ncu_tdnn.log 2>&1
python tools/parse_launches.py $O/launches_hr34_r01i.csv > $O/launches_hr34_r01i_summary.txt
python tools/parse_launches.py $O/launches_tdnn_r01i.csv > $O/launches_tdnn_r01i_summary.txt
head -3 $O/launches_hr34_r01i_summary.txt; head -8 $O/launches_tdnn_r01i_summary.txt
