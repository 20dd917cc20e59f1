# ncu --set full summaries of the kernels added / rewritten in the last sessions: resample, softmax_pool, meanstd
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:'resample_kernel' -c 1 -o $O/ri_rs python tools/resample_probe.py > $O/ncu_rs.log 2>&1
python tools/ncu_summary.py $O/ri_rs.ncu-rep 0 > $O/sumi_resample.txt 2>&1
$NCU -k regex:'softmax_pool_kernel|meanstd_kernel' -c 2 -o $O/ri_pool python tools/profile_step.py hr34 1 > $O/ncu_pool.log 2>&1
python tools/ncu_summary.py $O/ri_pool.ncu-rep 0 > $O/sumi_pool_0.txt 2>&1
python tools/ncu_summary.py $O/ri_pool.ncu-rep 1 > $O/sumi_pool_1.txt 2>&1
rm -f $O/*.ncu-rep
head -20 $O/sumi_resample.txt | cut -c1-150
