# ncu --set full of the first conv launch of a HalfResNet34 step (layer1.0 conv1 with the fused channel totals) and,
# for comparison, the same launch with SKB_NO_FUSED_SUMS=1
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:'conv_umma_kernel' -c 1 -o $O/c1_sums python tools/profile_step.py hr34 1 > $O/ncu_c1.log 2>&1
NCU_CTX=2 python tools/ncu_summary.py $O/c1_sums.ncu-rep 0 > $O/sum_conv1_fused_sums.txt 2>&1
SKB_NO_FUSED_SUMS=1 $NCU -k regex:'conv_umma_kernel' -c 1 -o $O/c1_nosums python tools/profile_step.py hr34 1 > $O/ncu_c1b.log 2>&1
NCU_CTX=2 python tools/ncu_summary.py $O/c1_nosums.ncu-rep 0 > $O/sum_conv1_no_sums.txt 2>&1
rm -f $O/*.ncu-rep
head -24 $O/sum_conv1_fused_sums.txt | cut -c1-150
