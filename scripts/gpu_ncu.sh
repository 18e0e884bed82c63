mkdir -p gpurun_out
cap() { name=$1; target=$2; regex=$3; skip=$4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/ncu_$name python bench/ncu_targets.py $target > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"; }
cap fused_cw fused_cw fused_ps_cw_kernel 2
cap gram_tma gram_tma gram_umma_tma_kernel 2
cap wsum_multi wsum_multi wsum_multi_kernel 2
cap bn_cluster_fwd bn_cluster "bn_cluster_kernel.*Lb0" 2
cap bn_cluster_bwd bn_cluster "bn_cluster_kernel.*Lb1" 2
cap nnm_map preagg preagg_map_kernel 2
cap caf preagg caf_kernel 2
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
