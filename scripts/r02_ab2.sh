set -u
O=gpurun_out/r02g; mkdir -p $O
. scripts/ab.sh
for cfg in "A=1" "DMU_GN_BWD_SMEM_MIN_KB=32" "DMU_GN_BWD_SMEM_MIN_KB=16" "DMU_GN_BWD_SMEM_MIN_KB=8" "DMU_GN_BWD_SMEM_MIN_KB=16 DMU_GN_BWD_SMEM_KB=140" "DMU_GN_BWD_SMEM_KB=140" "DMU_GN_FWD_FUSED_MB=2" "DMU_GN_FWD_FUSED_MB=0" "A=2"; do
  run $cfg | tee -a $O/ab.txt
done
