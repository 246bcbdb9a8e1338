set -u
O=gpurun_out/r02f; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "groupnorm or gn_" 2>&1 | tail -3 | tee $O/tests.txt
for cfg in "A=1" "DMU_GN_BWD_SMEM_MIN_KB=16" "DMU_GN_BWD_SMEM_KB=36"; do
  echo "== $cfg" | tee -a $O/sweep2.txt
  env $cfg python scripts/gn_time.py 2>&1 | cut -c1-200 | head -9 | tee -a $O/sweep2.txt
done
bash scripts/r02_check.sh r02f
