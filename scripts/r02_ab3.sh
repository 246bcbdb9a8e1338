set -u
O=gpurun_out/r02j; mkdir -p $O
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "adam" 2>&1 | tail -2 | tee $O/ab.txt
. scripts/ab.sh
for cfg in "DMU_ADAM_OVERLAP=0" "DMU_ADAM_OVERLAP=1" "DMU_ADAM_OVERLAP=1 DMU_ADAM_CTAS_PER_SM=4" "DMU_ADAM_OVERLAP=1 DMU_ADAM_CTAS_PER_SM=2" "DMU_ADAM_OVERLAP=1 DMU_ADAM_CTAS_PER_SM=1" "DMU_ADAM_OVERLAP=0 DMU_ADAM_CTAS_PER_SM=4"; do
  run $cfg | tee -a $O/ab.txt
done
timeout 300 python scripts/phase_times.py 2>&1 | tail -7 | tee -a $O/ab.txt
