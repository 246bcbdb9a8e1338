# GroupNorm-in-the-epilogue: kernel parity, whole-network parity, A/B inside the step graph
set -u
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gn_epilogue" > $O/pytest_epi.log 2>&1; echo "epi rc=$?" | tee $O/summary.txt
tail -25 $O/pytest_epi.log | cut -c1-300 | tee -a $O/summary.txt
DMU_DRIFT_OUT=$O/drift timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -12 $O/pytest.log | cut -c1-300 | tee -a $O/summary.txt
. scripts/ab.sh
run DMU_GN_EPI=0 | tee -a $O/summary.txt
run DMU_GN_EPI=1 | tee -a $O/summary.txt
DMU_GN_EPI=0 timeout 300 python scripts/phase_times.py 2>&1 | tail -7 > $O/phase_off.txt; DMU_GN_EPI=1 timeout 300 python scripts/phase_times.py 2>&1 | tail -7 > $O/phase_on.txt
paste $O/phase_off.txt $O/phase_on.txt | cut -c1-200 | tee -a $O/summary.txt
rund DMU_GN_EPI=0 | tee -a $O/summary.txt
rund DMU_GN_EPI=1 | tee -a $O/summary.txt
