# full GPU test suite + phase times + train / ddim lines (quick)
set -u
O=gpurun_out/${1:-r02e}; mkdir -p $O
DMU_DRIFT_OUT=$O/drift timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt
tail -6 $O/pytest.log | cut -c1-300 | tee -a $O/summary.txt
grep -h "snr time weights" $O/pytest.log | tee -a $O/summary.txt
timeout 300 python scripts/phase_times.py 2>&1 | tail -7 | tee -a $O/summary.txt
. scripts/ab.sh
run A=1 | tee -a $O/summary.txt
rund A=1 | tee -a $O/summary.txt
