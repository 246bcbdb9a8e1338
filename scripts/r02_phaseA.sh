# Round 2, phase A: the new parity tests (configs[2] shape, full chains, energy TrainStep) and one line per bench workload.
set -u
O=gpurun_out/r02b; mkdir -p $O
DMU_DRIFT_OUT=$O/drift timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee $O/summary.txt
tail -15 $O/pytest.log | tee -a $O/summary.txt
grep -h "drift\|cfg3" $O/pytest.log | tee -a $O/summary.txt
DMU_DRIFT_OUT=$O/drift timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q -s -k "cfg3 or full_chain" > $O/pytest_chain.log 2>&1
grep -h "drift\|cfg3" $O/pytest_chain.log | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench_train.json 2> $O/bench_train.err; echo "train rc=$?" | tee -a $O/summary.txt; cut -c1-300 $O/bench_train.json | tee -a $O/summary.txt
for w in ddim ddpm_sample score energy; do
  timeout 600 python bench.py --workload $w --cpu-budget 8 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?" | tee -a $O/summary.txt
  cut -c1-300 $O/bench_$w.json | tee -a $O/summary.txt; tail -3 $O/bench_$w.err
done
