# Final state of round 2: GPU test suite, the five bench workloads, smoke, run-to-run spread
set -u
O=gpurun_out/r02r; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -2 $O/pytest_gpu.log
timeout 400 python bench.py --steps 50 --warmup 10 > $O/bench_train.json 2> $O/bench_train.err; echo "bench rc=$?"; cut -c1-300 $O/bench_train.json
for w in ddim ddpm_sample score energy; do
timeout 300 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 $O/bench_$w.json
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 120 python scripts/determinism_check.py > $O/determinism.txt 2>&1; cat $O/determinism.txt
