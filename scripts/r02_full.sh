set -u
O=${1:-gpurun_out/r02v}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 10 > $O/bench_train.json 2> $O/bench_train.err; echo "bench rc=$?"; cut -c1-400 $O/bench_train.json
for w in ddim ddpm_sample score energy; do
timeout 900 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"; cut -c1-300 $O/bench_$w.json
done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
