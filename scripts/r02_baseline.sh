# Round-2 baseline on a fresh box: tests, the two switches left unmeasured by round 1, phase times, bench line, launch list.
set -u
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
timeout 300 python scripts/phase_times.py > $O/phase_times.txt 2>&1; cat $O/phase_times.txt | tee -a $O/summary.txt
timeout 600 bash scripts/next_gpu_checks.sh > $O/next_checks.txt 2>&1; tail -12 $O/next_checks.txt | tee -a $O/summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ddim > $O/ncu.log 2>&1
python scripts/ncu_launch_summary.py $O/launches.csv 60 > $O/launch_summary.txt 2>&1; head -30 $O/launch_summary.txt
