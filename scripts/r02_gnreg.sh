set -u
O=${1:-gpurun_out/r02p}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gn_epilogue" > $O/pytest_epi.log 2>&1; echo "pytest epi rc=$?"; tail -4 $O/pytest_epi.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -3 $O/pytest_unet.log
for e in A=1 DMU_GN_PIX256=0; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $O/bench_$e.json 2> $O/bench_$e.err; python -c "
import json
for l in open('$O/bench_$e.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,14p $O/step_trace.txt
