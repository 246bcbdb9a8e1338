set -u
O=gpurun_out/r02x; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gn_statistics_epilogue or halo" > $O/pytest_stats.log 2>&1; echo "pytest stats rc=$?"; tail -4 $O/pytest_stats.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -2 $O/pytest_unet.log
for e in A=1 DMU_GN_STATS=0; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench train $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'), d['roofline']['frac'])"
env $e timeout 300 python bench.py --workload ddim --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench ddim $e', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"
done
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,14p $O/step_trace.txt
timeout 300 python scripts/eval_trace.py $O/ddim_eval.csv 256 64 > $O/ddim_eval.txt 2>&1; sed -n 3,14p $O/ddim_eval.txt
