set -u
O=gpurun_out/r02t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stem_and_head or fprop_dgrad_wgrad" > $O/pytest_edge.log 2>&1; echo "pytest edge rc=$?"; tail -2 $O/pytest_edge.log
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -2 $O/pytest_unet.log
for e in A=1 DMU_FRONT_PRIO=0 DMU_EDGE_WGRAD_CTAS=1; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,6p $O/step_trace.txt
