set -u
O=gpurun_out/r02r; mkdir -p $O
DMU_EDGE_WGRAD_TC=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stem_and_head or fprop_dgrad_wgrad" > $O/pytest_edge.log 2>&1; echo "pytest edge rc=$?"; tail -5 $O/pytest_edge.log
DMU_EDGE_WGRAD_TC=1 timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -3 $O/pytest_unet.log
for e in A=1 DMU_EDGE_WGRAD_TC=1 "DMU_EDGE_WGRAD_TC=1 DMU_EDGE_WGRAD_CTAS=1"; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
DMU_EDGE_WGRAD_TC=1 timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,6p $O/step_trace.txt; grep -i "edge" $O/step_trace.txt
