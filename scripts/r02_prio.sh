set -u
O=gpurun_out/r02s; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "stem_and_head or fprop_dgrad_wgrad" > $O/pytest_edge.log 2>&1; echo "pytest edge rc=$?"; tail -3 $O/pytest_edge.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -2 $O/pytest_unet.log
for e in A=1 DMU_LANE_PRIO=1 DMU_EDGE_WGRAD_CTAS=1 DMU_EDGE_WGRAD_CTAS=3; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
DMU_LANE_PRIO=1 timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q -k "trainstep or graph" > $O/pytest_prio.log 2>&1; echo "pytest prio rc=$?"; tail -2 $O/pytest_prio.log
DMU_LANE_PRIO=1 timeout 200 python scripts/step_trace.py $O/step_trace_prio.csv > $O/step_trace_prio.txt 2>&1; sed -n 3,6p $O/step_trace_prio.txt
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,6p $O/step_trace.txt
