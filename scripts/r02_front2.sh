set -u
O=gpurun_out/r02u; mkdir -p $O
for e in "DMU_REPACK_LANE=main DMU_FRONT_PRIO=0" "DMU_REPACK_LANE=main DMU_FRONT_PRIO=1" "DMU_REPACK_LANE=side DMU_FRONT_PRIO=0"; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench $e', round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))"
done
DMU_REPACK_LANE=main DMU_FRONT_PRIO=0 timeout 200 python scripts/step_trace.py $O/step_trace_a.csv > $O/step_trace_a.txt 2>&1; sed -n 3,6p $O/step_trace_a.txt; head -16 $O/step_trace_a.csv
DMU_REPACK_LANE=main DMU_FRONT_PRIO=1 timeout 200 python scripts/step_trace.py $O/step_trace_b.csv > $O/step_trace_b.txt 2>&1; sed -n 3,6p $O/step_trace_b.txt; head -16 $O/step_trace_b.csv
