# A/B of the graph-structure changes (auxiliary lanes, early memsets, end-of-backward order), PDL on/off, repack placement
set -u
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_unet.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d.get('gpu_launches_per_step'))
"; }
run A=1 >> $O/ab.txt 2>&1
run DMU_AUX_LANES=0 >> $O/ab.txt 2>&1
run DMU_PDL=0 >> $O/ab.txt 2>&1
run DMU_REPACK_LANE=main >> $O/ab.txt 2>&1
run A=2 >> $O/ab.txt 2>&1
cat $O/ab.txt
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; head -8 $O/step_trace.txt | tail -5
