set -u
for e in DMU_REPACK_GX=64 DMU_REPACK_GX=128 DMU_REPACK_GX=256 DMU_REPACK_GX=32; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench train $e', round(d['value']), d['ms_per_step'], d['roofline']['by_entry_point_ms'].get('dmu_repack_weights'))"
done
