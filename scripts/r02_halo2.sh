set -u
O=gpurun_out/r02h2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gn_statistics_epilogue or halo or fused_groupnorm or head_conv" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 120 python scripts/halo_timeline.py 2>&1 | head -12
for e in A=1 DMU_HALO_STATS_RUN_SH=0 DMU_GN_STATS=0; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench train $e', round(d['value']), d['ms_per_step'], d['roofline']['largest_launch']['us'])"
env $e timeout 300 python bench.py --workload ddim --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench ddim $e', round(d['value']), d['ms_per_step'])"
done
