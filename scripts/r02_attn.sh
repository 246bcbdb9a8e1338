set -u
O=gpurun_out/r02y; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention_core" > $O/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -12 $O/pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -2 $O/pytest_unet.log
for e in A=1 DMU_ATTN_TC=0; do
env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench train $e', round(d['value']), d['ms_per_step'])"
env $e timeout 300 python bench.py --workload ddim --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('bench ddim $e', round(d['value']), d['ms_per_step'])"
done
timeout 300 python scripts/eval_trace.py $O/ddim_eval.csv 256 64 > $O/ddim_eval.txt 2>&1; grep -i "attn" $O/ddim_eval.txt
