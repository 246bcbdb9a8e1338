set -u
O=gpurun_out/r02o; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gn_epilogue" > $O/pytest_epi.log 2>&1; echo "pytest epi rc=$?"; tail -5 $O/pytest_epi.log
timeout 300 python scripts/gn_epi_timeline.py > $O/timeline.txt 2>&1; cat $O/timeline.txt
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -x -q > $O/pytest_unet.log 2>&1; echo "pytest unet rc=$?"; tail -3 $O/pytest_unet.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-extras > $O/bench.json 2> $O/bench.err; python -c "
import json
for l in open('$O/bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', round(d['value']), d['ms_per_step'])"
timeout 200 python scripts/step_trace.py $O/step_trace.csv > $O/step_trace.txt 2>&1; sed -n 3,12p $O/step_trace.txt
