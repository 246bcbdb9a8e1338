run() { env "$@" python bench.py --steps 30 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);b=d['roofline']['by_entry_point_ms'];print('$*', round(d['value']),d['ms_per_step'], {k:b[k] for k in list(b)[:6]})"; }
rund() { env "$@" python bench.py --workload ddim --steps 3 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$*', 'ddim50', round(d['value'],1), d['ms_per_step'])"; }
