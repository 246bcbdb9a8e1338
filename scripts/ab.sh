run() { env "$@" python bench.py --steps 30 --warmup 5 --no-cpu --no-ddim 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);b=d['roofline']['by_entry_point_ms'];print('$*', round(d['value']),d['ms_per_step'], {k:b[k] for k in list(b)[:6]})"; }
rund() { env "$@" python bench.py --steps 10 --warmup 8 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$*', round(d['value']), round(d['extra']['ddim50_64x64']['img_per_s'],1))"; }
