"""Development aid: CUPTI timeline of ONE UNet evaluation of the sampling path (no-grad plan, frozen weights), e.g. the DDIM
config: python scripts/eval_trace.py out.csv 256 64"""
import collections, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from bench import model_config, reseed_zero_init
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/eval_trace.csv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R = int(sys.argv[3]) if len(sys.argv) > 3 else 64
m = D.DDIM(model_config(R, "bf16")); reseed_zero_init(m, 7); m.cuda()
m.model.engine.frozen = False
x = torch.randn(B, 3, R, R, device="cuda"); t = torch.randint(0, 1000, (B,), device="cuda")
with torch.no_grad():
    for _ in range(4): m(x, t)
    m.model.engine.frozen = True
    for _ in range(3): m(x, t)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2): m(x, t)
        torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
half = [i for i, e in enumerate(ev) if "sinusoidal" in e["name"]]
step = ev[half[-1]:] if half else ev
t0 = step[0]["ts"]
short = lambda n: n.split("(")[0].replace("void ", "").replace("dmu::", "").replace("__nv_bfloat16", "bf16")[:60]
pe = t0
agg = collections.defaultdict(lambda: [0, 0.0])
with open(out, "w") as f:
    f.write("name,stream,start_us,dur_us,grid,d_end\n")
    for e in step:
        a = e.get("args", {}); end = e["ts"] + e["dur"]
        de = max(0.0, end - pe); pe = max(pe, end)
        agg[short(e["name"]) + " " + str(a.get("grid", ""))][0] += 1; agg[short(e["name"]) + " " + str(a.get("grid", ""))][1] += de
        f.write(f"\"{short(e['name'])}\",{a.get('stream', '')},{e['ts'] - t0:.2f},{e['dur']:.2f},\"{a.get('grid', '')}\",{de:.2f}\n")
print(f"{len(step)} kernels, evaluation span {pe - t0:.1f} us (B={B}, {R}x{R})")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {k:72s} n={v[0]:3d} completion-increment sum {v[1]:8.1f} us")
