"""Development aid: the REAL timeline of the one-graph training step (both lanes running concurrently), from CUPTI through
torch.profiler.  Writes one CSV row per kernel of the last traced step: name, stream, start (us from the step's first kernel),
duration (us); and prints the per-family busy time, the idle time of the device (no kernel running) and the longest gaps.
Usage: python scripts/step_trace.py [out.csv] [B] [R]"""
import collections, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from diffusion_model_universal_b200.trainer import TrainStep
from bench import model_config, reseed_zero_init

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_trace.csv"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
R = int(sys.argv[3]) if len(sys.argv) > 3 else 32
m = D.DDPM(model_config(R, "bf16")); reseed_zero_init(m, 7); m.cuda()
ts = TrainStep(m)
xs = [torch.randn(B, 3, R, R, device="cuda") for _ in range(4)]
for i in range(6): ts.step(xs[i % 4])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3): ts.step(xs[i % 4])
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
# the last step = kernels after the second-to-last adam_ema launch
adam = [i for i, e in enumerate(ev) if "adam_ema" in e["name"]]
step = ev[adam[-2] + 1: adam[-1] + 1] if len(adam) >= 2 else ev
t0 = step[0]["ts"]
def short(n):
    return n.split("(")[0].replace("void ", "").replace("dmu::", "").replace("__nv_bfloat16", "bf16")[:60]
with open(out, "w") as f:
    f.write("name,stream,start_us,dur_us,grid\n")
    for e in step:
        a = e.get("args", {})
        f.write(f"\"{short(e['name'])}\",{a.get('stream', '')},{e['ts'] - t0:.2f},{e['dur']:.2f},\"{a.get('grid', '')}\"\n")
end = max(e["ts"] + e["dur"] for e in step)
print(f"{len(step)} kernels, step span {end - t0:.1f} us")
if len(adam) >= 2:
    pa = ev[adam[-2]]
    print(f"previous step's adam_ema: start {pa['ts'] - t0:.1f} us, end {pa['ts'] + pa['dur'] - t0:.1f} us relative to this step's first kernel")
# device idle time: union of kernel intervals
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s; gaps.append((s - cur_e, cur_e - t0)); cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"device busy (>= 1 kernel running) {busy:.1f} us, idle {end - t0 - busy:.1f} us in {len(gaps)} gaps; largest: " + ", ".join(f"{g:.1f}@{at:.0f}" for g, at in sorted(gaps, reverse=True)[:8]))
fam = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    fam[short(e["name"])][0] += 1; fam[short(e["name"])][1] += e["dur"]
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {k:60s} n={v[0]:3d} tot={v[1]:8.1f} avg={v[1] / v[0]:6.1f}")
streams = collections.defaultdict(float)
for e in step: streams[e.get("args", {}).get("stream")] += e["dur"]
print("busy per stream:", {k: round(v, 1) for k, v in streams.items()})
