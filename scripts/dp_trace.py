"""Development aid: CUPTI timeline of one DATA-PARALLEL training step on rank 0 (torchrun --nproc-per-node 2 scripts/dp_trace.py out.csv)."""
import collections, json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
import diffusion_model_universal_b200 as D
from diffusion_model_universal_b200.trainer import TrainStep
from bench import model_config, reseed_zero_init
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dp_trace.csv"
torch.manual_seed(1234)
m = D.DDPM(model_config(32, "bf16")); reseed_zero_init(m, 7); m.cuda()
ts = TrainStep(m)
g = torch.Generator().manual_seed(1234 + rank)
xs = [torch.randn(128, 3, 32, 32, generator=g).cuda() for _ in range(4)]
for i in range(8): ts.step(xs[i % 4])
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3): ts.step(xs[i % 4])
    torch.cuda.synchronize()
if rank == 0:
    path = os.path.join(tempfile.mkdtemp(), "t.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    adam = [i for i, e in enumerate(ev) if "adam_ema" in e["name"]]
    step = ev[adam[-2] + 1: adam[-1] + 1]
    t0 = step[0]["ts"]
    short = lambda n: n.split("(")[0].replace("void ", "").replace("dmu::", "").replace("__nv_bfloat16", "bf16")[:60]
    with open(out, "w") as f:
        f.write("name,stream,start_us,dur_us,grid\n")
        for e in step:
            a = e.get("args", {})
            f.write(f"\"{short(e['name'])}\",{a.get('stream', '')},{e['ts'] - t0:.2f},{e['dur']:.2f},\"{a.get('grid', '')}\"\n")
    end = max(e["ts"] + e["dur"] for e in step)
    print(f"{len(step)} kernels, step span {end - t0:.1f} us; previous adam end {ev[adam[-2]]['ts'] + ev[adam[-2]]['dur'] - t0:.1f}")
    for e in step:
        if "nccl" in e["name"].lower() or "adam" in e["name"] or "repack" in e["name"] or "gn_param" in e["name"] or "loss_final" in e["name"] or "stem_tc" in e["name"]:
            print(f"  {short(e['name']):50s} start {e['ts'] - t0:8.1f} dur {e['dur']:7.1f} end {e['ts'] + e['dur'] - t0:8.1f}")
dist.barrier(); dist.destroy_process_group()
