"""HBM roofline of the fused per-step update / loss / optimizer kernels (north star: "achieved HBM GB/s against the B200 peak for
the norm, elementwise and update kernels").  Each kernel is timed back to back with CUDA events over ROTATING buffer sets whose
total footprint exceeds the 126 MB L2 (so every launch streams from HBM), at the sizes of BASELINE configs[2] (256 x 3 x 64 x 64 per
rank: 12.6 MB per tensor) and at a 16x larger batch where launch latency no longer matters.  Algorithmic bytes per element are
SURVEY.md §8(d)'s.  Writes one JSON object to stdout."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from oracle import process as P

dev = torch.device("cuda:0")
PEAK = 6550.0
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fns, reps=5):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for f in fns:
            f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(fns)) * 1e3      # us per launch


def sets_for(nbytes_per_set):
    return max(2, int(400e6 // nbytes_per_set) + 1)


out = {"peak_gbs": PEAK, "method": "CUDA events around back-to-back launches over rotating buffer sets (> 126 MB in total), us per launch", "kernels": []}
betas, alphas, acp = (t.to(dev) for t in P.linear_schedule(1e-4, 0.02, 1000))
tabs = [t.to(dev) for t in P.ddim_tables(acp.cpu(), P.ddim_timesteps(1000, 50), 0.5)]
sig = P.score_sigma_ladder(0.01, 50.0, 10).to(dev)

for B, R in ((256, 64), (4096, 64), (128, 32)):
    n = B * 3 * R * R
    K = sets_for(4 * n * 4)
    bufs = [[torch.randn(B, 3, R, R, device=dev) for _ in range(4)] for _ in range(K)]
    t = torch.randint(1, 1000, (B,), device=dev)
    idx = torch.randint(0, 50, (B,), device=dev)
    w = torch.rand(B, device=dev)
    cases = [
        ("dmu_q_sample", 12, [lambda b=b: ops.q_sample(b[0], t, b[1], acp, out=b[2]) for b in bufs]),
        ("dmu_ddpm_step (t > 0)", 16, [lambda b=b: ops.ddpm_step(b[0], b[1], t, b[2], betas, alphas, acp, out=b[3]) for b in bufs]),
        ("dmu_ddim_step (eta 0)", 12, [lambda b=b: ops.ddim_step(b[0], b[1], idx, None, *tabs, out=b[3]) for b in bufs]),
        ("dmu_ddim_step (eta > 0)", 16, [lambda b=b: ops.ddim_step(b[0], b[1], idx, b[2], *tabs, out=b[3]) for b in bufs]),
        ("dmu_langevin_score_step", 16, [lambda b=b: ops.langevin_score_step(b[0], b[1], b[2], sig, 5, 1.0, out=b[3]) for b in bufs]),
        ("dmu_langevin_energy_step", 16, [lambda b=b: ops.langevin_energy_step(b[0], b[1], b[2], 0.01, out=b[3]) for b in bufs]),
        ("dmu_diffusion_loss (+ dL/dpred)", 12, [lambda b=b: ops.diffusion_loss(b[0], b[1], w, 1.0, 0.0, 0.0, 1.0, True, dpred_out=b[2]) for b in bufs]),
    ]
    for name, bpe, fns in cases:
        us = timeit(fns)
        gbs = bpe * n / us / 1e3
        out["kernels"].append({"kernel": name, "shape": [B, 3, R, R], "bytes_per_element": bpe, "algorithmic_mb": round(bpe * n / 1e6, 2), "us": round(us, 2),
                               "gb_per_s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 3), "buffer_sets": K})
    del bufs
    torch.cuda.empty_cache()

# fused Adam + EMA over the UNet's arena (15.9 M parameters: 40 B per parameter), rotating over three arena sets
n = 15_909_955 // 4 * 4
sets = [[torch.randn(n, device=dev) for _ in range(5)] for _ in range(3)]
lib = _abi.lib()
step = torch.ones((), device=dev, dtype=torch.int64)
s = ops._stream()
fns = [lambda a=a: lib.dmu_adam_ema(a[0].data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), a[4].data_ptr(), n, 2e-4, 0.9, 0.999, 1e-8, 0.0, 1, 0.9999,
                                    1.0, step.data_ptr(), s) for a in sets]
for a in sets:
    a[3].abs_()
us = timeit(fns)
out["kernels"].append({"kernel": "dmu_adam_ema", "shape": [n], "bytes_per_element": 40, "algorithmic_mb": round(40 * n / 1e6, 1), "us": round(us, 2),
                       "gb_per_s": round(40 * n / us / 1e3, 1), "frac_of_hbm_peak": round(40 * n / us / 1e3 / PEAK, 3), "buffer_sets": 3})
print(json.dumps(out))
