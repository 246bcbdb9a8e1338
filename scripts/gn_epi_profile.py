"""Development aid: per-launch device time of every conv / GroupNorm launch of the training plan, with and without the
GroupNorm-in-the-epilogue fusion (same layer order: line up the two lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from diffusion_model_universal_b200 import ops
from bench import model_config, reseed_zero_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128


def profile(fuse):
    torch.manual_seed(0)
    m = D.DDPM(model_config(32, "bf16")); reseed_zero_init(m, 7); m.cuda()
    eng = m.model.engine
    eng.fuse_gn_epi = fuse
    x = torch.randn(B, 3, 32, 32, device="cuda"); t = torch.randint(0, 1000, (B,), device="cuda")
    eng.use_graphs = False
    for _ in range(2):
        y = m(x, t); y.backward(torch.randn_like(y))
    plan = eng.get_plan(x.shape, True)
    stream = ops._stream()
    rows = []
    for which in ("fwd", "bwd"):
        lst = [(op[0], op[1]) for op in getattr(plan, which) if op[0] is not None]
        for rep in range(2):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(lst) + 1)]
            torch.cuda._sleep(int(60e6))
            evs[0].record()
            for i, (fn, a) in enumerate(lst):
                fn(*a, stream); evs[i + 1].record()
            torch.cuda.synchronize()
        for i, (fn, a) in enumerate(lst):
            us = evs[i].elapsed_time(evs[i + 1]) * 1e3
            nm = fn.__name__
            if nm == "dmu_conv2d":
                p = a[0]._obj
                rows.append((which, "conv", f"{p.N}x{p.Hi}x{p.Wi}x{p.Ck}->{p.Ho}x{p.Wo}x{p.Cj} k{p.R}s{p.stride}g{p.gather} gn{p.gn_fuse_mode}", us))
            elif nm in ("dmu_gn_forward", "dmu_gn_backward"):
                p = a[0]._obj
                rows.append((which, nm[4:], f"{p.N}x{p.H}x{p.W}x{p.C}", us))
    return rows, plan.gn_fused


a, fa = profile(False)
b, fb = profile(True)
print("fused counts", fa, fb)
for tag, rows in (("OFF", a), ("ON", b)):
    tot = {}
    for w, k, d, us in rows:
        tot[(w, k)] = tot.get((w, k), 0) + us
    print(tag, {k: round(v) for k, v in tot.items()})
# per layer: walk both lists; in ON the GN rows of fused layers are missing
print("---- ON list (conv rows with gn1/gn2 are fused launches)")
for w, k, d, us in b:
    if "gn1" in d or "gn2" in d:
        print(f"  {w} {k:12s} {d:48s} {us:7.1f}")
print("---- OFF list, small layers")
for w, k, d, us in a:
    print(f"  {w} {k:12s} {d:48s} {us:7.1f}")
