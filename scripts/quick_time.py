"""Ad-hoc timing of forward / train step (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from oracle import weights as W

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for precision in ("bf16", "fp32"):
        cfg = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": 64,
               "loss_type": "mse", "precision": precision}
        m = D.DDPM(cfg)
        sd = m.state_dict(); sd.update(W.make_state_dict(W.unet_param_spec(64, 3, "model."), 1)); m.load_state_dict(sd)
        m.cuda()
        x = torch.randn(B, 3, 32, 32, device="cuda")
        t = torch.randint(0, 1000, (B,), device="cuda")
        def timeit(fn, n=5):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time(); e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n, (time.time() - t0) / n * 1e3
        def fwd():
            with torch.no_grad(): m(x, t)
        def train():
            m.zero_grad(set_to_none=True)
            m.loss_function(x).backward()
        f = timeit(fwd); tr = timeit(train)
        print(f"{precision} B={B}: fwd {f[0]:.2f} ms (wall {f[1]:.2f})  train {tr[0]:.2f} ms (wall {tr[1]:.2f})  -> {B/tr[0]*1e3:.0f} img/s", flush=True)
        # per-op timing of the forward plan
        eng = m.model.engine
        plan = eng.get_plan(x.shape, False)
        from diffusion_model_universal_b200 import ops
        stream = ops._stream()
        acc = {}
        for fn, args in plan.fwd:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(*args, stream); e1.record(); torch.cuda.synchronize()
            acc[fn.__name__] = acc.get(fn.__name__, 0) + e0.elapsed_time(e1)
        print("  fwd by op (ms):", {k: round(v, 3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])}, flush=True)

main()
