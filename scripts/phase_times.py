"""Development aid: in-graph time of the forward and backward replays and of the remaining step pieces (B=128, 32x32, bf16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from diffusion_model_universal_b200 import ops
from diffusion_model_universal_b200.trainer import TrainStep
from bench import model_config, reseed_zero_init

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
m = D.DDPM(model_config(R, "bf16")); reseed_zero_init(m, 7); m.cuda()
ts = TrainStep(m)
x = torch.randn(B, 3, R, R, device="cuda")
for _ in range(4): ts.step(x)
eng = m.model.engine
plan = eng.get_plan(x.shape, True)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("fwd graph   %.3f ms" % timeit(lambda: plan.graphs["fwd"].replay()))
print("bwd graph   %.3f ms" % timeit(lambda: eng.run_backward(plan, None)))
print("repack      %.3f ms" % timeit(lambda: eng.repack(ops._stream())))
print("adam        %.3f ms" % timeit(lambda: ts.opt.step()))
print("full step   %.3f ms" % timeit(lambda: ts.step(x)))
t = torch.randint(0, 1000, (B,), device="cuda")
eng.frozen = True
with torch.no_grad():
    for _ in range(3): m(x, t)
    print("infer fwd   %.3f ms (no-grad plan, frozen weights)" % timeit(lambda: m(x, t)))
