import os, sys
sys.path.insert(0, "/root/repo")
import torch
import diffusion_model_universal_b200 as D
from bench import model_config, reseed_zero_init
m = D.DDPM(model_config(32, "bf16")); reseed_zero_init(m, 7); m.cuda()
x = torch.randn(128, 3, 32, 32, device="cuda"); t = torch.randint(0, 1000, (128,), device="cuda")
eng = m.model.engine
eng.frozen = True
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    eng.use_graphs = False
    print("eager fwd (PDL=%s): %.3f ms" % (os.environ.get("DMU_PDL", "1"), timeit(lambda: m(x, t))))
    eng.use_graphs = True
    print("graph fwd (PDL=%s): %.3f ms" % (os.environ.get("DMU_PDL", "1"), timeit(lambda: m(x, t))))
