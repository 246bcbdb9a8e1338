"""Target program of the ncu passes: warm up, then run the profiled region (cudaProfilerStart / Stop) only:
   python scripts/ncu_step.py train [steps]   - training steps (B = 128, 32x32: the bench's step graph)
   python scripts/ncu_step.py ddim            - UNet evaluations of the DDIM config (B = 256, 64x64)
Use with `ncu --profile-from-start off ...` (scripts/r02_evidence.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from diffusion_model_universal_b200.trainer import TrainStep
from bench import model_config, reseed_zero_init
what = sys.argv[1] if len(sys.argv) > 1 else "train"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if what == "train":
    m = D.DDPM(model_config(32, "bf16")); reseed_zero_init(m, 7); m.cuda()
    ts = TrainStep(m)
    xs = [torch.randn(128, 3, 32, 32, device="cuda") for _ in range(4)]
    for i in range(8): ts.step(xs[i % 4])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(n): ts.step(xs[i % 4])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    m = D.DDIM(model_config(64, "bf16")); reseed_zero_init(m, 7); m.cuda()
    x = torch.randn(256, 3, 64, 64, device="cuda"); t = torch.randint(0, 1000, (256,), device="cuda")
    with torch.no_grad():
        for _ in range(4): m(x, t)
        m.model.engine.frozen = True
        for _ in range(3): m(x, t)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(n): m(x, t)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
