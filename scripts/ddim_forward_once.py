"""Development aid: a few UNet evaluations at the DDIM config (64x64, batch 256, bf16) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from bench import model_config, reseed_zero_init
m = D.DDIM(model_config(64, "bf16")); reseed_zero_init(m, 7); m.cuda()
m.model.engine.use_graphs = False
x = torch.randn(256, 3, 64, 64, device="cuda"); t = torch.full((256,), 500, device="cuda")
with torch.no_grad():
    for _ in range(3):
        m(x, t)
torch.cuda.synchronize()
print("ok")
