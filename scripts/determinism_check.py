"""Development aid: run-to-run spread of the UNet forward (eager vs eager vs graph replay), per precision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_model_universal_b200 as D
from oracle import weights as W, unet as U


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for precision in ("fp32", "bf16"):
    cfg = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": 64, "loss_type": "mse", "precision": precision}
    m = D.DDPM(cfg)
    wsd = W.make_state_dict(W.unet_param_spec(64, 3, "model."), 1)
    sd = m.state_dict(); sd.update(wsd); m.load_state_dict(sd); m.cuda()
    g = torch.Generator().manual_seed(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128      # 128: the batch at which the statistics ride in the persistent 3x3 kernel
    x = torch.randn(B, 3, 32, 32, generator=g).cuda(); t = torch.randint(0, 1000, (B,), generator=g).cuda()
    eng = m.model.engine
    with torch.no_grad():
        eng.use_graphs = False
        a = m(x, t); b = m(x, t)
        eng.use_graphs = True
        c = m(x, t); d = m(x, t); e = m(x, t)
        ref = U.unet_forward({k: v.cuda() for k, v in wsd.items()}, x, t)
    print(f"{precision}: eager/eager {rel(b, a):.2e}  graph/eager {rel(c, a):.2e}  graph/graph {rel(e, d):.2e}  "
          f"eager/oracle {rel(a, ref):.2e}  graph/oracle {rel(d, ref):.2e}", flush=True)
