#!/usr/bin/env python
"""Additional energy-model fixture from the LIVE reference (build container only; see make_golden.py for the protocol).

Stores every random draw of one ``EnergyBasedDiffusion.loss_function`` call (timesteps, q_sample noise, the Langevin
noises, the interpolation weights) so that the CUDA implementation can be driven with the same tensors on another device,
plus: the energies, the Langevin output x_fake, the input gradient of the energy at the interpolate, the loss with the
shipped regularization weight, and loss + parameter gradients with regularization_weight = 0 (contrastive-divergence only).
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import weights as W  # noqa: E402


def main(ref="/root/reference"):
    sys.path.insert(0, ref)
    from models.energy_based import EnergyBasedDiffusion  # noqa

    class RepairedEnergy(EnergyBasedDiffusion):      # SURVEY.md §8(c) repairs, as in make_golden.py
        def forward(self, x, t=None):
            return self.model(x)

        def generate_samples(self, batch_size, device):
            return self.sample(batch_size, device)

        def _langevin_sampling(self, x, t):
            x.requires_grad_(True)
            for _ in range(self.langevin_steps):
                energy = self.forward(x, t)
                grad = torch.autograd.grad(energy.sum(), x)[0]
                noise = torch.randn_like(x)
                x = x - self.langevin_step_size * grad + math.sqrt(2 * self.langevin_step_size) * noise
                x = x.detach().requires_grad_(True)
            return x.detach()

    torch.set_num_threads(8)
    out = {}
    for tag, EC, R in (("c16", 16, 16), ("c64", 64, 32)):
        cfg = {"num_timesteps": 1000, "beta_start": 1e-4, "beta_end": 0.02, "use_time_conditioning": False, "in_channels": 3,
               "model_channels": EC, "image_size": R, "image_channels": 3, "loss_type": "energy_based", "energy_scale": 1.0,
               "regularization_weight": 0.01, "langevin_steps": 3, "langevin_step_size": 0.01}
        B = 4
        xe = torch.randn(B, 3, R, R, generator=torch.Generator().manual_seed(611))
        rec = {"C": EC, "R": R, "wseed": 62, "cfg": cfg, "x": xe}
        for lam, key in ((0.01, "gp"), (0.0, "cd")):
            c = dict(cfg, regularization_weight=lam)
            em = RepairedEnergy(c)
            sd = W.make_state_dict(W.energynet_param_spec(EC, 3, "model."), 62)
            full = em.state_dict(); full.update(sd); em.load_state_dict(full, strict=True)
            torch.manual_seed(94)
            # replay of loss_function's RNG order to record the draws
            t = torch.randint(0, 1000, (B,))
            noise = torch.randn_like(xe)
            lang = [torch.randn_like(xe) for _ in range(3)]
            alpha = torch.rand(B, 1, 1, 1)
            torch.manual_seed(94)
            loss = em.loss_function(xe)
            loss.backward()
            rec.update({"t": t, "noise": noise, "lang": lang, "alpha": alpha})
            rec["loss_" + key] = loss.detach()
            if key == "cd":
                rec["grads_cd"] = {k: p.grad.clone() for k, p in em.named_parameters()}
            else:
                rec["grads_gp"] = {k: p.grad.clone() for k, p in em.named_parameters()}
        em.zero_grad()
        with torch.no_grad():
            rec["energy"] = em.forward(xe)
        x_noisy = em._add_noise(xe, t, noise)
        torch.manual_seed(94)
        torch.randint(0, 1000, (B,)); torch.randn_like(xe)          # advance to the Langevin draws
        rec["x_fake"] = em._langevin_sampling(x_noisy.clone(), t)
        xh = (alpha * xe + (1 - alpha) * rec["x_fake"]).requires_grad_(True)
        e = em.forward(xh)
        rec["grad_x_hat"] = torch.autograd.grad(e.sum(), xh)[0].detach()
        for key in ("grads_cd", "grads_gp"):       # keep the fixture small: full tensors only where they are small
            rec[key + "_norm"] = {k: float(v.norm()) for k, v in rec[key].items()}
            rec[key] = {k: v for k, v in rec[key].items() if v.numel() <= 20000}
        out[tag] = rec
        print(tag, "loss gp", float(rec["loss_gp"]), "loss cd", float(rec["loss_cd"]))
    torch.save(out, os.path.join(HERE, "energy_draws.pt"))


if __name__ == "__main__":
    main()
