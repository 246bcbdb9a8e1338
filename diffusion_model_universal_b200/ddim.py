"""DDIM behind the reference API (models/ddim.py) on the CUDA engine."""

from typing import Dict, List, Optional

import torch

from . import ops
from .ddpm import DDPM


class DDIM(DDPM):
    def __init__(self, config: Dict):
        super().__init__(config)
        self.ddim_sampling_steps = config.get("ddim_sampling_steps", 50)
        self.ddim_discretize = config.get("ddim_discretize_method", "uniform")
        self.eta = config.get("eta", 0.0)
        self.ddim_timesteps = self._get_ddim_timesteps()
        self._precompute_ddim_sampling_parameters()

    def _get_ddim_timesteps(self) -> torch.Tensor:
        """ddim.py:49-65 (plain CPU tensor attribute, like the reference)."""
        if self.ddim_discretize == "uniform":
            c = self.num_timesteps // self.ddim_sampling_steps
            return torch.arange(0, self.num_timesteps, c)
        if self.ddim_discretize == "quad":
            ts = torch.linspace(0, torch.sqrt(torch.tensor(self.num_timesteps * .8)), self.ddim_sampling_steps) ** 2
            return ts.long()
        raise NotImplementedError(f"Unknown discretization method: {self.ddim_discretize}")

    def _precompute_ddim_sampling_parameters(self):
        """ddim.py:67-81 — same torch expressions, so the buffers are bit-identical."""
        alphas = self.alphas_cumprod[self.ddim_timesteps]
        alphas_prev = torch.cat([self.alphas_cumprod[0:1], self.alphas_cumprod[self.ddim_timesteps[:-1]]])
        sigmas = self.eta * torch.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        self.register_buffer("ddim_alphas", alphas)
        self.register_buffer("ddim_alphas_prev", alphas_prev)
        self.register_buffer("ddim_sigmas", sigmas)
        self.register_buffer("ddim_sqrt_one_minus_alphas", torch.sqrt(1. - alphas))

    def _ddim_sample(self, x: torch.Tensor, t: torch.Tensor, t_prev: Optional[torch.Tensor] = None,
                     pred_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ddim.py:83-126 — ``t`` indexes the S-entry tables (what the arithmetic
        of ddim.py:97-100 requires); one fused launch after the eps prediction."""
        if pred_noise is None:
            pred_noise = self.forward(x, t)
        noise = torch.randn_like(x) if self.eta > 0 else None
        return ops.ddim_step(x, pred_noise, t, noise, self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sigmas,
                             self.ddim_sqrt_one_minus_alphas)

    def _loop(self, batch_size, device, save_interval):
        """ddim.py:128-187 with the table index repaired (the shipped loop feeds
        the raw timestep, e.g. 980, into the 50-entry tables: IndexError)."""
        x = torch.randn(self._sample_shape(batch_size), device=device)
        out = [x.clone()] if save_interval is not None else []
        S = len(self.ddim_timesteps)
        ts = self.ddim_timesteps.to(device=device, dtype=torch.int64)[:, None].expand(-1, batch_size).contiguous()
        idx = torch.arange(S, device=device, dtype=torch.int64)[:, None].expand(-1, batch_size).contiguous()
        eng = self.model.engine
        with torch.no_grad():
            prev_frozen = eng.frozen
            try:
                for i in range(S - 1, -1, -1):
                    eps = self.forward(x, ts[i])
                    eng.frozen = True
                    x = self._ddim_sample(x, idx[i], None, pred_noise=eps)
                    if save_interval is not None and (i % save_interval == 0 or i == 0):
                        out.append(x.clone())
            finally:
                eng.frozen = prev_frozen
        return out if save_interval is not None else [x]

    def generate_samples(self, batch_size: int, device: torch.device) -> torch.Tensor:
        return self._loop(batch_size, device, None)[-1]

    def generate_samples_with_intermediates(self, batch_size: int, device: torch.device, save_interval: int = 2) -> List[torch.Tensor]:
        return self._loop(batch_size, device, save_interval)
