"""DDPM behind the reference API (models/ddpm.py:137-329) on the CUDA engine."""

from typing import Dict, List, Optional

import torch

from . import ops
from .base_model import BaseDiffusion
from .losses import DiffusionLoss
from .unet import UNet


class DDPM(BaseDiffusion):
    """Same constructor contract as the reference: ``DDPM(config['model_config'])``.

    Extra optional key (defaults to reference behaviour): ``precision`` in
    {"fp32", "bf16"} — the arithmetic type of the denoiser's activations and
    conv operands (accumulation is always fp32)."""

    def __init__(self, config: Dict):
        super().__init__(config)
        self.beta_start = config.get("beta_start", 1e-4)
        self.beta_end = config.get("beta_end", 1e-2)
        self.num_timesteps = config.get("num_timesteps", 1000)
        # ddpm.py:176-178 — built with the same torch ops so the tables are bit-identical
        self.register_buffer("betas", torch.linspace(self.beta_start, self.beta_end, self.num_timesteps))
        self.register_buffer("alphas", 1 - self.betas)
        self.register_buffer("alphas_cumprod", torch.cumprod(self.alphas, dim=0))
        self.model = UNet(in_channels=config.get("in_channels", 3), model_channels=config.get("model_channels", 64),
                          out_channels=config.get("in_channels", 3), precision=config.get("precision", "fp32"))
        self.loss_fn = DiffusionLoss(loss_type=config.get("loss_type", "mse"), loss_config=config.get("loss_config", {}))
        self.loss_fn.max_t = self.num_timesteps    # t < num_timesteps: lets the 'snr' time weights stay on the device

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return self.model(x, t)

    def loss_function(self, x: torch.Tensor) -> torch.Tensor:
        """ddpm.py:207-235 — same RNG call order: randint, then randn_like."""
        batch_size = x.shape[0]
        t = torch.randint(0, self.num_timesteps, (batch_size,), device=x.device)
        noise = torch.randn_like(x)
        noisy_x = self._add_noise(x, t, noise)
        noise_pred = self.forward(noisy_x, t)
        return self.loss_fn(noise_pred, noise, t)

    def _sample_shape(self, batch_size):
        return (batch_size, self.config["image_channels"], self.config["image_size"], self.config["image_size"])

    def generate_samples(self, batch_size: int, device: torch.device) -> torch.Tensor:
        """ddpm.py:237-255."""
        return self._sample_loop(batch_size, device, None)[-1]

    def generate_samples_with_intermediates(self, batch_size: int, device: torch.device, save_interval: int = 100) -> List[torch.Tensor]:
        """ddpm.py:257-284."""
        return self._sample_loop(batch_size, device, save_interval)

    def _sample_loop(self, batch_size, device, save_interval):
        x = torch.randn(self._sample_shape(batch_size), device=device)
        out = [x.clone()] if save_interval is not None else []
        # one device tensor of all timesteps instead of a host->device copy per step (ddpm.py:252)
        t_all = torch.arange(self.num_timesteps, device=device, dtype=torch.int64)[:, None].expand(-1, batch_size).contiguous()
        eng = self.model.engine
        with torch.no_grad():
            prev_frozen = eng.frozen
            try:
                for t in reversed(range(self.num_timesteps)):
                    x = self._reverse_diffusion_step(x, t_all[t], _t_host=t)
                    eng.frozen = True       # weights cannot change inside the loop: repack once
                    if save_interval is not None and (t % save_interval == 0 or t == 0):
                        out.append(x.clone())
            finally:
                eng.frozen = prev_frozen
        return out if save_interval is not None else [x]

    def _add_noise(self, x: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ddpm.py:286-296 as one fused launch."""
        if noise is None:
            noise = torch.randn_like(x)
        return ops.q_sample(x.contiguous(), t, noise, self.alphas_cumprod)

    def _reverse_diffusion_step(self, x: torch.Tensor, t: torch.Tensor, _t_host: Optional[int] = None) -> torch.Tensor:
        """ddpm.py:298-329: eps prediction, then one fused posterior-step launch.

        The reference branches on ``t[0] > 0`` with a host sync; callers that
        know the timestep pass it as ``_t_host`` so nothing syncs.  The noise
        is drawn with the same ``randn_like`` call (none at t == 0)."""
        noise_pred = self.forward(x, t)
        t0 = int(t[0].item()) if _t_host is None else _t_host
        noise = torch.randn_like(x) if t0 > 0 else None
        return ops.ddpm_step(x, noise_pred, t, noise, self.betas, self.alphas, self.alphas_cumprod)
