"""Score-based model behind the reference API (models/score_based.py) on the CUDA engine.

The reference class cannot run as shipped (abstract ``generate_samples`` is
missing and ``ScoreNet.forward`` calls layers that do not exist,
score_based.py:84-99,209); this class implements the documented intent with
the repairs of SURVEY.md §8(c): the UNet body conditioned on
``time_embed(log sigma)``, ``generate_samples = sample``.
"""

from typing import Dict

import numpy as np
import torch

from . import ops
from .base_model import BaseDiffusion
from .losses import DiffusionLoss, _LossFn
from .unet import UNet


class ScoreNet(UNet):
    """score_based.py:25-61: UNet + ``time_embed`` MLP on log(sigma).  The inherited
    ``time_embedding`` parameters stay in the state_dict (dead, as in the reference)."""

    def __init__(self, in_channels: int, model_channels: int, out_channels: int, num_scales: int = 1000, precision: str = "fp32"):
        super().__init__(in_channels, model_channels, out_channels, precision=precision, sigma_embed=True)


class ScoreBasedDiffusion(BaseDiffusion):
    def __init__(self, config: Dict):
        super().__init__(config)
        self.sigma_min = config.get("sigma_min", 0.01)
        self.sigma_max = config.get("sigma_max", 50.0)
        self.num_scales = config.get("num_scales", 1000)
        self.beta = config.get("beta", 1.0)
        self.model = ScoreNet(in_channels=config.get("in_channels", 3), model_channels=config.get("model_channels", 64),
                              out_channels=config.get("in_channels", 3), num_scales=self.num_scales,
                              precision=config.get("precision", "fp32"))
        loss_type = config.get("loss_type", "score_matching")
        self.loss_type = loss_type
        self.loss_fn = None if loss_type == "score_matching" else DiffusionLoss(loss_type=loss_type, loss_config=config.get("loss_config", None))

    def forward(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        return self.model(x, sigma.float())

    def loss_function(self, x: torch.Tensor) -> torch.Tensor:
        """score_based.py:180-207 + utils/losses.py:226-242.  RNG order: rand, randn_like, then the loss's own randn_like."""
        batch_size = x.shape[0]
        u = torch.rand(batch_size, device=x.device)
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** u
        noise = torch.randn_like(x)
        noisy_x = ops.scale_add(x, noise, None, sigma)                       # x + sigma*noise
        score = self.forward(noisy_x, sigma)
        if self.loss_fn is not None:
            return self.loss_fn(score, noise, sigma)
        fresh = torch.randn_like(x)                                          # losses.py:238 draws fresh noise
        target = ops.scale_add(fresh, fresh, torch.zeros_like(sigma), -1.0 / sigma)   # -fresh/sigma
        return _LossFn.apply(score, target, None, 1.0, 0.0, 0.0, 1.0)        # F.mse_loss

    def sample(self, batch_size: int, device: torch.device) -> torch.Tensor:
        """score_based.py:209-247 — annealed Langevin dynamics, one fused update per inner step."""
        x = torch.randn((batch_size, self.config["in_channels"], self.config["image_size"], self.config["image_size"]), device=device)
        sigmas = torch.exp(torch.linspace(np.log(self.sigma_max), np.log(self.sigma_min), self.num_scales, device=device))
        sigma_rows = sigmas[:, None].expand(-1, batch_size).contiguous()
        eng = self.model.engine
        with torch.no_grad():
            prev = eng.frozen
            try:
                for k in range(self.num_scales):
                    for _ in range(self.config.get("langevin_steps", 10)):
                        score = self.forward(x, sigma_rows[k])
                        eng.frozen = True
                        noise = torch.randn_like(x)
                        x = ops.langevin_score_step(x, score, noise, sigmas, k, float(self.beta))
            finally:
                eng.frozen = prev
        return x

    def generate_samples(self, batch_size: int, device: torch.device) -> torch.Tensor:
        return self.sample(batch_size, device)

    def _get_sigma(self, t: torch.Tensor) -> torch.Tensor:
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** (t.float() / self.num_scales)
