"""One optimisation step of the reference trainer's hot loop on the CUDA engine.

Mirror of trainers/ddpm_trainer.py:539-555 (``images.to(device)`` ->
``zero_grad`` -> ``loss_function`` -> ``backward`` -> ``optimizer.step`` ->
``_update_ema_model``) with the three host-side costs removed:
  * gradients live in one flat arena, averaged across ranks by bucketed
    asynchronous all-reduces (the reference's DDP wrapper never reduces:
    trainers/ddpm_trainer.py:130-136 vs :543-547);
  * Adam over 314 tensors + the 314-iteration Python EMA loop become one
    ``dmu_adam_ema`` launch over the arena;
  * the loss stays on the device — callers read it when they want it.
"""

from typing import Optional

import torch

from .optim import FusedAdamEMA
from .parallel import GradAllReducer


class TrainStep:
    def __init__(self, model, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_decay: Optional[float] = 0.9999, bucket_mb: float = 16.0, group=None):
        self.model = model
        self.opt = FusedAdamEMA(model.model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay)
        self.reducer = GradAllReducer(model.model, bucket_mb=bucket_mb, group=group)
        self._stage = None

    def step(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [B,C,H,W] on the device, or a (pinned) host tensor which is
        copied asynchronously first (trainers/ddpm_trainer.py:539).  Returns the 0-dim
        device loss tensor of this rank (no host sync)."""
        if not images.is_cuda:
            dev = next(self.model.parameters()).device
            if self._stage is None or self._stage.shape != images.shape:
                self._stage = torch.empty(images.shape, device=dev, dtype=torch.float32)
            self._stage.copy_(images, non_blocking=True)
            images = self._stage
        loss = self.model.loss_function(images)
        loss.backward()
        scale = self.reducer.allreduce()
        self.opt.step(grad_scale=scale)
        # the arena holds this step's gradients; param.grad views must not accumulate into the next step
        for p in self.model.parameters():
            p.grad = None
        return loss.detach()
