"""Fused Adam + EMA over the flat parameter arena.

Replaces ``optimizer.step()`` over 314 tensors plus the per-parameter Python
EMA loop of the reference trainer (trainers/ddpm_trainer.py:139-143,463-480)
with ONE launch of ``dmu_adam_ema`` (28 + 12 bytes per parameter of HBM
traffic).  Semantics are torch.optim.Adam's (bias-corrected, eps added after
the sqrt, optional L2 weight decay) followed by
``ema = decay * ema + (1 - decay) * param``.
"""

import torch

from . import _abi, ops


class FusedAdamEMA:
    def __init__(self, unet, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=None):
        self.unet = unet
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.step_count = 0
        self.m = self.v = self.ema = None

    def _ensure(self):
        eng = self.unet.engine
        if eng.flat is None:
            raise RuntimeError("FusedAdamEMA: run a forward pass first (the parameter arena is created lazily on the device)")
        if self.m is None or self.m.data_ptr() == 0 or self.m.numel() != eng.flat.numel() or self.m.device != eng.flat.device:
            self.m = torch.zeros_like(eng.flat)
            self.v = torch.zeros_like(eng.flat)
            self.ema = eng.flat.clone() if self.ema_decay is not None else None
        return eng

    def step(self, grad_scale: float = 1.0):
        """Apply one update using the gradient arena filled by the last backward."""
        eng = self._ensure()
        self.step_count += 1
        _abi.check(_abi.lib().dmu_adam_ema(
            eng.flat.data_ptr(), eng.gflat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
            self.ema.data_ptr() if self.ema is not None else None, eng.flat.numel(),
            self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
            self.ema_decay if self.ema_decay is not None else 0.0, grad_scale, ops._stream()), "adam_ema")
        ops.LAUNCHES += 1

    def ema_state_dict(self, prefix=""):
        """EMA weights under the model's own parameter names."""
        eng = self._ensure()
        out = {}
        for k, (o, n) in eng.offs.items():
            out[prefix + k] = self.ema[o:o + n].view(eng.named[k].shape).clone()
        return out
