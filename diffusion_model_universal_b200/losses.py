"""DiffusionLoss — host mirror of utils/losses.py:8-181 over one fused kernel.

The [B]-sized time weights are computed with the reference's own torch
expressions (bit-identical on the same device); the
per-element work — difference, mse/l1/huber mix, weighting, mean and the
gradient w.r.t. the prediction — is a single pass of ``dmu_diffusion_loss``.
"""

from typing import Dict, Optional

import torch

from . import ops


class DiffusionLoss:
    LOSS_TYPES = ("mse", "l1", "huber", "hybrid")

    def __init__(self, loss_type: str = "mse", loss_config: Optional[Dict] = None):
        self.loss_type = loss_type.lower()
        self.loss_config = loss_config or {}
        if self.loss_type not in self.LOSS_TYPES:
            raise ValueError(f"Unsupported loss type: {loss_type}")
        c = self.loss_config
        self.mse_weight = c.get("mse_weight", 1.0)
        self.l1_weight = c.get("l1_weight", 0.0)
        self.huber_weight = c.get("huber_weight", 0.0)
        self.huber_delta = c.get("huber_delta", 1.0)
        self.use_hybrid = c.get("use_hybrid", False)
        if self.use_hybrid:
            w = c.get("hybrid_weights", {})
            self.hybrid_weights = {"mse": w.get("mse", 1.0), "l1": w.get("l1", 0.0), "huber": w.get("huber", 0.0)}
        self.use_time_weighting = c.get("use_time_weighting", True)
        self.time_weight_type = c.get("time_weight_type", "snr")
        self.time_weight_params = c.get("time_weight_params", {"min_weight": 0.1, "max_weight": 1.0})
        self.perceptual_weight = c.get("perceptual_weight", 0.0)
        self.adversarial_weight = c.get("adversarial_weight", 0.0)
        if self.perceptual_weight > 0:
            raise NotImplementedError("perceptual (VGG16) term is outside the hot path (weight 0.0 in every shipped config)")

    def coefficients(self):
        """(w_mse, w_l1, w_huber) the way losses.py:105-131 resolves them."""
        if self.use_hybrid:
            hw = self.hybrid_weights
            return tuple(float(v) if v > 0 else 0.0 for v in (hw["mse"], hw["l1"], hw["huber"]))
        if self.loss_type == "mse":
            return (float(self.mse_weight), 0.0, 0.0)
        if self.loss_type == "l1":
            return (0.0, float(self.l1_weight), 0.0)
        if self.loss_type == "huber":
            return (0.0, 0.0, float(self.huber_weight))
        raise ValueError(f"Unsupported single loss type: {self.loss_type}")  # 'hybrid' without use_hybrid, losses.py:113-114

    max_t = None   # upper bound on the timesteps (set by the owning model): enables the host-sync-free 'snr' weights
    _snr_tables = None     # device -> fp32 [max_t, max_t]: row tm = cumprod(1 - linspace(1e-4, 2e-2, tm + 1)) (dmu_snr_time_weights)

    def _snr_table(self, device) -> torch.Tensor:
        """Every table the reference could build at losses.py:150-157, one per possible t_max, made once per device with the
        reference's own calls (so each row is bit-identical to what it computes there).  4 MB at 1000 timesteps."""
        if self._snr_tables is None:
            self._snr_tables = {}
        key = (str(device), int(self.max_t))
        tab = self._snr_tables.get(key)
        if tab is None:
            T = int(self.max_t)
            tab = torch.zeros(T, T, device=device, dtype=torch.float32)
            for tm in range(T):
                tab[tm, :tm + 1] = torch.cumprod(1 - torch.linspace(1e-4, 2e-2, tm + 1, device=device), dim=0)
            self._snr_tables[key] = tab
        return tab

    def _snr_alphas_cumprod(self, timesteps: torch.Tensor) -> torch.Tensor:
        """losses.py:150-157: ``cumprod(1 - linspace(1e-4, 2e-2, t_max + 1))[t]``.  The reference reads ``t_max`` back to the
        host (``timesteps.max().item()``) to size the table; when an upper bound is known the same table is built for the
        bound with ``t_max`` kept on the device (linspace's own two-sided formula), so the training loop never syncs."""
        dev = timesteps.device
        if self.max_t is None:
            betas = torch.linspace(1e-4, 2e-2, timesteps.max().item() + 1, device=dev)
            return torch.cumprod(1 - betas, dim=0).index_select(0, timesteps)
        tmax = timesteps.max()
        i = torch.arange(int(self.max_t), device=dev)
        start = torch.full((), 1e-4, device=dev, dtype=torch.float32)     # fill kernels, not host copies: capturable in a CUDA graph
        end = torch.full((), 2e-2, device=dev, dtype=torch.float32)
        step = (end - start) / tmax.clamp(min=1).to(torch.float32)
        steps = tmax + 1
        lo = start + step * i.to(torch.float32)
        hi = end - step * (steps - 1 - i).to(torch.float32)
        betas = torch.where(i < steps // 2, lo, hi)
        betas = torch.where(steps == 1, start, betas)        # linspace(a, b, 1) == [a]
        return torch.cumprod(1 - betas, dim=0).index_select(0, timesteps)

    def time_weights(self, timesteps: torch.Tensor) -> Optional[torch.Tensor]:
        """The [B] weights this loss would apply for ``timesteps`` (None = unweighted)."""
        if not self.use_time_weighting or timesteps is None:
            return None
        if (self.time_weight_type == "snr" and self.max_t is not None and timesteps.is_cuda and timesteps.dtype == torch.int64
                and timesteps.dim() == 1 and timesteps.is_contiguous()):
            # one launch, no host sync (the torch-op formulation below stays as the general path and as the test's comparator)
            return ops.snr_time_weights(timesteps, self._snr_table(timesteps.device), self.time_weight_params["min_weight"],
                                        self.time_weight_params["max_weight"])
        return self._get_time_weights(timesteps).float().contiguous()

    def _get_time_weights(self, timesteps: torch.Tensor) -> torch.Tensor:
        """losses.py:133-181, returned flat [B]."""
        lo, hi = self.time_weight_params["min_weight"], self.time_weight_params["max_weight"]
        if self.time_weight_type == "snr":
            acp = self._snr_alphas_cumprod(timesteps)
            snr = acp / (1 - acp)
            w = (snr / snr.max()).clamp(min=1e-5)
        elif self.time_weight_type == "linear":
            w = 1 - (timesteps.float() / timesteps.max())
        elif self.time_weight_type == "inverse":
            w = 1 / (timesteps.float() + 1)
        else:
            w = torch.ones_like(timesteps, dtype=torch.float)
        return lo + (hi - lo) * ((w - w.min()) / (w.max() - w.min() + 1e-5))

    def __call__(self, pred: torch.Tensor, target: torch.Tensor, timesteps: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.time_weights(timesteps)
        wm, wl, wh = self.coefficients()
        return _LossFn.apply(pred, target, w, wm, wl, wh, float(self.huber_delta))


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, w, wm, wl, wh, delta):
        need = pred.requires_grad
        loss, dpred = ops.diffusion_loss(pred.contiguous().float(), target.contiguous().float(), w, wm, wl, wh, delta, need)
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.dpred
        if d is None:
            return (None,) * 7
        # g is the scalar upstream gradient (1.0 for loss.backward()); keep it on device, no sync
        return (d * g, None, None, None, None, None, None)
