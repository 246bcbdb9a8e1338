"""Loss arithmetic of the reference, restated functionally.  TEST INFRASTRUCTURE.

Citations: upstream ``utils/losses.py``.
"""

import torch
import torch.nn.functional as F


def loss_terms(cfg: dict, loss_type: str = "mse"):
    """Resolve (w_mse, w_l1, w_huber, delta) the way utils/losses.py:42-56,
    105-131 does.  ``use_hybrid`` selects hybrid_weights and skips components
    whose weight is not > 0 (losses.py:120-129)."""
    cfg = cfg or {}
    delta = cfg.get("huber_delta", 1.0)
    if cfg.get("use_hybrid", False):
        w = cfg.get("hybrid_weights", {})
        wm, wl, wh = w.get("mse", 1.0), w.get("l1", 0.0), w.get("huber", 0.0)
        return (wm if wm > 0 else 0.0, wl if wl > 0 else 0.0, wh if wh > 0 else 0.0, delta)
    lt = loss_type.lower()
    if lt == "mse":
        return (cfg.get("mse_weight", 1.0), 0.0, 0.0, delta)
    if lt == "l1":
        return (0.0, cfg.get("l1_weight", 0.0), 0.0, delta)
    if lt == "huber":
        return (0.0, 0.0, cfg.get("huber_weight", 0.0), delta)
    raise ValueError(f"Unsupported single loss type: {lt}")


def time_weights(timesteps: torch.Tensor, kind: str = "snr", min_weight: float = 0.1, max_weight: float = 1.0):
    """utils/losses.py:133-181 — per-sample weights [B] (min-max normalised
    over the batch; the 'snr' branch rebuilds a linspace(1e-4, 2e-2, t_max+1)
    schedule every call, losses.py:144-160)."""
    if kind == "snr":
        betas = torch.linspace(1e-4, 2e-2, timesteps.max().item() + 1, device=timesteps.device)
        acp = torch.cumprod(1 - betas, dim=0).index_select(0, timesteps)
        snr = acp / (1 - acp)
        w = (snr / snr.max()).clamp(min=1e-5)
    elif kind == "linear":
        w = 1 - (timesteps.float() / timesteps.max())
    elif kind == "inverse":
        w = 1 / (timesteps.float() + 1)
    else:
        w = torch.ones_like(timesteps, dtype=torch.float)
    return min_weight + (max_weight - min_weight) * ((w - w.min()) / (w.max() - w.min() + 1e-5))


def diffusion_loss(pred, target, timesteps=None, loss_type: str = "mse", cfg: dict = None):
    """utils/losses.py:74-103 (perceptual term excluded: weight 0.0 in every
    shipped config and it needs a VGG download)."""
    cfg = cfg or {}
    wm, wl, wh, delta = loss_terms(cfg, loss_type)
    if cfg.get("use_hybrid", False):
        base = torch.zeros_like(pred)
        if wm > 0:
            base = base + wm * F.mse_loss(pred, target, reduction="none")
        if wl > 0:
            base = base + wl * F.l1_loss(pred, target, reduction="none")
        if wh > 0:
            base = base + wh * F.smooth_l1_loss(pred, target, reduction="none", beta=delta)
    else:
        lt = loss_type.lower()
        if lt == "mse":
            base = wm * F.mse_loss(pred, target, reduction="none")
        elif lt == "l1":
            base = wl * F.l1_loss(pred, target, reduction="none")
        else:
            base = wh * F.smooth_l1_loss(pred, target, reduction="none", beta=delta)
    if cfg.get("use_time_weighting", True) and timesteps is not None:
        p = cfg.get("time_weight_params", {"min_weight": 0.1, "max_weight": 1.0})
        w = time_weights(timesteps, cfg.get("time_weight_type", "snr"), p["min_weight"], p["max_weight"])
        base = base * w.view(-1, 1, 1, 1)
    return base.mean()


def score_matching_loss(score, fresh_noise, sigma):
    """utils/losses.py:238-242 — target uses a FRESH noise draw, not the noise
    that perturbed x (the caller passes the tensor randn_like would return)."""
    target = -fresh_noise / sigma.view(-1, 1, 1, 1)
    return F.mse_loss(score, target)


def energy_loss(energy_fn, x_real, x_fake, alpha, reg_weight: float):
    """utils/losses.py:264-285 — CD + gradient penalty; ``alpha`` [B,1,1,1] is
    the tensor torch.rand would return at losses.py:272.  The norm is over
    dim=1 (channels) only."""
    cd = torch.mean(energy_fn(x_real)) - torch.mean(energy_fn(x_fake))
    inter = (alpha * x_real + (1 - alpha) * x_fake).requires_grad_(True)
    e = energy_fn(inter)
    g = torch.autograd.grad(e, inter, grad_outputs=torch.ones_like(e), create_graph=True, retain_graph=True)[0]
    gp = ((g.norm(2, dim=1) - 1) ** 2).mean()
    return cd + reg_weight * gp
