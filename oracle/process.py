"""Diffusion-process arithmetic of the reference, restated.  TEST INFRASTRUCTURE.

Schedules, q_sample, the DDPM ancestral step, DDIM tables/step, the score and
energy Langevin updates.  Noise is always an explicit argument so the CUDA
path can be fed the same tensors.  Citations: upstream file:line.
"""

import math
import numpy as np
import torch


# ---------------------------------------------------------------- DDPM
def linear_schedule(beta_start: float, beta_end: float, T: int, device="cpu"):
    """models/ddpm.py:176-178 — betas, alphas, alphas_cumprod (fp32[T]).

    The reference builds them on CPU at construction and moves them with the
    module, so they are always computed on CPU here too."""
    betas = torch.linspace(beta_start, beta_end, T)
    alphas = 1 - betas
    acp = torch.cumprod(alphas, dim=0)
    return betas.to(device), alphas.to(device), acp.to(device)


def q_sample(x0, t, noise, acp):
    """models/ddpm.py:295-296."""
    a = acp[t][:, None, None, None]
    return torch.sqrt(a) * x0 + torch.sqrt(1 - a) * noise


def ddpm_reverse_step(x, eps, t, noise, betas, alphas, acp):
    """models/ddpm.py:306-329 with eps = model(x, t) injected.

    ``noise`` is the tensor the reference would draw with randn_like at
    ddpm.py:324; it is ignored when t[0] == 0 (ddpm.py:323-327)."""
    alpha_t = alphas[t][:, None, None, None]
    acp_t = acp[t][:, None, None, None]
    acp_prev = acp[t - 1][:, None, None, None] if t[0] > 0 else torch.ones_like(acp_t)
    beta_t = betas[t][:, None, None, None]
    beta_tilde = (1 - acp_prev) / (1 - acp_t) * beta_t
    mean = torch.pow(alpha_t, -0.5) * (x - beta_t / torch.sqrt(1 - acp_t) * eps)
    if t[0] > 0:
        return mean + torch.sqrt(beta_tilde) * noise
    return mean


# ---------------------------------------------------------------- DDIM
def ddim_timesteps(T: int, S: int, method: str = "uniform") -> torch.Tensor:
    """models/ddim.py:55-63."""
    if method == "uniform":
        return torch.arange(0, T, T // S)
    if method == "quad":
        ts = torch.linspace(0, torch.sqrt(torch.tensor(T * .8)), S) ** 2
        return ts.long()
    raise NotImplementedError(f"Unknown discretization method: {method}")


def ddim_tables(acp: torch.Tensor, steps: torch.Tensor, eta: float):
    """models/ddim.py:70-81 — (alphas, alphas_prev, sigmas, sqrt_one_minus_alphas).

    alphas_prev[0] is alphas_cumprod[0], not 1.0 (ddim.py:71)."""
    a = acp[steps]
    a_prev = torch.cat([acp[0:1], acp[steps[:-1]]])
    sig = eta * torch.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return a, a_prev, sig, torch.sqrt(1. - a)


def ddim_step(x, eps, idx, tables, eta: float, noise=None):
    """models/ddim.py:97-124 with ``idx`` indexing the S-entry tables (the
    repaired driver of SURVEY.md §3.3) and eps injected (``pred_noise``)."""
    a, a_prev, sig, s1m = tables
    a_t, ap_t, s_t, r_t = a[idx], a_prev[idx], sig[idx], s1m[idx]
    x0 = (x - r_t[:, None, None, None] * eps) / torch.sqrt(a_t)[:, None, None, None]
    x0 = x0.clamp(-1, 1)
    dir_xt = torch.sqrt(1. - ap_t - s_t ** 2)[:, None, None, None] * eps
    if eta > 0:
        z = noise.clamp(-3, 3)
    else:
        z = 0
    return torch.sqrt(ap_t)[:, None, None, None] * x0 + dir_xt + s_t[:, None, None, None] * z


# ---------------------------------------------------------------- score
def score_sigma_from_u(u, sigma_min: float, sigma_max: float):
    """models/score_based.py:197."""
    return sigma_min * (sigma_max / sigma_min) ** u


def score_sigma_ladder(sigma_min: float, sigma_max: float, n: int, device="cpu"):
    """models/score_based.py:228-231."""
    return torch.exp(torch.linspace(np.log(sigma_max), np.log(sigma_min), n, device=device))


def score_langevin_step(x, score, noise, sigma, beta: float):
    """models/score_based.py:236-245: step = 2(sigma*beta)^2,
    x + step*score + sqrt(2*step)*noise (sigma is a 0-dim tensor)."""
    step = (sigma * beta) ** 2 * 2
    return x + step * score + torch.sqrt(step * 2) * noise


# ---------------------------------------------------------------- energy
def energy_langevin_step(x, grad, noise, step_size: float):
    """models/energy_based.py:271-273 with ``math.sqrt`` for the float step
    (the shipped ``torch.sqrt(float)`` raises TypeError)."""
    return x - step_size * grad + math.sqrt(2 * step_size) * noise


def energy_renoise(x, noise, acp, t: int):
    """models/energy_based.py:240-246 (t > 0)."""
    a_next = acp[t - 1]
    a = acp[t]
    sigma = torch.sqrt((1 - a_next) / (1 - a)) * torch.sqrt(1 - a / a_next)
    return torch.sqrt(a_next / a) * x + sigma * noise
