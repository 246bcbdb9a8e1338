"""Functional fp32 restatement of the reference networks.  TEST INFRASTRUCTURE.

All functions take a ``state_dict``-style mapping (same keys as the reference,
see oracle/weights.py) and run on whatever device the tensors live on using
plain ATen ops.  File:line citations are relative to the upstream repository.
"""

import math
import torch
import torch.nn.functional as F

from .weights import gn_groups, down_plan, up_plan


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """models/layers/embeddings.py:24-39 — sin first, denominator half-1."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -k)
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def time_embedding(sd, pfx: str, t: torch.Tensor, base_dim: int) -> torch.Tensor:
    """models/layers/embeddings.py:52-58,66-75 — Linear, exact GELU, Linear."""
    e = sinusoidal_embedding(t, base_dim)
    e = F.linear(e, sd[pfx + "1.weight"], sd[pfx + "1.bias"])
    e = F.gelu(e)
    return F.linear(e, sd[pfx + "3.weight"], sd[pfx + "3.bias"])


def residual_block(sd, p: str, x: torch.Tensor, temb: torch.Tensor) -> torch.Tensor:
    """models/layers/residual.py:54-68."""
    cin = x.shape[1]
    cout = sd[p + "conv1.weight"].shape[0]
    h = F.group_norm(x, gn_groups(cin), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = h + F.linear(temb, sd[p + "time_mlp.weight"], sd[p + "time_mlp.bias"])[..., None, None]
    h = F.group_norm(h, gn_groups(cout), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "shortcut.weight"], sd[p + "shortcut.bias"])
    return h + x


def attention_block(sd, p: str, x: torch.Tensor, heads: int = 4) -> torch.Tensor:
    """models/layers/attention.py:36-68 — no pre-norm, post GroupNorm(proj + x)."""
    b, c, hh, ww = x.shape
    tok = x.reshape(b, c, hh * ww).transpose(1, 2)
    d = c // heads

    def split(z):
        return z.reshape(b, hh * ww, heads, d).transpose(1, 2)

    q = split(F.linear(tok, sd[p + "query_projection.weight"], sd[p + "query_projection.bias"]))
    k = split(F.linear(tok, sd[p + "key_projection.weight"], sd[p + "key_projection.bias"]))
    v = split(F.linear(tok, sd[p + "value_projection.weight"], sd[p + "value_projection.bias"]))
    att = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = torch.matmul(att, v).permute(0, 2, 1, 3).reshape(b, hh * ww, c)
    o = F.linear(o, sd[p + "final_projection.weight"], sd[p + "final_projection.bias"])
    o = o.transpose(-1, -2).reshape(b, c, hh, ww)
    return F.group_norm(o + x, gn_groups(c), sd[p + "norm.weight"], sd[p + "norm.bias"], eps=1e-5)


def _stage(sd, p, kind, x, temb):
    for j in range(2):
        x = residual_block(sd, f"{p}res_blocks.{j}.", x, temb)
        if kind == "attn":
            x = attention_block(sd, f"{p}attention_blocks.{j}.", x)
    return x


def unet_body(sd, x: torch.Tensor, temb: torch.Tensor, prefix: str = "model.") -> torch.Tensor:
    """models/ddpm.py:106-135 given a ready time embedding [B, 4C]."""
    C = sd[prefix + "initial_conv.weight"].shape[0]
    h = F.conv2d(x, sd[prefix + "initial_conv.weight"], sd[prefix + "initial_conv.bias"], padding=1)
    skips = [h]
    for i, (kind, _, _) in enumerate(down_plan(C)):
        p = f"{prefix}down_blocks.{i}."
        h = _stage(sd, p, kind, h, temb)
        h = F.conv2d(h, sd[p + "downsample.weight"], sd[p + "downsample.bias"], stride=2, padding=1)
        skips.append(h)
    h = residual_block(sd, prefix + "bottleneck.0.", h, temb)
    h = attention_block(sd, prefix + "bottleneck.1.", h)
    h = residual_block(sd, prefix + "bottleneck.2.", h, temb)
    # ddpm.py:126-130: zip() stops after five blocks, so skips[0] (stem output) is never used.
    for i, ((kind, _, _), skip) in enumerate(zip(up_plan(C), reversed(skips))):
        p = f"{prefix}up_blocks.{i}."
        h = _stage(sd, p, kind, torch.cat([h, skip], dim=1), temb)
        h = F.conv_transpose2d(h, sd[p + "upsample.weight"], sd[p + "upsample.bias"], stride=2, padding=1)
    h = F.group_norm(h, 32, sd[prefix + "output_conv.0.weight"], sd[prefix + "output_conv.0.bias"], eps=1e-5)
    return F.conv2d(F.silu(h), sd[prefix + "output_conv.2.weight"], sd[prefix + "output_conv.2.bias"], padding=1)


def unet_forward(sd, x: torch.Tensor, t: torch.Tensor, prefix: str = "model.") -> torch.Tensor:
    """models/ddpm.py:93-135 — eps_theta(x, t)."""
    C = sd[prefix + "initial_conv.weight"].shape[0]
    temb = time_embedding(sd, prefix + "time_embedding.positional_encoding.", t, C)
    return unet_body(sd, x, temb, prefix)


def scorenet_forward(sd, x: torch.Tensor, sigma: torch.Tensor, prefix: str = "model.") -> torch.Tensor:
    """Repaired ``ScoreNet.forward``: UNet body conditioned on
    ``time_embed(log sigma)`` (models/score_based.py:57-61,82-83); the shipped
    body (score_based.py:84-99) references layers that do not exist."""
    e = torch.log(sigma).view(-1, 1)
    e = F.linear(e, sd[prefix + "time_embed.0.weight"], sd[prefix + "time_embed.0.bias"])
    e = F.linear(F.silu(e), sd[prefix + "time_embed.2.weight"], sd[prefix + "time_embed.2.bias"])
    return unet_body(sd, x, e, prefix)


def energynet_forward(sd, x: torch.Tensor, prefix: str = "model.") -> torch.Tensor:
    """models/energy_based.py:62-85 — three convs, GroupNorm(8), SiLU, mean, dense -> [B]."""
    h = F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, 8, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], eps=1e-5))
    h = F.conv2d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"], padding=1)
    h = F.silu(F.group_norm(h, 8, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], eps=1e-5))
    h = F.silu(F.conv2d(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"], padding=1))
    h = h.mean(dim=[2, 3])
    return F.linear(h, sd[prefix + "dense.weight"], sd[prefix + "dense.bias"]).squeeze(-1)
