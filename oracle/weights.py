"""State-dict contract of the reference networks + deterministic test weights.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The key order and shapes restate what ``models/ddpm.py:45-91`` (UNet),
``models/layers/residual.py:20-52,81-91,111-121,154-181,218-245`` (blocks),
``models/layers/attention.py:12-27`` (attention) and
``models/layers/embeddings.py:52-58`` (time embedding) register.  The golden
generator loads these tensors into the live reference with ``strict=True``,
so any drift in names or shapes fails at fixture-generation time.
"""

from collections import OrderedDict
import math
import torch


def gn_groups(channels: int, num_groups: int = 32) -> int:
    """Group count rule of ``models/layers/residual.py:22-29``."""
    g = min(num_groups, channels)
    while channels % g != 0 and g > 1:
        g -= 1
    return g


def _res(spec, p, cin, cout, temb):
    # residual.py:31-46 registration order: norm1, conv1, time_mlp, norm2, conv2, shortcut
    spec[p + "norm1.weight"] = (cin,)
    spec[p + "norm1.bias"] = (cin,)
    spec[p + "conv1.weight"] = (cout, cin, 3, 3)
    spec[p + "conv1.bias"] = (cout,)
    spec[p + "time_mlp.weight"] = (cout, temb)
    spec[p + "time_mlp.bias"] = (cout,)
    spec[p + "norm2.weight"] = (cout,)
    spec[p + "norm2.bias"] = (cout,)
    spec[p + "conv2.weight"] = (cout, cout, 3, 3)
    spec[p + "conv2.bias"] = (cout,)
    if cin != cout:
        spec[p + "shortcut.weight"] = (cout, cin, 1, 1)
        spec[p + "shortcut.bias"] = (cout,)


def _attn(spec, p, c):
    # attention.py:21-27
    for n in ("query_projection", "key_projection", "value_projection", "final_projection"):
        spec[p + n + ".weight"] = (c, c)
        spec[p + n + ".bias"] = (c,)
    spec[p + "norm.weight"] = (c,)
    spec[p + "norm.bias"] = (c,)


# (kind, cin, cout) per block; kind: 'conv' | 'attn'   (ddpm.py:56-84)
def down_plan(C):
    return [("conv", C, C), ("conv", C, C), ("conv", C, 2 * C), ("attn", 2 * C, 2 * C), ("conv", 2 * C, 4 * C)]


def up_plan(C):
    return [("conv", 8 * C, 4 * C), ("attn", 6 * C, 2 * C), ("conv", 4 * C, 2 * C), ("conv", 3 * C, C), ("conv", 2 * C, C)]


def unet_param_spec(model_channels: int = 64, in_channels: int = 3, prefix: str = "model.") -> "OrderedDict[str, tuple]":
    """Ordered name -> shape map of ``UNet.state_dict()`` (ddpm.py:45-91)."""
    C, T = model_channels, 4 * model_channels
    s = OrderedDict()
    s[prefix + "initial_conv.weight"] = (C, in_channels, 3, 3)
    s[prefix + "initial_conv.bias"] = (C,)
    te = prefix + "time_embedding.positional_encoding."
    s[te + "1.weight"] = (T, C)
    s[te + "1.bias"] = (T,)
    s[te + "3.weight"] = (T, T)
    s[te + "3.bias"] = (T,)
    for i, (kind, cin, cout) in enumerate(down_plan(C)):
        p = f"{prefix}down_blocks.{i}."
        for j in range(2):
            _res(s, f"{p}res_blocks.{j}.", cin if j == 0 else cout, cout, T)
        if kind == "attn":
            for j in range(2):
                _attn(s, f"{p}attention_blocks.{j}.", cout)
        s[p + "downsample.weight"] = (cout, cout, 4, 4)
        s[p + "downsample.bias"] = (cout,)
    _res(s, prefix + "bottleneck.0.", 4 * C, 4 * C, T)
    _attn(s, prefix + "bottleneck.1.", 4 * C)
    _res(s, prefix + "bottleneck.2.", 4 * C, 4 * C, T)
    for i, (kind, cin, cout) in enumerate(up_plan(C)):
        p = f"{prefix}up_blocks.{i}."
        for j in range(2):
            _res(s, f"{p}res_blocks.{j}.", cin if j == 0 else cout, cout, T)
        if kind == "attn":
            for j in range(2):
                _attn(s, f"{p}attention_blocks.{j}.", cout)
        s[p + "upsample.weight"] = (cout, cout, 4, 4)  # ConvTranspose2d: [in, out, kh, kw]
        s[p + "upsample.bias"] = (cout,)
    s[prefix + "output_conv.0.weight"] = (C,)
    s[prefix + "output_conv.0.bias"] = (C,)
    s[prefix + "output_conv.2.weight"] = (in_channels, C, 3, 3)
    s[prefix + "output_conv.2.bias"] = (in_channels,)
    return s


def scorenet_param_spec(model_channels: int = 64, in_channels: int = 3, prefix: str = "model.") -> "OrderedDict[str, tuple]":
    """UNet spec + the sigma embedding MLP of ``models/score_based.py:57-61``."""
    s = unet_param_spec(model_channels, in_channels, prefix)
    C = model_channels
    s[prefix + "time_embed.0.weight"] = (C, 1)
    s[prefix + "time_embed.0.bias"] = (C,)
    s[prefix + "time_embed.2.weight"] = (4 * C, C)
    s[prefix + "time_embed.2.bias"] = (4 * C,)
    return s


def energynet_param_spec(model_channels: int = 64, in_channels: int = 3, prefix: str = "model.") -> "OrderedDict[str, tuple]":
    """``EnergyNet`` registration order, ``models/energy_based.py:51-60``."""
    C = model_channels
    s = OrderedDict()
    s[prefix + "conv1.weight"] = (C, in_channels, 3, 3)
    s[prefix + "conv1.bias"] = (C,)
    s[prefix + "conv2.weight"] = (2 * C, C, 3, 3)
    s[prefix + "conv2.bias"] = (2 * C,)
    s[prefix + "conv3.weight"] = (4 * C, 2 * C, 3, 3)
    s[prefix + "conv3.bias"] = (4 * C,)
    s[prefix + "norm1.weight"] = (C,)
    s[prefix + "norm1.bias"] = (C,)
    s[prefix + "norm2.weight"] = (2 * C,)
    s[prefix + "norm2.bias"] = (2 * C,)
    s[prefix + "dense.weight"] = (1, 4 * C)
    s[prefix + "dense.bias"] = (1,)
    return s


def make_state_dict(spec, seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, fully non-zero test weights for a param spec.

    The reference zero-initialises ``conv2`` and ``time_mlp`` of every
    ResidualBlock (residual.py:49-52), which makes a fresh network blind to
    ``t`` (SURVEY.md §4 pitfall), so parity weights re-randomise everything:
    matrices/filters ~ N(0, 1/fan_in) (keeps activations O(1) through 60+
    layers), norm scales ~ 1 + 0.1 N(0,1), biases ~ 0.05 N(0,1).  Uses a CPU
    torch.Generator so the same seed yields the same tensors in the build
    container and on the GPU box.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    for name, shape in spec.items():
        if len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if "upsample.weight" in name:  # ConvTranspose2d [in,out,k,k]: each output sums in*k*k/4 taps
                fan_in = shape[0] * 4
            w = torch.randn(shape, generator=g, dtype=torch.float32) / math.sqrt(fan_in)
        elif name.endswith("weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        else:
            w = 0.05 * torch.randn(shape, generator=g, dtype=torch.float32)
        sd[name] = w.to(dtype)
    return sd
