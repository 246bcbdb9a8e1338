"""Energy-based model behind the reference API (models/energy_based.py) on the CUDA library.

The reference class cannot run as shipped: ``generate_samples`` is abstract, ``forward`` passes ``t`` to a
network that takes only ``x`` (energy_based.py:177), time conditioning widens conv1 by channels nothing produces
(:136-138) and ``torch.sqrt(float)`` raises (:273).  This class implements the repaired behaviour the oracle pins:
``forward(x, t) = EnergyNet(x)``, ``use_time_conditioning`` must be False, ``math.sqrt`` for the step size.

EnergyNet (energy_based.py:51-85) is NOT a UNet: conv3x3(in->C) -> GN(8) -> SiLU -> conv3x3(C->2C) -> GN(8) -> SiLU ->
conv3x3(2C->4C) -> SiLU -> global mean -> Linear(4C, 1).  Forward and the first-order backward (input gradient for the
Langevin dynamics, parameter gradients of the contrastive-divergence terms) run on the same kernels as the UNet.

NOT built in this round: the parameter gradient of the gradient-penalty term of ``EnergyBasedLoss``
(utils/losses.py:277-285, ``create_graph=True``), which needs second-order backward kernels (SiLU'' and the GroupNorm
double backward).  The loss VALUE includes the penalty exactly; ``backward()`` of a loss with ``regularization_weight > 0``
raises unless the config opts into ``gradient_penalty_grad: "skip"`` (penalty treated as a constant for the gradient).
"""

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _abi, ops
from ._abi import F32, BF16, Tensor4, ConvParams, WgradParams, GnParams, RepackDesc
from .base_model import BaseDiffusion
from .losses import DiffusionLoss


def _null():
    return Tensor4(None, 0, 0, 0, 0, 0, 0)


class EnergyNet(nn.Module):
    """energy_based.py:42-60: same parameter names, shapes and default initialisation."""

    def __init__(self, in_channels: int, model_channels: int, precision: str = "fp32"):
        super().__init__()
        Cm = model_channels
        self.conv1 = nn.Conv2d(in_channels, Cm, 3, padding=1)
        self.conv2 = nn.Conv2d(Cm, Cm * 2, 3, padding=1)
        self.conv3 = nn.Conv2d(Cm * 2, Cm * 4, 3, padding=1)
        self.norm1 = nn.GroupNorm(8, Cm)
        self.norm2 = nn.GroupNorm(8, Cm * 2)
        self.dense = nn.Linear(Cm * 4, 1)
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
        self.precision = precision
        self.in_channels, self.model_channels = in_channels, Cm

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """E(x): fp32 [B,C,H,W] -> fp32 [B]."""
        ops._need_cuda(x)
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected x of shape [B,{self.in_channels},H,W], got {tuple(x.shape)}")
        params = [p for _, p in self.named_parameters()]
        return _EnergyFn.apply(self, x.contiguous().float(), *params)

    # ------------------------------------------------------------------ launches
    def _run(self, x: torch.Tensor, dE: Optional[torch.Tensor], want_dx: bool, want_dw: bool):
        """Forward (and, when dE is given, backward) of the energy network.
        Returns (E [B], dx [B,C,H,W] or None, {param name: grad} or None)."""
        lib = _abi.lib()
        st = ops._stream()
        dev = x.device
        code = BF16 if self.precision == "bf16" else F32
        tdt = torch.bfloat16 if code == BF16 else torch.float32
        N, Ci, H, W = x.shape
        Cm = self.model_channels
        chans = [Ci, Cm, 2 * Cm, 4 * Cm]
        keep = []

        def chk(rc, what):
            ops._launched()
            _abi.check(rc, what)

        # ---- filters in the compute dtype: [O][R][S][I] for fprop, [I][R][S][O] for dgrad
        convs = [self.conv1, self.conv2, self.conv3]
        wf, wb, descs = [], [], []
        for i, cv in enumerate(convs):
            O, I = chans[i + 1], chans[i]
            f = torch.empty(O * 9 * I, device=dev, dtype=tdt)
            b = torch.empty(O * 9 * I, device=dev, dtype=tdt)
            wsrc = cv.weight.detach().contiguous()
            keep.append(wsrc)
            descs.append(RepackDesc(wsrc.data_ptr(), f.data_ptr(), O, I, 3, 3, 0, code))
            descs.append(RepackDesc(wsrc.data_ptr(), b.data_ptr(), I, O, 3, 3, 1, code))
            wf.append(f)
            wb.append(b)
        arr = (RepackDesc * len(descs))(*descs)
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        chk(lib.dmu_repack_weights(table.data_ptr(), len(descs), max(c.weight.numel() for c in convs), st), "repack_weights")

        def conv(xt4, yt4, w, Ck, Cj, bias, gather, w_strides):
            p = ConvParams(xt4, yt4, _null(), w.data_ptr(), w_strides[0], w_strides[1], w_strides[2], bias, None, 0,
                           N, H, W, Ck, H, W, Cj, 3, 3, 1, 1, gather, code, 0, 0, None, 0)
            chk(lib.dmu_conv2d(C.byref(p), st), "conv2d")

        def gn(xh, yh, sums, norm, silu=1):
            p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(yh), _null(), _null(), _null(), sums.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(),
                         None, None, None, N, H, W, xh.shape[3], 8, silu, 1e-5, 0)
            chk(lib.dmu_gn_forward(C.byref(p), st), "gn_forward")

        # ---- forward
        h1 = torch.empty(N, H, W, chans[1], device=dev, dtype=tdt)
        conv(ops.t4_nchw(x), ops.t4_nhwc(h1), wf[0], Ci, chans[1], self.conv1.bias.data_ptr(), 0, (9 * Ci, 1, Ci))
        s1 = torch.zeros(N, 8, 2, device=dev)
        a1 = torch.empty_like(h1)
        gn(h1, a1, s1, self.norm1)
        h2 = torch.empty(N, H, W, chans[2], device=dev, dtype=tdt)
        conv(ops.t4_nhwc(a1), ops.t4_nhwc(h2), wf[1], chans[1], chans[2], self.conv2.bias.data_ptr(), 0, (9 * chans[1], 1, chans[1]))
        s2 = torch.zeros(N, 8, 2, device=dev)
        a2 = torch.empty_like(h2)
        gn(h2, a2, s2, self.norm2)
        h3 = torch.empty(N, H, W, chans[3], device=dev, dtype=tdt)
        conv(ops.t4_nhwc(a2), ops.t4_nhwc(h3), wf[2], chans[2], chans[3], self.conv3.bias.data_ptr(), 0, (9 * chans[2], 1, chans[2]))
        pooled = torch.zeros(N, chans[3], device=dev)
        t3 = ops.t4_nhwc(h3)
        chk(lib.dmu_silu_pool_fwd(C.byref(t3), N, H, W, chans[3], pooled.data_ptr(), chans[3], 1.0 / (H * W), st), "silu_pool_fwd")
        E = torch.empty(N, 1, device=dev)
        pl = ConvParams(ops.t4_rows(pooled), ops.t4_rows(E), _null(), self.dense.weight.data_ptr(), chans[3], 1, 0, self.dense.bias.data_ptr(), None, 0,
                        N, 1, 1, chans[3], 1, 1, 1, 1, 1, 1, 0, 0, F32, 0, 0, None, 0)
        chk(lib.dmu_conv2d(C.byref(pl), st), "dense")
        if dE is None:
            return E.view(N), None, None

        # ---- backward: dE [N] -> dx (and parameter gradients)
        grads = {}
        dE2 = dE.reshape(N, 1).contiguous().float()
        dpooled = torch.empty(N, chans[3], device=dev)
        pd = ConvParams(ops.t4_rows(dE2), ops.t4_rows(dpooled), _null(), self.dense.weight.data_ptr(), 1, chans[3], 0, None, None, 0,
                        N, 1, 1, 1, 1, 1, chans[3], 1, 1, 1, 0, 0, F32, 0, 0, None, 0)
        chk(lib.dmu_conv2d(C.byref(pd), st), "dense dgrad")

        def wgrad(p4, q4, Ca, Cb, name, cv, rows=False):
            dw = torch.zeros_like(cv.weight)
            db = torch.zeros_like(cv.bias)
            if rows:
                p = WgradParams(p4, q4, dw.data_ptr(), Cb, 1, 0, db.data_ptr(), N, 1, 1, Ca, 1, 1, Cb, 1, 1, 1, 0, 0)
            else:
                p = WgradParams(p4, q4, dw.data_ptr(), Cb * 9, 9, 1, db.data_ptr(), N, H, W, Ca, H, W, Cb, 3, 3, 1, 1, 0)
            chk(lib.dmu_conv2d_wgrad(C.byref(p), st), "wgrad " + name)
            grads[name + ".weight"], grads[name + ".bias"] = dw, db

        def gn_bwd(xh, dyh, dxh, sums, norm, name):
            red = torch.zeros(N, xh.shape[3], 2, device=dev)
            dg, db = (torch.zeros_like(norm.weight), torch.zeros_like(norm.bias)) if want_dw else (None, None)
            p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), ops.t4_nhwc(dxh), _null(), _null(), sums.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(),
                         red.data_ptr(), dg.data_ptr() if want_dw else None, db.data_ptr() if want_dw else None, N, H, W, xh.shape[3], 8, 1, 1e-5, 0)
            chk(lib.dmu_gn_backward(C.byref(p), st), "gn_backward")
            if want_dw:
                grads[name + ".weight"], grads[name + ".bias"] = dg, db

        if want_dw:
            wgrad(ops.t4_rows(dE2), ops.t4_rows(pooled), 1, chans[3], "dense", self.dense, rows=True)
        dh3 = torch.empty_like(h3)
        t3d = ops.t4_nhwc(dh3)
        chk(lib.dmu_silu_pool_bwd(C.byref(t3), C.byref(t3d), N, H, W, chans[3], dpooled.data_ptr(), chans[3], 1.0 / (H * W), st), "silu_pool_bwd")
        da2 = torch.empty_like(a2)
        conv(ops.t4_nhwc(dh3), ops.t4_nhwc(da2), wb[2], chans[3], chans[2], None, 1, (9 * chans[3], 1, chans[3]))
        if want_dw:
            wgrad(ops.t4_nhwc(dh3), ops.t4_nhwc(a2), chans[3], chans[2], "conv3", self.conv3)
        dh2 = torch.empty_like(h2)
        gn_bwd(h2, da2, dh2, s2, self.norm2, "norm2")
        da1 = torch.empty_like(a1)
        conv(ops.t4_nhwc(dh2), ops.t4_nhwc(da1), wb[1], chans[2], chans[1], None, 1, (9 * chans[2], 1, chans[2]))
        if want_dw:
            wgrad(ops.t4_nhwc(dh2), ops.t4_nhwc(a1), chans[2], chans[1], "conv2", self.conv2)
        dh1 = torch.empty_like(h1)
        gn_bwd(h1, da1, dh1, s1, self.norm1, "norm1")
        dx = None
        if want_dx:
            dx = torch.empty_like(x)
            conv(ops.t4_nhwc(dh1), ops.t4_nchw(dx), wb[0], chans[1], Ci, None, 1, (9 * chans[1], 1, chans[1]))
        if want_dw:
            wgrad(ops.t4_nhwc(dh1), ops.t4_nchw(x), chans[1], Ci, "conv1", self.conv1)
        return E.view(N), dx, grads if want_dw else None

    def energy_and_input_grad(self, x: torch.Tensor):
        """(E(x) [B], d sum(E) / dx) in one forward+backward pass: what a Langevin step needs (energy_based.py:266-268)."""
        ones = torch.ones(x.shape[0], device=x.device)
        E, dx, _ = self._run(x.contiguous().float(), ones, True, False)
        return E, dx


class _EnergyFn(torch.autograd.Function):
    """First-order autograd of EnergyNet (recomputes the forward inside backward: the net is tiny)."""

    @staticmethod
    def forward(ctx, net: EnergyNet, x, *params):
        E, _, _ = net._run(x, None, False, False)
        ctx.net = net
        ctx.save_for_backward(x)
        ctx.need_x = x.requires_grad
        return E

    @staticmethod
    def backward(ctx, dE):
        net = ctx.net
        (x,) = ctx.saved_tensors
        names = [n for n, _ in net.named_parameters()]
        need_w = any(p.requires_grad for _, p in net.named_parameters())
        _, dx, grads = net._run(x, dE.contiguous(), ctx.need_x, need_w)
        return (None, dx) + tuple(grads[n] if grads is not None else None for n in names)


class _PenaltyFn(torch.autograd.Function):
    """Value of the gradient penalty as a function of the network output it is attached to; its parameter gradient is the
    second-order term this round does not implement."""

    @staticmethod
    def forward(ctx, anchor, value, skip):
        ctx.skip = skip
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        if not ctx.skip:
            raise NotImplementedError(
                "EnergyBasedLoss gradient penalty (utils/losses.py:277-285): its parameter gradient needs second-order backward "
                "kernels that are not built yet.  Set regularization_weight to 0, or opt into treating the penalty as a constant "
                "for the gradient with model_config['gradient_penalty_grad'] = 'skip'.")
        return torch.zeros_like(g), None, None


class EnergyBasedDiffusion(BaseDiffusion):
    def __init__(self, config: Dict):
        super().__init__(config)
        self.num_timesteps = config.get("num_timesteps", 1000)
        self.beta_start = config.get("beta_start", 0.0001)
        self.beta_end = config.get("beta_end", 0.02)
        self.register_buffer("betas", torch.linspace(self.beta_start, self.beta_end, self.num_timesteps))
        self.register_buffer("alphas", 1 - self.betas)
        self.register_buffer("alphas_cumprod", torch.cumprod(self.alphas, dim=0))
        if config.get("use_time_conditioning", True):
            raise NotImplementedError(
                "use_time_conditioning=True is not runnable in the reference either: conv1 is widened by model_channels "
                "(energy_based.py:136-138) but nothing produces those channels.  Use use_time_conditioning: False.")
        self.model = EnergyNet(config.get("in_channels", 3), config.get("model_channels", 64), precision=config.get("precision", "fp32"))
        loss_type = config.get("loss_type", "energy_based")
        self.loss_type = loss_type
        self.energy_scale = config.get("energy_scale", 1.0)
        self.regularization_weight = config.get("regularization_weight", 0.1)
        self.loss_fn = None if loss_type == "energy_based" else DiffusionLoss(loss_type=loss_type, loss_config=config.get("loss_config", None))
        self.langevin_steps = config.get("langevin_steps", 10)
        self.langevin_step_size = config.get("langevin_step_size", 0.01)

    def forward(self, x: torch.Tensor, t: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.model(x)

    # ------------------------------------------------------------------ training
    def loss_function(self, x: torch.Tensor) -> torch.Tensor:
        """energy_based.py:179-211 + utils/losses.py:252-286.  RNG order: randint, randn_like, langevin_steps x randn_like,
        then the loss's rand for the interpolation weights."""
        B = x.shape[0]
        t = torch.randint(0, self.num_timesteps, (B,), device=x.device)
        noise = torch.randn_like(x)
        lang = [torch.randn_like(x) for _ in range(self.langevin_steps)]
        alpha = torch.rand(B, 1, 1, 1, device=x.device) if self.loss_fn is None else None
        return self._loss_from_draws(x, t, noise, lang, alpha)

    def _loss_from_draws(self, x, t, noise, lang: List[torch.Tensor], alpha):
        x = x.contiguous().float()
        x_noisy = self._add_noise(x, t, noise)
        x_fake = self._langevin_sampling(x_noisy, t, _noises=lang)
        if self.loss_fn is not None:
            return self.loss_fn(self.forward(x, t), self.forward(x_fake, t), t)
        e_real, e_fake = self.forward(x), self.forward(x_fake)
        cd = e_real.mean() - e_fake.mean()
        lam = float(self.regularization_weight)
        # interpolation x_hat = alpha x + (1 - alpha) x_fake: one fused launch (per-sample a, c coefficients)
        a = alpha.reshape(x.shape[0]).contiguous()
        x_hat = ops.scale_add(x, x_fake, a, 1.0 - a)
        with torch.no_grad():
            _, g = self.model.energy_and_input_grad(x_hat)
            penalty = ((g.norm(2, dim=1) - 1) ** 2).mean()         # [B,3,H,W]-sized reduction of the input gradient
        skip = self.config.get("gradient_penalty_grad", "error") == "skip" or lam == 0.0
        return cd + lam * _PenaltyFn.apply(cd, penalty, skip)

    # ------------------------------------------------------------------ sampling
    def _langevin_sampling(self, x: torch.Tensor, t: torch.Tensor, _noises: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """energy_based.py:250-278: x <- x - eta * grad_x E + sqrt(2 eta) z, one fused update launch per step."""
        x = x.detach()
        for i in range(self.langevin_steps):
            _, grad = self.model.energy_and_input_grad(x)
            z = _noises[i] if _noises is not None else torch.randn_like(x)
            x = ops.langevin_energy_step(x, grad, z, float(self.langevin_step_size))
        return x

    def sample(self, batch_size: int, device: torch.device) -> torch.Tensor:
        """energy_based.py:213-248."""
        x = torch.randn((batch_size, self.config["in_channels"], self.config["image_size"], self.config["image_size"]), device=device)
        with torch.no_grad():
            for t in reversed(range(self.num_timesteps)):
                x = self._langevin_sampling(x, None)
                if t > 0:
                    x = ops.energy_renoise(x, torch.randn_like(x), self.alphas_cumprod, t)
        return x

    def generate_samples(self, batch_size: int, device: torch.device) -> torch.Tensor:
        return self.sample(batch_size, device)

    def _add_noise(self, x: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = torch.randn_like(x)
        return ops.q_sample(x.contiguous(), t, noise, self.alphas_cumprod)
