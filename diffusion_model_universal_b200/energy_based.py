"""Energy-based model behind the reference API (models/energy_based.py) on the CUDA library.

The reference class cannot run as shipped: ``generate_samples`` is abstract, ``forward`` passes ``t`` to a
network that takes only ``x`` (energy_based.py:177), time conditioning widens conv1 by channels nothing produces
(:136-138) and ``torch.sqrt(float)`` raises (:273).  This class implements the repaired behaviour the oracle pins:
``forward(x, t) = EnergyNet(x)``, ``use_time_conditioning`` must be False, ``math.sqrt`` for the step size.

EnergyNet (energy_based.py:51-85) is NOT a UNet: conv3x3(in->C) -> GN(8) -> SiLU -> conv3x3(C->2C) -> GN(8) -> SiLU ->
conv3x3(2C->4C) -> SiLU -> global mean -> Linear(4C, 1).  Forward and the first-order backward (input gradient for the
Langevin dynamics, parameter gradients of the contrastive-divergence terms) run on the same kernels as the UNet.

The gradient penalty of ``EnergyBasedLoss`` (utils/losses.py:277-285, ``create_graph=True``) is differentiated exactly:
the double backward through the network runs on the conv kernels (a convolution's second-order terms are convolutions)
plus ``dmu_gn_bwd_bwd`` and ``dmu_silu_pool_bwd_bwd`` (include/dmu_b200.h).
"""

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _abi, ops
from ._abi import F32, BF16, Tensor4, ConvParams, WgradParams, GnParams, GnBwd2Params, RepackDesc
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
        k = _Launcher(self, x)
        E = k.forward()
        if dE is None:
            return E, None, None
        dx = k.backward(dE, want_dx, want_dw)
        return E, dx, (k.grads if want_dw else None)

    def energy_and_input_grad(self, x: torch.Tensor):
        """(E(x) [B], d sum(E) / dx) in one forward+backward pass: what a Langevin step needs (energy_based.py:266-268)."""
        ones = torch.ones(x.shape[0], device=x.device)
        E, dx, _ = self._run(x.contiguous().float(), ones, True, False)
        return E, dx

    def gradient_penalty(self, x_hat: torch.Tensor, want_param_grads: bool):
        """utils/losses.py:275-285: penalty = mean_{n,h,w} (||grad_x E(x_hat)||_2 over channels - 1)^2, and (optionally) its
        gradient with respect to every parameter - the double backward through the energy network.  Convolutions are linear,
        so their second-order terms are the ordinary fprop / dgrad / wgrad kernels applied to cotangents; the GroupNorm+SiLU
        backward and the SiLU+pool backward have dedicated derivative kernels (dmu_gn_bwd_bwd, dmu_silu_pool_bwd_bwd)."""
        k = _Launcher(self, x_hat.contiguous().float())
        k.forward()
        g = k.backward(torch.ones(x_hat.shape[0], device=x_hat.device), True, False, keep=True)
        nrm = g.norm(2, dim=1, keepdim=True)                       # [B,1,H,W]: host-side reduction of a 3-channel tensor
        penalty = ((nrm - 1) ** 2).mean()
        if not want_param_grads:
            return penalty, None
        v = (2.0 / nrm.numel()) * (nrm - 1) * g / nrm                # d penalty / d g
        k.second_order(v.contiguous())
        return penalty, k.grads


class _Launcher:
    """One evaluation of EnergyNet on the C ABI (eager launches; the network is 3 convolutions)."""

    def __init__(self, net: EnergyNet, x: torch.Tensor):
        self.net, self.x = net, x
        self.lib = _abi.lib()
        self.st = ops._stream()
        self.dev = x.device
        self.code = BF16 if net.precision == "bf16" else F32
        self.tdt = torch.bfloat16 if self.code == BF16 else torch.float32
        self.N, self.Ci, self.H, self.W = x.shape
        Cm = net.model_channels
        self.ch = [self.Ci, Cm, 2 * Cm, 4 * Cm]
        self.convs = [net.conv1, net.conv2, net.conv3]
        self.norms = [net.norm1, net.norm2]
        self.grads = {n: torch.zeros_like(p) for n, p in net.named_parameters()}
        self.keep = []
        # filters in the compute dtype: [O][R][S][I] for fprop, [I][R][S][O] for dgrad
        self.wf, self.wb, descs = [], [], []
        for i, cv in enumerate(self.convs):
            O, I = self.ch[i + 1], self.ch[i]
            f = torch.empty(O * 9 * I, device=self.dev, dtype=self.tdt)
            b = torch.empty(O * 9 * I, device=self.dev, dtype=self.tdt)
            wsrc = cv.weight.detach().contiguous()
            self.keep.append(wsrc)
            descs.append(RepackDesc(wsrc.data_ptr(), f.data_ptr(), O, I, 3, 3, 0, self.code))
            descs.append(RepackDesc(wsrc.data_ptr(), b.data_ptr(), I, O, 3, 3, 1, self.code))
            self.wf.append(f)
            self.wb.append(b)
        arr = (RepackDesc * len(descs))(*descs)
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
        self.keep.append(table)
        self._chk(self.lib.dmu_repack_weights(table.data_ptr(), len(descs), max(c.weight.numel() for c in self.convs), self.st), "repack_weights")

    def _chk(self, rc, what):
        ops._launched()
        _abi.check(rc, what)

    def _act(self, c):
        return torch.empty(self.N, self.H, self.W, c, device=self.dev, dtype=self.tdt)

    # ---- primitives
    def conv(self, i, xt4, yt4, bias=True):
        """layer i fprop: channels ch[i] -> ch[i+1]"""
        Ck, Cj = self.ch[i], self.ch[i + 1]
        p = ConvParams(xt4, yt4, _null(), self.wf[i].data_ptr(), 9 * Ck, 1, Ck, self.convs[i].bias.data_ptr() if bias else None, None, 0,
                       self.N, self.H, self.W, Ck, self.H, self.W, Cj, 3, 3, 1, 1, 0, self.code, 0, 0, None, 0)
        self._chk(self.lib.dmu_conv2d(C.byref(p), self.st), "conv2d")

    def conv_t(self, i, dyt4, dxt4):
        """layer i dgrad: channels ch[i+1] -> ch[i]"""
        Ck, Cj = self.ch[i + 1], self.ch[i]
        p = ConvParams(dyt4, dxt4, _null(), self.wb[i].data_ptr(), 9 * Ck, 1, Ck, None, None, 0,
                       self.N, self.H, self.W, Ck, self.H, self.W, Cj, 3, 3, 1, 1, 1, self.code, 0, 0, None, 0)
        self._chk(self.lib.dmu_conv2d(C.byref(p), self.st), "conv2d dgrad")

    def wgrad(self, i, p4, q4, bias: bool):
        """layer i: dW[o][i][r][s] += sum P[.., o] Q[shifted, i]  (+ dbias += sum P)"""
        name = "conv%d" % (i + 1)
        Ca, Cb = self.ch[i + 1], self.ch[i]
        p = WgradParams(p4, q4, self.grads[name + ".weight"].data_ptr(), Cb * 9, 9, 1, self.grads[name + ".bias"].data_ptr() if bias else None,
                        self.N, self.H, self.W, Ca, self.H, self.W, Cb, 3, 3, 1, 1, 0)
        self._chk(self.lib.dmu_conv2d_wgrad(C.byref(p), self.st), "wgrad " + name)

    def gn_fwd(self, j, xh, yh):
        sums = torch.zeros(self.N, 8, 2, device=self.dev)
        nm = self.norms[j]
        p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(yh), _null(), _null(), _null(), sums.data_ptr(), nm.weight.data_ptr(), nm.bias.data_ptr(),
                     None, None, None, self.N, self.H, self.W, xh.shape[3], 8, 1, 1e-5, 0)
        self._chk(self.lib.dmu_gn_forward(C.byref(p), self.st), "gn_forward")
        return sums

    def gn_bwd(self, j, xh, dyh, dxh, sums, want_dw, add0=None):
        nm = self.norms[j]
        name = "norm%d" % (j + 1)
        red = torch.zeros(self.N, xh.shape[3], 2, device=self.dev)
        p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), ops.t4_nhwc(dxh), ops.t4_nhwc(add0) if add0 is not None else _null(), _null(), sums.data_ptr(),
                     nm.weight.data_ptr(), nm.bias.data_ptr(), red.data_ptr(),
                     self.grads[name + ".weight"].data_ptr() if want_dw else None, self.grads[name + ".bias"].data_ptr() if want_dw else None,
                     self.N, self.H, self.W, xh.shape[3], 8, 1, 1e-5, 0)
        self._chk(self.lib.dmu_gn_backward(C.byref(p), self.st), "gn_backward")

    def gn_bwd2(self, j, xh, dyh, ch, gx, gdy, sums):
        nm = self.norms[j]
        name = "norm%d" % (j + 1)
        p = GnBwd2Params(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), ops.t4_nhwc(ch), ops.t4_nhwc(gx), ops.t4_nhwc(gdy), sums.data_ptr(), nm.weight.data_ptr(),
                         nm.bias.data_ptr(), self.grads[name + ".weight"].data_ptr(), self.grads[name + ".bias"].data_ptr(),
                         self.N, self.H, self.W, xh.shape[3], 8, 1, 1e-5, 0)
        self._chk(self.lib.dmu_gn_bwd_bwd(C.byref(p), self.st), "gn_bwd_bwd")

    # ---- passes
    def forward(self) -> torch.Tensor:
        N, H, W, ch, net = self.N, self.H, self.W, self.ch, self.net
        self.h1 = self._act(ch[1])
        self.conv(0, ops.t4_nchw(self.x), ops.t4_nhwc(self.h1))
        self.a1 = torch.empty_like(self.h1)
        self.s1 = self.gn_fwd(0, self.h1, self.a1)
        self.h2 = self._act(ch[2])
        self.conv(1, ops.t4_nhwc(self.a1), ops.t4_nhwc(self.h2))
        self.a2 = torch.empty_like(self.h2)
        self.s2 = self.gn_fwd(1, self.h2, self.a2)
        self.h3 = self._act(ch[3])
        self.conv(2, ops.t4_nhwc(self.a2), ops.t4_nhwc(self.h3))
        self.pooled = torch.zeros(N, ch[3], device=self.dev)
        t3 = ops.t4_nhwc(self.h3)
        self._chk(self.lib.dmu_silu_pool_fwd(C.byref(t3), N, H, W, ch[3], self.pooled.data_ptr(), ch[3], 1.0 / (H * W), self.st), "silu_pool_fwd")
        E = torch.empty(N, 1, device=self.dev)
        pl = ConvParams(ops.t4_rows(self.pooled), ops.t4_rows(E), _null(), net.dense.weight.data_ptr(), ch[3], 1, 0, net.dense.bias.data_ptr(), None, 0,
                        N, 1, 1, ch[3], 1, 1, 1, 1, 1, 1, 0, 0, F32, 0, 0, None, 0)
        self._chk(self.lib.dmu_conv2d(C.byref(pl), self.st), "dense")
        return E.view(N)

    def _dense_wgrad(self, dE2, q):
        """ddense.weight[0, c] += sum_n dE[n] q[n, c];  ddense.bias += sum dE (only when q is the pooled activations)"""
        ch = self.ch
        p = WgradParams(ops.t4_rows(dE2), ops.t4_rows(q), self.grads["dense.weight"].data_ptr(), ch[3], 1, 0,
                        self.grads["dense.bias"].data_ptr() if q is self.pooled else None, self.N, 1, 1, 1, 1, 1, ch[3], 1, 1, 1, 0, 0)
        self._chk(self.lib.dmu_conv2d_wgrad(C.byref(p), self.st), "wgrad dense")

    def backward(self, dE, want_dx, want_dw, keep=False):
        """First-order backward from dE [N]; returns dx (fp32 NCHW) if wanted.  keep=True retains the intermediate gradients
        the second-order pass differentiates."""
        N, H, W, ch, net = self.N, self.H, self.W, self.ch, self.net
        dE2 = dE.reshape(N, 1).contiguous().float()
        dpooled = torch.empty(N, ch[3], device=self.dev)
        pd = ConvParams(ops.t4_rows(dE2), ops.t4_rows(dpooled), _null(), net.dense.weight.data_ptr(), 1, ch[3], 0, None, None, 0,
                        N, 1, 1, 1, 1, 1, ch[3], 1, 1, 1, 0, 0, F32, 0, 0, None, 0)
        self._chk(self.lib.dmu_conv2d(C.byref(pd), self.st), "dense dgrad")
        if want_dw:
            self._dense_wgrad(dE2, self.pooled)
        dh3 = torch.empty_like(self.h3)
        t3, t3d = ops.t4_nhwc(self.h3), ops.t4_nhwc(dh3)
        self._chk(self.lib.dmu_silu_pool_bwd(C.byref(t3), C.byref(t3d), N, H, W, ch[3], dpooled.data_ptr(), ch[3], 1.0 / (H * W), self.st), "silu_pool_bwd")
        da2 = torch.empty_like(self.a2)
        self.conv_t(2, ops.t4_nhwc(dh3), ops.t4_nhwc(da2))
        if want_dw:
            self.wgrad(2, ops.t4_nhwc(dh3), ops.t4_nhwc(self.a2), True)
        dh2 = torch.empty_like(self.h2)
        self.gn_bwd(1, self.h2, da2, dh2, self.s2, want_dw)
        da1 = torch.empty_like(self.a1)
        self.conv_t(1, ops.t4_nhwc(dh2), ops.t4_nhwc(da1))
        if want_dw:
            self.wgrad(1, ops.t4_nhwc(dh2), ops.t4_nhwc(self.a1), True)
        dh1 = torch.empty_like(self.h1)
        self.gn_bwd(0, self.h1, da1, dh1, self.s1, want_dw)
        dx = None
        if want_dx:
            dx = torch.empty_like(self.x)
            self.conv_t(0, ops.t4_nhwc(dh1), ops.t4_nchw(dx))
        if want_dw:
            self.wgrad(0, ops.t4_nhwc(dh1), ops.t4_nchw(self.x), True)
        if keep:
            self.dE2, self.dpooled, self.dh3, self.da2, self.dh2, self.da1, self.dh1 = dE2, dpooled, dh3, da2, dh2, da1, dh1
        return dx

    def second_order(self, v: torch.Tensor):
        """Accumulate into self.grads the parameter gradient of <v, grad_x E(x)> (v fp32 [N,Cin,H,W], treated as constant):
        back-propagation through the first-order backward pass kept by backward(keep=True), then through the forward pass."""
        N, H, W, ch = self.N, self.H, self.W, self.ch
        # --- through the backward pass, in reverse: g = ConvT1(dh1) <- dh1 = B1(h1, da1) <- da1 = ConvT2(dh2) <- ...
        c_dh1 = self._act(ch[1])
        self.conv(0, ops.t4_nchw(v), ops.t4_nhwc(c_dh1), bias=False)                       # d<v,g>/d dh1 = Conv1(v)
        self.wgrad(0, ops.t4_nhwc(self.dh1), ops.t4_nchw(v), False)                          # dW1 += dh1 (x) v
        c_h1, c_da1 = torch.empty_like(self.h1), torch.empty_like(self.a1)
        self.gn_bwd2(0, self.h1, self.da1, c_dh1, c_h1, c_da1, self.s1)
        c_dh2 = self._act(ch[2])
        self.conv(1, ops.t4_nhwc(c_da1), ops.t4_nhwc(c_dh2), bias=False)
        self.wgrad(1, ops.t4_nhwc(self.dh2), ops.t4_nhwc(c_da1), False)
        c_h2, c_da2 = torch.empty_like(self.h2), torch.empty_like(self.a2)
        self.gn_bwd2(1, self.h2, self.da2, c_dh2, c_h2, c_da2, self.s2)
        c_dh3 = self._act(ch[3])
        self.conv(2, ops.t4_nhwc(c_da2), ops.t4_nhwc(c_dh3), bias=False)
        self.wgrad(2, ops.t4_nhwc(self.dh3), ops.t4_nhwc(c_da2), False)
        c_h3 = torch.empty_like(self.h3)
        c_dpooled = torch.zeros(N, ch[3], device=self.dev)
        t3, tc, tg = ops.t4_nhwc(self.h3), ops.t4_nhwc(c_dh3), ops.t4_nhwc(c_h3)
        self._chk(self.lib.dmu_silu_pool_bwd_bwd(C.byref(t3), C.byref(tc), C.byref(tg), N, H, W, ch[3], self.dpooled.data_ptr(), ch[3],
                                                 c_dpooled.data_ptr(), ch[3], 1.0 / (H * W), self.st), "silu_pool_bwd_bwd")
        self._dense_wgrad(self.dE2, c_dpooled)                                               # dpooled = dE * w_dense
        # --- through the forward pass with the cotangents of h3, h2, h1 collected above
        self.wgrad(2, ops.t4_nhwc(c_h3), ops.t4_nhwc(self.a2), True)
        c_a2 = torch.empty_like(self.a2)
        self.conv_t(2, ops.t4_nhwc(c_h3), ops.t4_nhwc(c_a2))
        t_h2 = torch.empty_like(self.h2)
        self.gn_bwd(1, self.h2, c_a2, t_h2, self.s2, True, add0=c_h2)
        self.wgrad(1, ops.t4_nhwc(t_h2), ops.t4_nhwc(self.a1), True)
        c_a1 = torch.empty_like(self.a1)
        self.conv_t(1, ops.t4_nhwc(t_h2), ops.t4_nhwc(c_a1))
        t_h1 = torch.empty_like(self.h1)
        self.gn_bwd(0, self.h1, c_a1, t_h1, self.s1, True, add0=c_h1)
        self.wgrad(0, ops.t4_nhwc(t_h1), ops.t4_nchw(self.x), True)


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
    """Gradient penalty as a node of the autograd graph: the forward computes the value and (if anything requires grad) its
    parameter gradients with the second-order launches of EnergyNet.gradient_penalty; the backward hands them out."""

    @staticmethod
    def forward(ctx, net: EnergyNet, x_hat, *params):
        need = any(p.requires_grad for p in params)
        value, grads = net.gradient_penalty(x_hat, need)
        ctx.names = [n for n, _ in net.named_parameters()]
        ctx.grads = grads
        return value

    @staticmethod
    def backward(ctx, g):
        if ctx.grads is None:
            return (None, None) + (None,) * len(ctx.names)
        return (None, None) + tuple(ctx.grads[n] * g for n in ctx.names)


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
        params = [p for _, p in self.model.named_parameters()]
        return cd + lam * _PenaltyFn.apply(self.model, x_hat, *params)

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
