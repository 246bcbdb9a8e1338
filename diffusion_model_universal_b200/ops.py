"""Tensor-level wrappers over the C ABI (include/dmu_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; all
arithmetic happens in libdmu_b200.so.  Every wrapper requires CUDA tensors and
raises otherwise — there is no fallback path.
"""

import ctypes as C
import math

import torch

from . import _abi
from ._abi import F32, BF16, Tensor4, ConvParams, WgradParams, GnParams, AttnParams, check


LAUNCHES = 0   # kernels launched through this module / the engine (bench.py reports it as gpu_launches)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("diffusion_model_universal_b200 runs on CUDA (sm_100a) only: got a CPU tensor. "
                               "Move the model and inputs to a B200 device; there is no CPU fallback.")


def _launched(n=1):
    global LAUNCHES
    LAUNCHES += n


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _f32c(t, name):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return t


def _index_vec(t, batch, name):
    """Timestep / table-index vectors are read as int64 [B] by the kernels: anything else would be reinterpreted, not cast."""
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64 (got {t.dtype}); cast with .long()")
    if t.dim() != 1 or t.shape[0] != batch or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous [B] vector with B = {batch}, got shape {tuple(t.shape)}")


def t4_nchw(t: torch.Tensor) -> Tensor4:
    """Describe a contiguous [N,C,H,W] tensor."""
    n, c, h, w = t.shape
    assert t.is_contiguous()
    return Tensor4(t.data_ptr(), c * h * w, w, 1, h * w, dtype_code(t), 0)


def t4_nhwc(t: torch.Tensor, c0: int = 0, c: int = None) -> Tensor4:
    """Describe channels [c0, c0+c) of a contiguous [N,H,W,Ctot] tensor."""
    n, h, w, ct = t.shape
    assert t.is_contiguous()
    return Tensor4(t.data_ptr() + c0 * t.element_size(), h * w * ct, w * ct, ct, 1, dtype_code(t), 0)


def t4_rows(t: torch.Tensor) -> Tensor4:
    """[M, C] row matrix as an N=M, H=W=1 tensor."""
    m, c = t.shape
    assert t.stride(1) == 1
    p = t.stride(0)
    return Tensor4(t.data_ptr(), p, p, p, 1, dtype_code(t), 0)


# ------------------------------------------------------------------ process updates
def snr_time_weights(t, table, min_weight: float, max_weight: float):
    """utils/losses.py:144-181 ('snr'): the [B] weight vector in one launch, t_max kept on the device.  table: fp32 [T, T]."""
    _need_cuda(t, table)
    _index_vec(t, t.shape[0], "t")
    _f32c(table, "table")
    w = torch.empty(t.shape[0], device=t.device, dtype=torch.float32)
    _launched()
    check(_abi.lib().dmu_snr_time_weights(t.data_ptr(), table.data_ptr(), table.shape[0], t.shape[0], float(min_weight),
                                          float(max_weight - min_weight), w.data_ptr(), _stream()), "snr_time_weights")
    return w


def q_sample(x0, t, noise, alphas_cumprod, out=None):
    """models/ddpm.py:286-296."""
    _need_cuda(x0, t, noise, alphas_cumprod)
    x0, noise = _f32c(x0, "x0"), _f32c(noise, "noise")
    _index_vec(t, x0.shape[0], "t")
    _f32c(alphas_cumprod, "alphas_cumprod")
    out = torch.empty_like(x0) if out is None else _f32c(out, "out")
    b = x0.shape[0]
    _launched()
    check(_abi.lib().dmu_q_sample(x0.data_ptr(), noise.data_ptr(), t.data_ptr(), alphas_cumprod.data_ptr(), out.data_ptr(),
                                  b, x0.numel() // max(b, 1), _stream()), "q_sample")
    return out


def ingest_u8(img, mean=None, std=None, layout="NCHW", t=None, noise=None, alphas_cumprod=None, want_x0=True, xt_out=None):
    """datasets/dataset_utils.py:58-61 (ToTensor + Normalize) on the device, optionally fused with q_sample
    (models/ddpm.py:286-296).  img: uint8 [B,C,H,W] (layout "NCHW") or [B,H,W,C] ("NHWC"); mean / std: fp32 [C] device
    tensors or None.  Returns (x0, xt) as fp32 [B,C,H,W]; x0 is None when want_x0 is False, xt is None without noise."""
    _need_cuda(img, mean, std, t, noise, alphas_cumprod)
    if img.dtype != torch.uint8 or img.dim() != 4 or not img.is_contiguous():
        raise TypeError("img must be a contiguous 4-D uint8 tensor")
    if layout not in ("NCHW", "NHWC"):
        raise ValueError(f"layout must be 'NCHW' or 'NHWC', got {layout!r}")
    if layout == "NHWC":
        b, h, w, c = img.shape
    else:
        b, c, h, w = img.shape
    for v, name in ((mean, "mean"), (std, "std")):
        if v is not None and (v.dtype != torch.float32 or v.numel() != c or not v.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous float32 tensor with one entry per channel")
    fused = noise is not None
    if fused:
        if t is None or alphas_cumprod is None:
            raise ValueError("the fused q_sample needs t, noise and alphas_cumprod")
        _f32c(noise, "noise")
        if tuple(noise.shape) != (b, c, h, w):
            raise ValueError("noise must be [B,C,H,W]")
        if t.dtype != torch.int64 or t.numel() != b:
            raise TypeError("t must be int64 [B]")
    elif not want_x0:
        raise ValueError("nothing to compute: want_x0 is False and no noise was given")
    x0 = torch.empty((b, c, h, w), device=img.device, dtype=torch.float32) if want_x0 else None
    xt = (torch.empty((b, c, h, w), device=img.device, dtype=torch.float32) if xt_out is None else _f32c(xt_out, "xt_out")) if fused else None
    ptr = lambda v: v.data_ptr() if v is not None else None
    _launched()
    check(_abi.lib().dmu_ingest_u8(img.data_ptr(), 1 if layout == "NHWC" else 0, ptr(mean), ptr(std), ptr(noise),
                                   ptr(t) if fused else None, ptr(alphas_cumprod) if fused else None, ptr(x0), ptr(xt),
                                   b, c, h * w, _stream()), "ingest_u8")
    return x0, xt


def image_grid_shape(n, c, h, w, nrow=8, padding=2):
    """(grid_h, grid_w, grid_c) of torchvision.utils.make_grid for n images of [c,h,w]."""
    gh, gw, gc = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    check(_abi.lib().dmu_image_grid_shape(n, c, h, w, nrow, padding, C.byref(gh), C.byref(gw), C.byref(gc)), "image_grid_shape")
    return gh.value, gw.value, gc.value


def image_grid_u8(x, nrow=8, padding=2, pad_value=0.0, transpose=False, value_range=None):
    """trainers/ddpm_trainer.py:821-834: make_grid + save_image's 8-bit quantisation in one launch.
    value_range=(lo, hi) is make_grid's normalize=True, value_range=... (scripts/generate.py:119-133).
    x: fp32 [N,C,H,W] (cells filled in order) or [A,B,C,H,W]; with transpose=True the 5-D tensor is read as
    image k = x[k % A, k // A] — the trainer's "row per sample, column per saved step" view of a stacked list of
    intermediates.  Returns uint8 [grid_h, grid_w, grid_c] on the device."""
    _need_cuda(x)
    _f32c(x, "x")
    if x.dim() == 4:
        n, c, h, w = x.shape
        period, s_mod, s_div = max(n, 1), c * h * w, 0
    elif x.dim() == 5:
        a, b, c, h, w = x.shape
        n = a * b
        if transpose:
            period, s_mod, s_div = a, b * c * h * w, c * h * w
        else:
            period, s_mod, s_div = max(n, 1), c * h * w, 0
    else:
        raise ValueError("x must be [N,C,H,W] or [A,B,C,H,W]")
    gh, gw, gc = image_grid_shape(n, c, h, w, nrow, padding)
    out = torch.empty((gh, gw, gc), device=x.device, dtype=torch.uint8)
    _launched()
    if value_range is None:
        check(_abi.lib().dmu_image_grid_u8(x.data_ptr(), n, period, s_mod, s_div, c, h, w, nrow, padding, float(pad_value),
                                           out.data_ptr(), _stream()), "image_grid_u8")
    else:
        lo, hi = (float(v) for v in value_range)
        check(_abi.lib().dmu_image_grid_range_u8(x.data_ptr(), n, period, s_mod, s_div, c, h, w, nrow, padding, float(pad_value),
                                                 lo, hi, out.data_ptr(), _stream()), "image_grid_range_u8")
    return out


def ddpm_step(x, eps, t, noise, betas, alphas, alphas_cumprod, out=None):
    """models/ddpm.py:306-329 after the eps prediction (noise=None only when t == 0)."""
    _need_cuda(x, eps, t)
    x, eps = _f32c(x, "x"), _f32c(eps, "eps")
    if noise is not None:
        _f32c(noise, "noise")
    _index_vec(t, x.shape[0], "t")
    if not (betas.numel() == alphas.numel() == alphas_cumprod.numel()):
        raise ValueError("betas, alphas and alphas_cumprod must have the same length")
    out = torch.empty_like(x) if out is None else out
    b = x.shape[0]
    _launched()
    check(_abi.lib().dmu_ddpm_step(x.data_ptr(), eps.data_ptr(), noise.data_ptr() if noise is not None else None, t.data_ptr(),
                                   betas.data_ptr(), alphas.data_ptr(), alphas_cumprod.data_ptr(), betas.numel(), out.data_ptr(),
                                   b, x.numel() // max(b, 1), _stream()), "ddpm_step")
    return out


def ddim_step(x, eps, idx, noise, alphas, alphas_prev, sigmas, sqrt_one_minus_alphas, out=None):
    """models/ddim.py:97-124 after the eps prediction; idx indexes the S-entry tables."""
    _need_cuda(x, eps, idx)
    x, eps = _f32c(x, "x"), _f32c(eps, "eps")
    _index_vec(idx, x.shape[0], "idx")
    out = torch.empty_like(x) if out is None else out
    b = x.shape[0]
    _launched()
    check(_abi.lib().dmu_ddim_step(x.data_ptr(), eps.data_ptr(), noise.data_ptr() if noise is not None else None, idx.data_ptr(),
                                   alphas.data_ptr(), alphas_prev.data_ptr(), sigmas.data_ptr(), sqrt_one_minus_alphas.data_ptr(),
                                   out.data_ptr(), b, x.numel() // max(b, 1), _stream()), "ddim_step")
    return out


def langevin_score_step(x, score, noise, sigmas, k: int, beta: float, out=None):
    """models/score_based.py:236-245."""
    _need_cuda(x, score, noise, sigmas)
    out = torch.empty_like(x) if out is None else out
    _launched()
    check(_abi.lib().dmu_langevin_score_step(_f32c(x, "x").data_ptr(), _f32c(score, "score").data_ptr(), _f32c(noise, "noise").data_ptr(),
                                             sigmas.data_ptr(), k, beta, out.data_ptr(), x.numel(), _stream()), "langevin_score_step")
    return out


def langevin_energy_step(x, grad, noise, step_size: float, out=None):
    """models/energy_based.py:271-273 (math.sqrt repair)."""
    _need_cuda(x, grad, noise)
    out = torch.empty_like(x) if out is None else out
    _launched()
    check(_abi.lib().dmu_langevin_energy_step(_f32c(x, "x").data_ptr(), _f32c(grad, "grad").data_ptr(), _f32c(noise, "noise").data_ptr(),
                                              step_size, math.sqrt(2 * step_size), out.data_ptr(), x.numel(), _stream()), "langevin_energy_step")
    return out


def energy_renoise(x, noise, alphas_cumprod, t: int, out=None):
    """models/energy_based.py:240-246."""
    _need_cuda(x, noise, alphas_cumprod)
    out = torch.empty_like(x) if out is None else out
    _launched()
    check(_abi.lib().dmu_energy_renoise(_f32c(x, "x").data_ptr(), _f32c(noise, "noise").data_ptr(), alphas_cumprod.data_ptr(), t,
                                        out.data_ptr(), x.numel(), _stream()), "energy_renoise")
    return out


def scale_add(x, z, a, c, out=None):
    """out[b] = a[b]*x[b] + c[b]*z[b] (a None = 1): score_based.py:200-201, losses.py:240."""
    _need_cuda(x, z, c)
    x, z = _f32c(x, "x"), _f32c(z, "z")
    out = torch.empty_like(x) if out is None else out
    b = x.shape[0]
    c = c.float().contiguous()
    a = a.float().contiguous() if a is not None else None
    _launched()
    check(_abi.lib().dmu_scale_add(x.data_ptr(), z.data_ptr(), a.data_ptr() if a is not None else None, c.data_ptr(), out.data_ptr(),
                                   b, x.numel() // max(b, 1), _stream()), "scale_add")
    return out


def diffusion_loss(pred, target, w, wm, wl, wh, delta, want_grad: bool, dpred_out=None):
    """utils/losses.py:74-131 given per-sample weights w [B] (or None).  Returns (loss 0-dim, dpred or None); dpred_out = an
    existing fp32 tensor of pred's shape to receive the gradient (the engine's static upstream-gradient buffer)."""
    _need_cuda(pred, target, w)
    pred, target = _f32c(pred, "pred"), _f32c(target, "target")
    b = pred.shape[0]
    n = pred.numel()
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    dpred = (torch.empty_like(pred) if dpred_out is None else _f32c(dpred_out, "dpred_out")) if want_grad else None
    part = torch.empty(_abi.lib().dmu_loss_workspace_floats(n), device=pred.device, dtype=torch.float32)
    _launched(2)
    check(_abi.lib().dmu_diffusion_loss(pred.data_ptr(), target.data_ptr(), w.data_ptr() if w is not None else None,
                                        wm, wl, wh, delta, loss.data_ptr(), dpred.data_ptr() if want_grad else None,
                                        part.data_ptr(), b, n // b, _stream()), "diffusion_loss")
    return loss, dpred


# ------------------------------------------------------------------ UNet primitives (used by tests; the engine builds structs directly)
def conv2d_raw(p: ConvParams):
    _launched()
    check(_abi.lib().dmu_conv2d(C.byref(p), _stream()), "conv2d")


def wgrad_raw(p: WgradParams):
    _launched()
    check(_abi.lib().dmu_conv2d_wgrad(C.byref(p), _stream()), "conv2d_wgrad")


def copy4(src: Tensor4, dst: Tensor4, n, h, w, c):
    _launched()
    check(_abi.lib().dmu_copy4(C.byref(src), C.byref(dst), n, h, w, c, _stream()), "copy4")


def nchw_to_nhwc(x: torch.Tensor, dtype=None) -> torch.Tensor:
    _need_cuda(x)
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), device=x.device, dtype=dtype or x.dtype)
    copy4(t4_nchw(x.contiguous()), t4_nhwc(out), n, h, w, c)
    return out


def nhwc_to_nchw(x: torch.Tensor, dtype=None) -> torch.Tensor:
    _need_cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), device=x.device, dtype=dtype or x.dtype)
    copy4(t4_nhwc(x.contiguous()), t4_nchw(out), n, h, w, c)
    return out
