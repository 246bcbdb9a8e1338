"""ctypes binding of libdmu_b200.so (include/dmu_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing
this module never needs a GPU (so CPU-only tooling can import the package and
check the exported symbols), but every compute entry point runs CUDA kernels.
"""

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdmu_b200.so")

F32, BF16 = 0, 1

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class Tensor4(C.Structure):
    _fields_ = [("ptr", c_vp), ("sn", c_i64), ("sh", c_i64), ("sw", c_i64), ("sc", c_i64), ("dtype", c_i32), ("_pad", c_i32)]


class ConvParams(C.Structure):
    _fields_ = [
        ("x", Tensor4), ("y", Tensor4), ("res", Tensor4),
        ("w", c_vp), ("w_sn", c_i64), ("w_sk", c_i64), ("w_st", c_i64),
        ("bias", c_vp), ("temb", c_vp), ("temb_pitch", c_i64),
        ("N", c_i32), ("Hi", c_i32), ("Wi", c_i32), ("Ck", c_i32),
        ("Ho", c_i32), ("Wo", c_i32), ("Cj", c_i32),
        ("R", c_i32), ("S", c_i32), ("stride", c_i32), ("pad", c_i32),
        ("gather", c_i32), ("w_dtype", c_i32), ("impl", c_i32), ("_pad", c_i32),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
        ("gn_coef", c_vp), ("gn_silu", c_i32), ("_pad2", c_i32), ("a_out", Tensor4),
        ("gn_fuse", c_vp), ("gn_fuse_mode", c_i32), ("_pad3", c_i32),
    ]


class WgradParams(C.Structure):
    _fields_ = [
        ("p", Tensor4), ("q", Tensor4),
        ("dw", c_vp), ("dw_sa", c_i64), ("dw_sb", c_i64), ("dw_st", c_i64),
        ("dbias", c_vp),
        ("N", c_i32), ("Hp", c_i32), ("Wp", c_i32), ("Ca", c_i32),
        ("Hq", c_i32), ("Wq", c_i32), ("Cb", c_i32),
        ("R", c_i32), ("S", c_i32), ("stride", c_i32), ("pad", c_i32),
        ("impl", c_i32),
    ]


class GnParams(C.Structure):
    _fields_ = [
        ("x", Tensor4), ("y", Tensor4), ("dx", Tensor4), ("add0", Tensor4), ("add1", Tensor4),
        ("sums", c_vp), ("gamma", c_vp), ("beta", c_vp), ("red", c_vp), ("dgamma", c_vp), ("dbeta", c_vp),
        ("N", c_i32), ("H", c_i32), ("W", c_i32), ("C", c_i32), ("G", c_i32),
        ("silu", c_i32), ("eps", c_f32), ("flags", c_i32),      # flags: GN_FIXED_SUMS
    ]


GN_FIXED_SUMS = 1      # include/dmu_b200.h DMU_GN_FIXED_SUMS: int64 fixed-point accumulators behind the float sums


class GnBwd2Params(C.Structure):
    _fields_ = [
        ("x", Tensor4), ("dy", Tensor4), ("c", Tensor4), ("gx", Tensor4), ("gdy", Tensor4),
        ("sums", c_vp), ("gamma", c_vp), ("beta", c_vp), ("dgamma", c_vp), ("dbeta", c_vp),
        ("N", c_i32), ("H", c_i32), ("W", c_i32), ("C", c_i32), ("G", c_i32), ("silu", c_i32), ("eps", c_f32), ("_pad", c_i32),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("qkv", c_vp), ("qkv_pitch", c_i64), ("o", c_vp), ("o_pitch", c_i64),
        ("d_o", c_vp), ("do_pitch", c_i64), ("dqkv", c_vp), ("dqkv_pitch", c_i64),
        ("lse", c_vp),
        ("N", c_i32), ("S", c_i32), ("C", c_i32), ("heads", c_i32), ("dtype", c_i32), ("_pad", c_i32),
    ]


class GnPgDesc(C.Structure):
    _fields_ = [("red", c_vp), ("dgamma", c_vp), ("dbeta", c_vp), ("C", c_i32), ("count", c_i32)]


class RepackDesc(C.Structure):
    _fields_ = [("src", c_vp), ("dst", c_vp), ("O", c_i32), ("I", c_i32), ("R", c_i32), ("S", c_i32), ("kind", c_i32), ("dst_dtype", c_i32)]


class ColsumDesc(C.Structure):
    _fields_ = [("x", Tensor4), ("out_nc", c_vp), ("pitch", c_i64), ("out_c", c_vp), ("N", c_i32), ("H", c_i32), ("W", c_i32), ("C", c_i32),
                ("chunks", c_i32), ("scale", c_f32), ("cta0", c_i32), ("_pad", c_i32)]


STRUCTS = {
    "dmu_tensor4": Tensor4, "dmu_conv_params": ConvParams, "dmu_wgrad_params": WgradParams,
    "dmu_gn_params": GnParams, "dmu_attn_params": AttnParams, "dmu_repack_desc": RepackDesc, "dmu_gn_bwd2_params": GnBwd2Params,
    "dmu_colsum_desc": ColsumDesc,
}

P = C.POINTER
_SIGS = {
    "dmu_abi_version": (c_i32, []),
    "dmu_last_error": (C.c_char_p, []),
    "dmu_sizeof": (c_i32, [C.c_char_p]),
    "dmu_q_sample": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "dmu_ddpm_step": (c_i32, [c_vp] * 7 + [c_i64, c_vp, c_i64, c_i64, c_vp]),
    "dmu_ddim_step": (c_i32, [c_vp] * 9 + [c_i64, c_i64, c_vp]),
    "dmu_langevin_score_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_i64, c_vp]),
    "dmu_langevin_energy_step": (c_i32, [c_vp, c_vp, c_vp, c_f32, c_f32, c_vp, c_i64, c_vp]),
    "dmu_energy_renoise": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "dmu_scale_add": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "dmu_snr_time_weights": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "dmu_loss_workspace_floats": (c_i64, [c_i64]),
    "dmu_diffusion_loss": (c_i32, [c_vp, c_vp, c_vp, c_f32, c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp]),
    "dmu_conv2d": (c_i32, [P(ConvParams), c_vp]),
    "dmu_conv2d_workspace_bytes": (c_i64, []),
    "dmu_conv2d_gn_supported": (c_i32, [P(ConvParams)]),
    "dmu_conv2d_gn_fuse_supported": (c_i32, [P(ConvParams)]),
    "dmu_gn_coef": (c_i32, [P(GnParams), c_vp, c_vp]),
    "dmu_conv2d_wgrad": (c_i32, [P(WgradParams), c_vp]),
    "dmu_gn_forward": (c_i32, [P(GnParams), c_vp]),
    "dmu_gn_backward": (c_i32, [P(GnParams), c_vp]),
    "dmu_gn_param_grads": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp]),
    "dmu_gn_stats": (c_i32, [P(GnParams), c_vp]),
    "dmu_gn_apply": (c_i32, [P(GnParams), c_vp]),
    "dmu_gn_bwd_reduce": (c_i32, [P(GnParams), c_vp]),
    "dmu_gn_bwd_apply": (c_i32, [P(GnParams), c_vp]),
    "dmu_colsum": (c_i32, [P(Tensor4), c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_f32, c_vp]),
    "dmu_colsum_multi": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp]),
    "dmu_silu_pool_fwd": (c_i32, [P(Tensor4), c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_f32, c_vp]),
    "dmu_silu_pool_bwd": (c_i32, [P(Tensor4), P(Tensor4), c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_f32, c_vp]),
    "dmu_gn_bwd_bwd": (c_i32, [P(GnBwd2Params), c_vp]),
    "dmu_silu_pool_bwd_bwd": (c_i32, [P(Tensor4), P(Tensor4), P(Tensor4), c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_f32, c_vp]),
    "dmu_attn_fwd": (c_i32, [P(AttnParams), c_vp]),
    "dmu_attn_bwd": (c_i32, [P(AttnParams), c_vp]),
    "dmu_sinusoidal_embedding": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_i32, c_vp]),
    "dmu_act_fwd": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dmu_act_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    "dmu_repack_weights": (c_i32, [c_vp, c_i32, c_i64, c_vp]),
    "dmu_zero": (c_i32, [c_vp, c_i64, c_vp]),
    "dmu_copy4": (c_i32, [P(Tensor4), P(Tensor4), c_i32, c_i32, c_i32, c_i32, c_vp]),
    "dmu_adam_ema": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "dmu_ingest_u8": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i64, c_vp]),
    "dmu_image_grid_shape": (c_i32, [c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, P(c_i64), P(c_i64), P(c_i32)]),
    "dmu_image_grid_u8": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp]),
    "dmu_image_grid_range_u8": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32, C.c_double, C.c_double,
                                        c_vp, c_vp]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is the only compute path of this package. "
                "Build it with `python -m diffusion_model_universal_b200.build_ext` (needs nvcc, sm_100a).")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        if h.dmu_abi_version() != 1:
            raise RuntimeError("libdmu_b200.so ABI version mismatch")
        for name, st in STRUCTS.items():
            if h.dmu_sizeof(name.encode()) != C.sizeof(st):
                raise RuntimeError(f"struct {name}: ctypes mirror is {C.sizeof(st)} B, library says {h.dmu_sizeof(name.encode())} B")
        _lib = h
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise RuntimeError(f"dmu_b200 {what}: {lib().dmu_last_error().decode(errors='replace')}")
