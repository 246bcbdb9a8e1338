"""GPU parity of the UNet primitives (C ABI) against ATen fp32 on the same device."""

import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 2e-5, torch.bfloat16: 6e-3}


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def _mods():
    from diffusion_model_universal_b200 import ops, _abi
    return ops, _abi


def _null():
    from diffusion_model_universal_b200._abi import Tensor4
    return Tensor4(None, 0, 0, 0, 0, 0, 0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _repack(w, transposed, dtype, dgrad=False):
    """OIHW (or IOHW for ConvTranspose2d) -> [O][R][S][I] via the library's own repack kernel.
    dgrad=True gives the layout of the input-gradient contraction instead: [I][R][S][O] (rows = original input channels)."""
    ops, _abi = _mods()
    w = w.contiguous()
    if dgrad:   # the same kernel with the roles of the two channel axes exchanged
        transposed = not transposed
    O, I = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
    dst = torch.empty(O * w.shape[2] * w.shape[3] * I, device=w.device, dtype=dtype)
    d = _abi.RepackDesc(w.data_ptr(), dst.data_ptr(), O, I, w.shape[2], w.shape[3], 1 if transposed else 0, ops.dtype_code(dst))
    tab = torch.frombuffer(bytearray(bytes(d)), dtype=torch.uint8).cuda()
    _abi.check(_abi.lib().dmu_repack_weights(tab.data_ptr(), 1, w.numel(), _stream()))
    torch.cuda.synchronize()
    return dst


CONV_CASES = [
    # N, H, W, Ci, Co, R, stride, pad, kind
    (2, 32, 32, 64, 64, 3, 1, 1, "conv"),
    (3, 16, 16, 64, 128, 3, 1, 1, "conv"),
    (2, 8, 8, 192, 64, 3, 1, 1, "conv"),
    (4, 4, 4, 128, 128, 3, 1, 1, "conv"),
    (5, 1, 1, 512, 256, 3, 1, 1, "conv"),
    (3, 2, 2, 384, 128, 3, 1, 1, "conv"),
    (2, 8, 8, 64, 128, 1, 1, 0, "conv"),
    (2, 32, 32, 64, 64, 4, 2, 1, "conv"),
    (3, 2, 2, 256, 256, 4, 2, 1, "conv"),
    (2, 16, 16, 64, 64, 4, 2, 1, "convT"),
    (3, 1, 1, 256, 256, 4, 2, 1, "convT"),
    (2, 12, 20, 24, 40, 3, 1, 1, "conv"),     # ragged: non-square, channel counts off the tile sizes
]


TC_CASES = [c for c in CONV_CASES if c[3] % 64 == 0 and c[4] % 64 == 0] + [
    (128, 1, 1, 256, 256, 3, 1, 1, "conv"),   # bottleneck at the bench batch: 8 of 9 taps are entirely padding
    (16, 32, 32, 64, 64, 3, 1, 1, "conv"),    # several pixel tiles per image, several images
    (9, 8, 8, 128, 256, 3, 1, 1, "conv"),     # batch not a multiple of the images-per-box
    (8, 4, 4, 128, 128, 4, 2, 1, "convT"),
    (4, 16, 16, 128, 64, 1, 1, 0, "conv"),
]


HALO_CASES = [
    (2, 32, 32, 64, 64, 3, 1, 1, "conv"),       # resident weights, 4-stage halo ring
    (128, 32, 32, 64, 64, 3, 1, 1, "conv"),     # the bench layer: 7 tiles per CTA
    (3, 16, 16, 64, 128, 3, 1, 1, "conv"),      # NT = 128, resident
    (5, 16, 16, 128, 64, 3, 1, 1, "conv"),      # two 64-channel chunks, resident
    (4, 8, 8, 128, 128, 3, 1, 1, "conv"),       # streamed weights
    (3, 8, 8, 192, 64, 3, 1, 1, "conv"),        # three chunks, streamed
    (7, 8, 8, 64, 64, 3, 1, 1, "conv"),         # tiles straddle images
    (2, 64, 64, 64, 64, 3, 1, 1, "conv"),       # 64x64 (DDIM config)
    (1, 8, 16, 128, 128, 3, 1, 1, "conv"),      # non-square
]


@pytest.mark.parametrize("impl", [2, 4])
@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05(case, impl):
    """The tcgen05/TMA implicit-GEMM kernels against ATen on bf16-rounded operands: impl 2 = auto (persistent halo kernel
    for 3x3 stride-1 layers of 8x8 and up, per-tap kernel otherwise), impl 4 = per-tap kernel only."""
    _conv_case(case, torch.bfloat16, impl)


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_tcgen05_halo(case):
    """impl 5 = the persistent halo kernel, forced (impl 2 only picks it for large single-chunk layers)."""
    _conv_case(case, torch.bfloat16, 5)


SPLITK_CASES = [
    (128, 1, 1, 512, 256, 3, 1, 1, "conv"),
    (128, 2, 2, 384, 128, 3, 1, 1, "conv"),
    (32, 4, 4, 256, 128, 3, 1, 1, "conv"),
    (64, 2, 2, 256, 256, 4, 2, 1, "conv"),
    (64, 2, 2, 256, 256, 4, 2, 1, "convT"),
    (5, 4, 4, 128, 128, 3, 1, 1, "conv"),
]


@pytest.mark.parametrize("case", SPLITK_CASES)
def test_conv_tcgen05_split_k(case):
    """Same kernels with the split-K scratch supplied: few output tiles, long contraction -> a cluster of CTAs per tile.
    The scratch starts as garbage (NaN bit patterns) and is reused by a second launch: its prior contents must not matter."""
    ops, _abi = _mods()
    ws = torch.full((int(_abi.lib().dmu_conv2d_workspace_bytes()),), 0xFF, device="cuda:0", dtype=torch.uint8)
    _conv_case(case, torch.bfloat16, 2, ws)
    _conv_case(case, torch.bfloat16, 2, ws)


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_fprop_dgrad_wgrad(case, dtype):
    _conv_case(case, dtype, 1)


def _conv_case(case, dtype, impl, ws=None):
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, WgradParams
    N, H, W, Ci, Co, R, stride, pad, kind = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(sum(v for v in case if isinstance(v, int)))
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    if kind == "conv":
        w = (torch.randn(Co, Ci, R, R, generator=g) / math.sqrt(Ci * R * R)).to(dev)
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    else:
        w = (torch.randn(Ci, Co, R, R, generator=g) / math.sqrt(Ci * R * R / 4)).to(dev)
        Ho, Wo = (H - 1) * stride - 2 * pad + R, (W - 1) * stride - 2 * pad + R
    bias = torch.randn(Co, generator=g).to(dev)
    temb = torch.randn(N, Co + 8, generator=g).to(dev)     # pitched rows
    res_full = torch.randn(N, Ho, Wo, Co + 16, generator=g).to(dev).to(dtype)   # residual read from a channel slice
    xh = ops.nchw_to_nhwc(x, dtype)
    xq, wq = xh.float().permute(0, 3, 1, 2), w
    if dtype == torch.bfloat16:
        wq = w.to(dtype).float()
    wk = _repack(w, kind == "convT", dtype)
    y_full = torch.zeros(N, Ho, Wo, Co + 24, device=dev, dtype=dtype)           # output written into a channel slice
    code = ops.dtype_code(xh)
    p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y_full, 8, Co), ops.t4_nhwc(res_full, 16, Co), wk.data_ptr(), R * R * Ci, 1, Ci,
                   bias.data_ptr(), temb.data_ptr() + 4 * 4, Co + 8, N, H, W, Ci, Ho, Wo, Co, R, R, stride, pad,
                   0 if kind == "conv" else 1, code, impl, 0, ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0)
    ops.conv2d_raw(p)
    if kind == "conv":
        ref = F.conv2d(xq, wq, bias, stride=stride, padding=pad)
    else:
        ref = F.conv_transpose2d(xq, wq, bias, stride=stride, padding=pad)
    ref = ref + temb[:, 4:4 + Co, None, None] + res_full[..., 16:16 + Co].float().permute(0, 3, 1, 2)
    got = y_full[..., 8:8 + Co].float().permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < TOL[dtype], "fprop"
    assert y_full[..., :8].abs().max() == 0 and y_full[..., 8 + Co:].abs().max() == 0, "wrote outside its channel slice"

    # ---- dgrad: gradient w.r.t. x of the same layer
    dy = torch.randn(N, Co, Ho, Wo, generator=g).to(dev)
    dyh = ops.nchw_to_nhwc(dy, dtype)
    dyq = dyh.float().permute(0, 3, 1, 2)
    dx = torch.empty(N, H, W, Ci, device=dev, dtype=dtype)
    if impl in (2, 4, 5):   # tensor-core path contracts over a K-contiguous filter: [Ci][R][S][Co]
        wkt = _repack(w, kind == "convT", dtype, dgrad=True)
        p2 = ConvParams(ops.t4_nhwc(dyh), ops.t4_nhwc(dx), _null(), wkt.data_ptr(), R * R * Co, 1, Co, None, None, 0,
                        N, Ho, Wo, Co, H, W, Ci, R, R, stride, pad, 1 if kind == "conv" else 0, code, impl, 0,
                        ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0)
    else:
        p2 = ConvParams(ops.t4_nhwc(dyh), ops.t4_nhwc(dx), _null(), wk.data_ptr(), 1, R * R * Ci, Ci, None, None, 0,
                        N, Ho, Wo, Co, H, W, Ci, R, R, stride, pad, 1 if kind == "conv" else 0, code, impl, 0)
    ops.conv2d_raw(p2)
    if kind == "conv":
        dref = torch.nn.grad.conv2d_input((N, Ci, H, W), wq, dyq, stride=stride, padding=pad)
    else:
        dref = F.conv2d(dyq, wq, stride=stride, padding=pad)
    assert rel_l2(dx.float().permute(0, 3, 1, 2), dref) < TOL[dtype], "dgrad"

    # ---- wgrad (+ bias grad) in the parameter's own layout, accumulated onto existing content
    dw = torch.ones_like(w)
    db = torch.ones(Co, device=dev)
    if kind == "conv":
        p3 = WgradParams(ops.t4_nhwc(dyh), ops.t4_nhwc(xh), dw.data_ptr(), Ci * R * R, R * R, 1, db.data_ptr(),
                         N, Ho, Wo, Co, H, W, Ci, R, R, stride, pad, impl)
        wref = torch.nn.grad.conv2d_weight(xq, w.shape, dyq, stride=stride, padding=pad)
    else:
        p3 = WgradParams(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), dw.data_ptr(), Co * R * R, R * R, 1, None,
                         N, H, W, Ci, Ho, Wo, Co, R, R, stride, pad, impl)
        wref = torch.nn.grad.conv2d_weight(dyq, (Ci, Co, R, R), xq, stride=stride, padding=pad)
    ops.wgrad_raw(p3)
    assert rel_l2(dw - 1, wref) < 5e-5, "wgrad"
    if kind == "conv":
        assert rel_l2(db - 1, dyq.sum(dim=(0, 2, 3))) < 5e-5, "dbias"


# Every distinct implicit-GEMM shape of the C = 64 UNet at 32 x 32 input (M = B Ho Wo, N = C_out, K = C_in kh kw),
# at the benchmark's batch: 3x3 layers incl. the eight concatenated inputs of the up path, 4x4 stride-2 down- and up-sampling, 1x1 shortcuts.
# (The two 3-channel boundary layers have their own test: test_stem_and_head_layouts.)
UNET_SHAPES = [
    (128, 32, 32, 64, 64, 3, 1, 1, "conv"), (128, 16, 16, 64, 64, 3, 1, 1, "conv"), (128, 16, 16, 128, 64, 3, 1, 1, "conv"),
    (128, 8, 8, 128, 128, 3, 1, 1, "conv"), (128, 8, 8, 192, 64, 3, 1, 1, "conv"), (128, 8, 8, 64, 64, 3, 1, 1, "conv"),
    (128, 8, 8, 64, 128, 3, 1, 1, "conv"), (128, 4, 4, 128, 128, 3, 1, 1, "conv"), (128, 4, 4, 256, 128, 3, 1, 1, "conv"),
    (128, 2, 2, 256, 256, 3, 1, 1, "conv"), (128, 2, 2, 384, 128, 3, 1, 1, "conv"), (128, 2, 2, 128, 128, 3, 1, 1, "conv"),
    (128, 2, 2, 128, 256, 3, 1, 1, "conv"), (128, 1, 1, 256, 256, 3, 1, 1, "conv"), (128, 1, 1, 512, 256, 3, 1, 1, "conv"),
    (128, 32, 32, 64, 64, 4, 2, 1, "conv"), (128, 16, 16, 64, 64, 4, 2, 1, "conv"), (128, 8, 8, 128, 128, 4, 2, 1, "conv"),
    (128, 4, 4, 128, 128, 4, 2, 1, "conv"), (128, 2, 2, 256, 256, 4, 2, 1, "conv"),
    (128, 1, 1, 256, 256, 4, 2, 1, "convT"), (128, 2, 2, 128, 128, 4, 2, 1, "convT"), (128, 4, 4, 128, 128, 4, 2, 1, "convT"),
    (128, 8, 8, 64, 64, 4, 2, 1, "convT"), (128, 16, 16, 64, 64, 4, 2, 1, "convT"),
    (128, 16, 16, 128, 64, 1, 1, 0, "conv"), (128, 8, 8, 192, 64, 1, 1, 0, "conv"), (128, 8, 8, 64, 128, 1, 1, 0, "conv"),
    (128, 4, 4, 256, 128, 1, 1, 0, "conv"), (128, 2, 2, 384, 128, 1, 1, 0, "conv"), (128, 2, 2, 128, 256, 1, 1, 0, "conv"),
    (128, 1, 1, 512, 256, 1, 1, 0, "conv"),
]


@pytest.mark.parametrize("case", UNET_SHAPES)
def test_every_unet_gemm_shape_at_the_bench_batch(case):
    """fprop (+bias +temb +residual), dgrad and wgrad of every layer shape of the network at B = 128, through the kernels the engine's
    own dispatch picks (impl 2), against ATen fp32 on the same bf16-rounded operands."""
    _conv_case(case, torch.bfloat16, 2)


HALO_T_CASES = [
    # N, H, W (input), Co
    (256, 16, 16, 64),     # 648 tiles: several per CTA
    (128, 16, 16, 64),     # the training step's up4 / down0-dgrad shape
    (3, 32, 32, 64),       # tiles straddle images
    (5, 8, 8, 128),        # two output-channel tiles
    (1, 4, 12, 64),        # non-square, one partial tile
]


@pytest.mark.parametrize("case", HALO_T_CASES)
def test_conv_transposed_halo(case):
    """impl 5 on a 4x4 stride-2 pad-1 transposed gather with 64 input channels = the halo kernel of that read pattern (input
    positions of the padded flat space as the tile, four output parity classes as four accumulators), in both of its uses:
    nn.ConvTranspose2d forward (+bias) and the input gradient of the 4x4 stride-2 nn.Conv2d; against ATen on the same bf16
    operands, and against the per-tap kernel (impl 4)."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams
    N, H, W, Co = case
    Ci = 64
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N * 100 + H + Co)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    xh = ops.nchw_to_nhwc(x, torch.bfloat16)
    xq = xh.float().permute(0, 3, 1, 2)
    code = ops.dtype_code(xh)
    # (a) ConvTranspose2d: weight [Ci][Co][4][4]
    wt = (torch.randn(Ci, Co, 4, 4, generator=g) / math.sqrt(Ci * 4)).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    wk = _repack(wt, True, torch.bfloat16)
    ref = F.conv_transpose2d(xq, wt.to(torch.bfloat16).float(), bias, stride=2, padding=1)
    outs = {}
    for impl in (5, 4):
        y_full = torch.zeros(N, 2 * H, 2 * W, Co + 16, device=dev, dtype=torch.bfloat16)      # written into a channel slice
        p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y_full, 8, Co), _null(), wk.data_ptr(), 16 * Ci, 1, Ci, bias.data_ptr(), None, 0,
                       N, H, W, Ci, 2 * H, 2 * W, Co, 4, 4, 2, 1, 1, code, impl, 0, None, 0)
        ops.conv2d_raw(p)
        got = y_full[..., 8:8 + Co].float().permute(0, 3, 1, 2)
        assert rel_l2(got, ref) < TOL[torch.bfloat16], f"conv_transpose2d impl {impl}"
        assert y_full[..., :8].abs().max() == 0 and y_full[..., 8 + Co:].abs().max() == 0, "wrote outside its channel slice"
        outs[impl] = got
    assert rel_l2(outs[5], outs[4]) < 1e-3      # same products, fp32 accumulation in a different order, one bf16 rounding
    # (b) input gradient of Conv2d(Co' = 64 -> ..., k 4, s 2, p 1): dy on the small grid [N, Ci=64 channels], w [64(out)][Co(in)][4][4]
    w = (torch.randn(Ci, Co, 4, 4, generator=g) / math.sqrt(Co * 16)).to(dev)      # Conv2d weight [out = 64][in = Co]
    wkt = _repack(w, False, torch.bfloat16, dgrad=True)
    dx = torch.empty(N, 2 * H, 2 * W, Co, device=dev, dtype=torch.bfloat16)
    p2 = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(dx), _null(), wkt.data_ptr(), 16 * Ci, 1, Ci, None, None, 0,
                    N, H, W, Ci, 2 * H, 2 * W, Co, 4, 4, 2, 1, 1, code, 5, 0, None, 0)
    ops.conv2d_raw(p2)
    dref = torch.nn.grad.conv2d_input((N, Co, 2 * H, 2 * W), w.to(torch.bfloat16).float(), xq, stride=2, padding=1)
    assert rel_l2(dx.float().permute(0, 3, 1, 2), dref) < TOL[torch.bfloat16], "dgrad of the strided conv"


HALO_S_CASES = [
    # N, Ho, Wo (output), Co
    (256, 16, 16, 64),     # 648 tiles: several per CTA, three ring stages
    (128, 16, 16, 64),     # the training step's down0 / up4-dgrad shape
    (3, 32, 32, 64),       # two ring stages, tiles straddle images
    (5, 8, 8, 128),        # two output-channel tiles
    (1, 4, 12, 64),        # non-square, one partial tile
]


@pytest.mark.parametrize("case", HALO_S_CASES)
def test_conv_strided_halo(case):
    """impl 5 on a 4x4 stride-2 pad-1 convolution with 64 input channels = the halo kernel of that read pattern (the four parity
    sub-lattices of the input as four halo tiles per output tile), in both of its uses: nn.Conv2d downsampling forward (+bias) and
    the input gradient of nn.ConvTranspose2d; against ATen on the same bf16 operands, and against the per-tap kernel (impl 4)."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams
    N, Ho, Wo, Co = case
    Ci, H, W = 64, 2 * Ho, 2 * Wo
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N * 100 + Ho + Co)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    xh = ops.nchw_to_nhwc(x, torch.bfloat16)
    xq = xh.float().permute(0, 3, 1, 2)
    code = ops.dtype_code(xh)
    # (a) Conv2d: weight [Co][Ci][4][4]
    w = (torch.randn(Co, Ci, 4, 4, generator=g) / math.sqrt(Ci * 16)).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    wk = _repack(w, False, torch.bfloat16)
    ref = F.conv2d(xq, w.to(torch.bfloat16).float(), bias, stride=2, padding=1)
    outs = {}
    for impl in (5, 4):
        y_full = torch.zeros(N, Ho, Wo, Co + 16, device=dev, dtype=torch.bfloat16)      # written into a channel slice
        p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y_full, 8, Co), _null(), wk.data_ptr(), 16 * Ci, 1, Ci, bias.data_ptr(), None, 0,
                       N, H, W, Ci, Ho, Wo, Co, 4, 4, 2, 1, 0, code, impl, 0, None, 0)
        ops.conv2d_raw(p)
        got = y_full[..., 8:8 + Co].float().permute(0, 3, 1, 2)
        assert rel_l2(got, ref) < TOL[torch.bfloat16], f"conv2d impl {impl}"
        assert y_full[..., :8].abs().max() == 0 and y_full[..., 8 + Co:].abs().max() == 0, "wrote outside its channel slice"
        outs[impl] = got
    assert rel_l2(outs[5], outs[4]) < 1e-3      # same products, fp32 accumulation in a different order, one bf16 rounding
    # (b) input gradient of ConvTranspose2d(Co -> 64, k 4, s 2, p 1): dy = x on the big grid (64 channels), weight [Co(in)][64(out)][4][4]
    wt = (torch.randn(Co, Ci, 4, 4, generator=g) / math.sqrt(Ci * 4)).to(dev)
    wkt = _repack(wt, True, torch.bfloat16, dgrad=True)
    dx = torch.empty(N, Ho, Wo, Co, device=dev, dtype=torch.bfloat16)
    p2 = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(dx), _null(), wkt.data_ptr(), 16 * Ci, 1, Ci, None, None, 0,
                    N, H, W, Ci, Ho, Wo, Co, 4, 4, 2, 1, 0, code, 5, 0, None, 0)
    ops.conv2d_raw(p2)
    dref = F.conv2d(xq, wt.to(torch.bfloat16).float(), stride=2, padding=1)
    assert rel_l2(dx.float().permute(0, 3, 1, 2), dref) < TOL[torch.bfloat16], "dgrad of the transposed conv"


WGRAD_HALO_CASES = [
    (128, 32, 32, 64, 64),     # the bench layer: 1156 position tiles, both tap groups, 85 + 63 CTAs
    (3, 32, 32, 128, 64),      # two 64-channel chunks of q
    (5, 16, 16, 64, 128),      # two 64-channel tiles of p
    (2, 64, 64, 64, 64),       # 64x64: three padded rows per p tile
    (7, 8, 8, 192, 64),        # tiles straddle images, three chunks
    (1, 8, 24, 64, 64),        # non-square, fewer tiles than CTAs asked for
]


@pytest.mark.parametrize("case", WGRAD_HALO_CASES)
def test_wgrad_halo(case):
    """impl 5 = the halo weight-gradient kernel (3x3 stride 1: padded-flat position tiles, taps as shifted MN-major windows of one
    shared-memory tile, two taps per M = 128 instruction) against ATen's convolution_backward on the same bf16 operands, written
    through the engine's staging layout [a][r][s][b] and accumulated onto existing content; the per-tap kernel (impl 2 with the
    halo kernel's size threshold not met) must agree with it."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import WgradParams
    N, H, W, Ci, Co = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + H + Ci)
    xh = torch.randn(N, H, W, Ci, generator=g).to(dev).to(torch.bfloat16)
    dyh = torch.randn(N, H, W, Co, generator=g).to(dev).to(torch.bfloat16)
    wref = torch.nn.grad.conv2d_weight(xh.float().permute(0, 3, 1, 2), (Co, Ci, 3, 3), dyh.float().permute(0, 3, 1, 2), stride=1, padding=1)
    for impl in (5, 2):
        dw = torch.ones(Co, 3, 3, Ci, device=dev)
        db = torch.ones(Co, device=dev)
        ops.wgrad_raw(WgradParams(ops.t4_nhwc(dyh), ops.t4_nhwc(xh), dw.data_ptr(), 9 * Ci, 1, Ci, db.data_ptr(), N, H, W, Co, H, W, Ci, 3, 3, 1, 1, impl))
        assert rel_l2((dw - 1).permute(0, 3, 1, 2), wref) < 5e-5, f"wgrad impl {impl}"
        assert rel_l2(db - 1, dyh.float().sum(dim=(0, 1, 2))) < 5e-5, f"dbias impl {impl}"


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_and_head_layouts(dtype, impl):
    """3-channel NCHW fp32 boundary tensors: stem fprop/wgrad, head fprop (small-N kernel)/dgrad/wgrad."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, WgradParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    N, H, W, Cm = 3, 32, 32, 64
    x = torch.randn(N, 3, H, W, generator=g).to(dev)
    w = (torch.randn(Cm, 3, 3, 3, generator=g) / 5).to(dev)
    b = torch.randn(Cm, generator=g).to(dev)
    wk = _repack(w, False, dtype)
    code = ops.dtype_code(wk)
    y = torch.empty(N, H, W, Cm, device=dev, dtype=dtype)
    ops.conv2d_raw(ConvParams(ops.t4_nchw(x), ops.t4_nhwc(y), _null(), wk.data_ptr(), 27, 1, 3, b.data_ptr(), None, 0,
                              N, H, W, 3, H, W, Cm, 3, 3, 1, 1, 0, code, impl, 0))
    wq = w.to(dtype).float()
    assert rel_l2(y.float().permute(0, 3, 1, 2), F.conv2d(x, wq, b, padding=1)) < TOL[dtype]
    dy = torch.randn(N, Cm, H, W, generator=g).to(dev)
    dyh = ops.nchw_to_nhwc(dy, dtype)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.wgrad_raw(WgradParams(ops.t4_nhwc(dyh), ops.t4_nchw(x), dw.data_ptr(), 27, 9, 1, db.data_ptr(), N, H, W, Cm, H, W, 3, 3, 3, 1, 1, impl))
    assert rel_l2(dw, torch.nn.grad.conv2d_weight(x, w.shape, dyh.float().permute(0, 3, 1, 2), padding=1)) < 5e-5
    assert rel_l2(db, dyh.float().sum(dim=(0, 1, 2))) < 5e-5
    # head: Cm -> 3, NHWC in, NCHW fp32 out
    a = torch.randn(N, Cm, H, W, generator=g).to(dev)
    ah = ops.nchw_to_nhwc(a, dtype)
    wh = (torch.randn(3, Cm, 3, 3, generator=g) / 24).to(dev)
    bh = torch.randn(3, generator=g).to(dev)
    whk = _repack(wh, False, dtype)
    out = torch.empty(N, 3, H, W, device=dev)
    ops.conv2d_raw(ConvParams(ops.t4_nhwc(ah), ops.t4_nchw(out), _null(), whk.data_ptr(), 9 * Cm, 1, Cm, bh.data_ptr(), None, 0,
                              N, H, W, Cm, H, W, 3, 3, 3, 1, 1, 0, code, impl, 0))
    aq, whq = ah.float().permute(0, 3, 1, 2), wh.to(dtype).float()
    assert rel_l2(out, F.conv2d(aq, whq, bh, padding=1)) < TOL[dtype]
    dout = torch.randn(N, 3, H, W, generator=g).to(dev)
    da = torch.empty(N, H, W, Cm, device=dev, dtype=dtype)
    ops.conv2d_raw(ConvParams(ops.t4_nchw(dout), ops.t4_nhwc(da), _null(), whk.data_ptr(), 1, 9 * Cm, Cm, None, None, 0,
                              N, H, W, 3, H, W, Cm, 3, 3, 1, 1, 1, code, impl, 0))
    assert rel_l2(da.float().permute(0, 3, 1, 2), torch.nn.grad.conv2d_input(a.shape, whq, dout, padding=1)) < TOL[dtype]
    dwh, dbh = torch.zeros_like(wh), torch.zeros_like(bh)
    ops.wgrad_raw(WgradParams(ops.t4_nchw(dout), ops.t4_nhwc(ah), dwh.data_ptr(), Cm * 9, 9, 1, dbh.data_ptr(), N, H, W, 3, H, W, Cm, 3, 3, 1, 1, impl))
    assert rel_l2(dwh, torch.nn.grad.conv2d_weight(aq, wh.shape, dout, padding=1)) < 5e-5
    assert rel_l2(dbh, dout.sum(dim=(0, 2, 3))) < 5e-5


@pytest.mark.parametrize("N,H,W,Cm", [(3, 32, 32, 64), (2, 8, 24, 128), (1, 5, 7, 64), (128, 32, 32, 64)])
def test_stem_and_head_dgrad_tensor_core(N, H, W, Cm):
    """conv_stem.cu: the tcgen05 im2col kernel for the 3-channel NCHW fp32 boundary tensors (stem fprop, head dgrad), incl.
    pixel counts that are not a multiple of the 128-pixel tile."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 1000 + H)
    dtype = torch.bfloat16
    x = torch.randn(N, 3, H, W, generator=g).to(dev)
    w = (torch.randn(Cm, 3, 3, 3, generator=g) / 5).to(dev)
    b = torch.randn(Cm, generator=g).to(dev)
    wk = _repack(w, False, dtype)
    code = ops.dtype_code(wk)
    y = torch.full((N, H, W, Cm), float("nan"), device=dev, dtype=dtype)
    ops.conv2d_raw(ConvParams(ops.t4_nchw(x), ops.t4_nhwc(y), _null(), wk.data_ptr(), 27, 1, 3, b.data_ptr(), None, 0,
                              N, H, W, 3, H, W, Cm, 3, 3, 1, 1, 0, code, 2, 0))
    xq, wq = x.to(dtype).float(), w.to(dtype).float()
    assert rel_l2(y.float().permute(0, 3, 1, 2), F.conv2d(xq, wq, b, padding=1)) < TOL[dtype]
    # head dgrad: dout [N,3,H,W] fp32 NCHW -> da [N,H,W,Cm], filters read from the fprop cache [3][9][Cm] through strides
    wh = (torch.randn(3, Cm, 3, 3, generator=g) / 24).to(dev)
    whk = _repack(wh, False, dtype)
    dout = torch.randn(N, 3, H, W, generator=g).to(dev)
    da = torch.full((N, H, W, Cm), float("nan"), device=dev, dtype=dtype)
    ops.conv2d_raw(ConvParams(ops.t4_nchw(dout), ops.t4_nhwc(da), _null(), whk.data_ptr(), 1, 9 * Cm, Cm, None, None, 0,
                              N, H, W, 3, H, W, Cm, 3, 3, 1, 1, 1, code, 2, 0))
    ref = torch.nn.grad.conv2d_input((N, Cm, H, W), wh.to(dtype).float(), dout.to(dtype).float(), padding=1)
    assert rel_l2(da.float().permute(0, 3, 1, 2), ref) < TOL[dtype]


@pytest.mark.parametrize("N,H,W,Ci,Co,silu,res", [(8, 32, 32, 64, 64, 1, True), (6, 16, 16, 128, 64, 1, False), (5, 8, 24, 64, 128, 0, False),
                                                   (70, 32, 32, 64, 64, 1, False), (3, 64, 64, 64, 64, 1, True)])
def test_conv_with_fused_groupnorm(N, H, W, Ci, Co, silu, res):
    """Halo kernel with GroupNorm(+SiLU) applied to its operand tile in shared memory (dmu_conv_params.gn_coef): y and the
    side output a = act(GN(x)) against ATen on the same bf16-rounded operands; padding pixels must stay zero."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, GnParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 100 + H + Ci)
    dtype, G = torch.bfloat16, 32
    x = (torch.randn(N, Ci, H, W, generator=g) * 1.7 + 0.4).to(dev)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / math.sqrt(9 * Ci)).to(dev)
    b = torch.randn(Co, generator=g).to(dev)
    gamma, beta = (1 + 0.2 * torch.randn(Ci, generator=g)).to(dev), (0.1 * torch.randn(Ci, generator=g)).to(dev)
    xh = ops.nchw_to_nhwc(x, dtype)
    wk = _repack(w, False, dtype)
    lib = _abi.lib()
    sums = torch.zeros(N, G, 2, device=dev)
    coef = torch.empty(N, Ci, 2, device=dev)
    pg = GnParams(ops.t4_nhwc(xh), _null(), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  None, None, None, N, H, W, Ci, G, silu, 1e-5, 0)
    _abi.check(lib.dmu_gn_stats(C.byref(pg), _stream()))
    _abi.check(lib.dmu_gn_coef(C.byref(pg), coef.data_ptr(), _stream()))
    y = torch.full((N, H, W, Co), float("nan"), device=dev, dtype=dtype)
    a = torch.full((N, H, W, Ci), float("nan"), device=dev, dtype=dtype)
    r = torch.randn(N, H, W, Co, generator=g).to(dev).to(dtype) if res else None
    p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y), ops.t4_nhwc(r) if res else _null(), wk.data_ptr(), 9 * Ci, 1, Ci, b.data_ptr(), None, 0,
                   N, H, W, Ci, H, W, Co, 3, 3, 1, 1, 0, ops.dtype_code(wk), 5, 0, None, 0, coef.data_ptr(), silu, 0, ops.t4_nhwc(a))
    assert lib.dmu_conv2d_gn_supported(C.byref(p)) == 1
    ops.conv2d_raw(p)
    xq = xh.float().permute(0, 3, 1, 2)
    aref = F.group_norm(xq, G, gamma, beta, eps=1e-5)
    aref = F.silu(aref) if silu else aref
    assert rel_l2(a.float().permute(0, 3, 1, 2), aref) < TOL[dtype]
    yref = F.conv2d(aref.to(dtype).float(), w.to(dtype).float(), b, padding=1)
    if res:
        yref = yref + r.float().permute(0, 3, 1, 2)
    assert rel_l2(y.float().permute(0, 3, 1, 2), yref) < TOL[dtype]


@pytest.mark.parametrize("M,I,O", [(128, 256, 3136), (7, 64, 256), (2048, 128, 384), (5, 1, 64)])
def test_linear_via_conv(M, I, O):
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, WgradParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M + O)
    x, w, b = torch.randn(M, I, generator=g).to(dev), (torch.randn(O, I, generator=g) / math.sqrt(I)).to(dev), torch.randn(O, generator=g).to(dev)
    y = torch.empty(M, O, device=dev)
    ops.conv2d_raw(ConvParams(ops.t4_rows(x), ops.t4_rows(y), _null(), w.data_ptr(), I, 1, 0, b.data_ptr(), None, 0,
                              M, 1, 1, I, 1, 1, O, 1, 1, 1, 0, 0, 0, 1, 0))
    assert rel_l2(y, F.linear(x, w, b)) < 2e-5
    dy = torch.randn(M, O, generator=g).to(dev)
    dx = torch.empty(M, I, device=dev)
    ops.conv2d_raw(ConvParams(ops.t4_rows(dy), ops.t4_rows(dx), _null(), w.data_ptr(), 1, I, 0, None, None, 0,
                              M, 1, 1, O, 1, 1, I, 1, 1, 1, 0, 0, 0, 1, 0))
    assert rel_l2(dx, dy @ w) < 2e-5
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    ops.wgrad_raw(WgradParams(ops.t4_rows(dy), ops.t4_rows(x), dw.data_ptr(), I, 1, 0, db.data_ptr(), M, 1, 1, O, 1, 1, I, 1, 1, 1, 0, 1))
    assert rel_l2(dw, dy.t() @ x) < 5e-5 and rel_l2(db, dy.sum(0)) < 5e-5


@pytest.mark.parametrize("C_,G", [(64, 32), (128, 32), (192, 32), (256, 32), (384, 32), (512, 32), (32, 32), (16, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("silu", [1, 0])
@pytest.mark.parametrize("one_call", [False, True])
def test_groupnorm_fwd_bwd(C_, G, dtype, silu, one_call):
    N, H, W = (3, 8, 8) if C_ > 128 else (2, 16, 16)
    _groupnorm_case(N, H, W, C_, G, dtype, silu, one_call)


@pytest.mark.parametrize("shape", [(3, 32, 32, 64), (2, 32, 32, 128), (5, 1, 1, 256), (4, 2, 2, 384), (3, 4, 4, 128), (2, 64, 64, 64), (2, 12, 20, 48)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_groupnorm_single_pass_cluster_sizes(shape, dtype):
    """dmu_gn_forward / dmu_gn_backward over the image sizes of the UNet: 1-, 2-, 4- and 8-CTA clusters per image, and the
    two-pass fallback (64x64x64 does not fit eight CTAs' registers)."""
    N, H, W, C_ = shape
    _groupnorm_case(N, H, W, C_, 16 if C_ % 32 else 32, dtype, 1, True)


def _groupnorm_case(N, H, W, C_, G, dtype, silu, one_call):
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import GnParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C_ + silu)
    x = (torch.randn(N, C_, H, W, generator=g) * 2 + 0.5).to(dev)
    gamma, beta = (1 + 0.2 * torch.randn(C_, generator=g)).to(dev), (0.1 * torch.randn(C_, generator=g)).to(dev)
    xh = ops.nchw_to_nhwc(x, dtype)
    xq = xh.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yh = torch.empty_like(xh)
    sums = torch.zeros(N, G, 2, device=dev)
    lib = _abi.lib()
    p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(yh), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                 None, None, None, N, H, W, C_, G, silu, 1e-5, 0)
    if one_call:
        _abi.check(lib.dmu_gn_forward(C.byref(p), _stream()))
    else:
        _abi.check(lib.dmu_gn_stats(C.byref(p), _stream()))
        _abi.check(lib.dmu_gn_apply(C.byref(p), _stream()))
    gq, bq = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.group_norm(xq, G, gq, bq, eps=1e-5)
    ref = F.silu(ref) if silu else ref
    assert rel_l2(yh.float().permute(0, 3, 1, 2), ref) < TOL[dtype]
    dy = torch.randn(N, C_, H, W, generator=g).to(dev)
    dyh = ops.nchw_to_nhwc(dy, dtype)
    add = torch.randn(N, H, W, C_, generator=g).to(dev).to(dtype)
    dxh = torch.empty_like(xh)
    red = torch.zeros(N, C_, 2, device=dev)
    dgam, dbet = torch.zeros(C_, device=dev), torch.zeros(C_, device=dev)
    p2 = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), ops.t4_nhwc(dxh), ops.t4_nhwc(add), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  red.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), N, H, W, C_, G, silu, 1e-5, 0)
    if one_call:
        _abi.check(lib.dmu_gn_backward(C.byref(p2), _stream()))
    else:
        _abi.check(lib.dmu_gn_bwd_reduce(C.byref(p2), _stream()))
        _abi.check(lib.dmu_gn_bwd_apply(C.byref(p2), _stream()))
    ref.backward(dyh.float().permute(0, 3, 1, 2))
    tol = 3e-5 if dtype == torch.float32 else 8e-3
    assert rel_l2(dxh.float().permute(0, 3, 1, 2), xq.grad + add.float().permute(0, 3, 1, 2)) < tol
    assert rel_l2(dgam, gq.grad) < 1e-4 and rel_l2(dbet, bq.grad) < 1e-4


@pytest.mark.parametrize("N,S,C_,heads", [(3, 16, 128, 4), (2, 64, 128, 4), (5, 1, 256, 4), (4, 4, 128, 4), (2, 16, 64, 4), (3, 4, 32, 4),
                                          (9, 64, 128, 4), (130, 1, 256, 4), (64, 4, 256, 4), (128, 16, 128, 4), (1, 128, 64, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_core(N, S, C_, heads, dtype):
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import AttnParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(S * C_)
    qkv = torch.randn(N, S, 3 * C_, generator=g).to(dev).to(dtype)
    o = torch.empty(N, S, C_, device=dev, dtype=dtype)
    lse = torch.empty(N, heads, S, device=dev)
    do = torch.randn(N, S, C_, generator=g).to(dev).to(dtype)
    dqkv = torch.empty_like(qkv)
    p = AttnParams(qkv.data_ptr(), 3 * C_, o.data_ptr(), C_, do.data_ptr(), C_, dqkv.data_ptr(), 3 * C_, lse.data_ptr(), N, S, C_, heads, ops.dtype_code(qkv), 0)
    lib = _abi.lib()
    _abi.check(lib.dmu_attn_fwd(C.byref(p), _stream()))
    _abi.check(lib.dmu_attn_bwd(C.byref(p), _stream()))
    d = C_ // heads
    qr = qkv.float().clone().requires_grad_(True)
    q, k, v = [z.reshape(N, S, heads, d).transpose(1, 2) for z in qr.split(C_, dim=-1)]
    ref = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5, dim=-1), v).transpose(1, 2).reshape(N, S, C_)
    assert rel_l2(o, ref) < TOL[dtype]
    ref.backward(do.float())
    assert rel_l2(dqkv, qr.grad) < (5e-5 if dtype == torch.float32 else 1.5e-2)


def test_small_fp32_ops():
    ops, _abi = _mods()
    from oracle.unet import sinusoidal_embedding
    lib = _abi.lib()
    dev = torch.device("cuda:0")
    for dim in (32, 64):
        t = torch.tensor([0, 1, 17, 500, 999], device=dev)
        emb = torch.empty(5, dim, device=dev)
        _abi.check(lib.dmu_sinusoidal_embedding(t.data_ptr(), 0, emb.data_ptr(), 5, dim, _stream()))
        ref = sinusoidal_embedding(t.cpu(), dim)
        assert (emb.cpu() - ref).abs().max() < 1.5e-4   # |arg| up to 1e3 rad: 1 ulp of the frequency (CPU vs GPU expf) is 6e-5 rad
        tf = t.float()
        _abi.check(lib.dmu_sinusoidal_embedding(tf.data_ptr(), 1, emb.data_ptr(), 5, dim, _stream()))
        assert (emb.cpu() - ref).abs().max() < 1.5e-4
    x = torch.randn(1000, device=dev) * 3
    y, dx = torch.empty_like(x), torch.empty_like(x)
    dy = torch.randn_like(x)
    for kind, fn in ((0, F.gelu), (1, F.silu), (2, torch.log)):
        xi = x.abs() + 0.1 if kind == 2 else x
        xr = xi.clone().requires_grad_(True)
        _abi.check(lib.dmu_act_fwd(xi.data_ptr(), y.data_ptr(), 1000, kind, _stream()))
        _abi.check(lib.dmu_act_bwd(xi.data_ptr(), dy.data_ptr(), dx.data_ptr(), 1000, kind, _stream()))
        r = fn(xr)
        r.backward(dy)
        assert rel_l2(y, r) < 1e-6 and rel_l2(dx, xr.grad) < 1e-6, kind
    # colsum
    from diffusion_model_universal_b200 import ops
    for dtype in (torch.float32, torch.bfloat16):
        z = torch.randn(3, 8, 8, 192, device=dev).to(dtype)
        nc = torch.zeros(3, 200, device=dev)
        c = torch.ones(192, device=dev)
        t4 = ops.t4_nhwc(z)
        _abi.check(lib.dmu_colsum(C.byref(t4), 3, 8, 8, 192, nc.data_ptr(), 200, c.data_ptr(), 0.5, _stream()))
        ref = z.float().sum(dim=(1, 2)) * 0.5
        assert rel_l2(nc[:, :192], ref) < 1e-5 and rel_l2(c - 1, ref.sum(0)) < 1e-5


def test_adam_ema_matches_torch():
    ops, _abi = _mods()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    n = 10007
    p0 = torch.randn(n, generator=g).to(dev)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=2e-4, betas=(0.9, 0.999), eps=1e-8)
    p, m, v, ema = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev), p0.clone()
    ema_ref = p0.clone()
    # the second copy takes the step count from device memory and is updated in two ranges (what the overlapped step does)
    p2, m2, v2, ema2 = p0.clone(), torch.zeros(n + 1, device=dev)[:n], torch.zeros(n, device=dev), p0.clone()
    step_dev = torch.zeros((), device=dev, dtype=torch.int64)
    cut = 4096
    for step in range(1, 4):
        grad = torch.randn(n, generator=g).to(dev)
        p_ref.grad = grad.clone()
        opt.step()
        ema_ref.mul_(0.999).add_(p_ref.detach(), alpha=0.001)
        _abi.check(_abi.lib().dmu_adam_ema(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), ema.data_ptr(), n,
                                           2e-4, 0.9, 0.999, 1e-8, 0.0, step, 0.999, 1.0, None, _stream()))
        step_dev.add_(1)
        for lo, hi in ((cut, n), (0, cut)):
            _abi.check(_abi.lib().dmu_adam_ema(p2.data_ptr() + 4 * lo, grad.data_ptr() + 4 * lo, m2.data_ptr() + 4 * lo, v2.data_ptr() + 4 * lo,
                                               ema2.data_ptr() + 4 * lo, hi - lo, 2e-4, 0.9, 0.999, 1e-8, 0.0, 0, 0.999, 1.0,
                                               step_dev.data_ptr(), _stream()))
    assert rel_l2(p, p_ref) < 1e-6 and rel_l2(ema, ema_ref) < 1e-6
    assert rel_l2(p2, p_ref) < 1e-6 and rel_l2(ema2, ema_ref) < 1e-6 and rel_l2(m2, m) < 1e-6


@pytest.mark.parametrize("shape", [(3, 8, 8, 64, 8), (2, 16, 16, 128, 8), (2, 4, 4, 16, 8), (2, 8, 8, 64, 32)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("silu", [1, 0])
def test_groupnorm_silu_double_backward(shape, dtype, silu):
    """dmu_gn_bwd_bwd (derivative of the GroupNorm+SiLU backward) against torch's double backward."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import GnParams, GnBwd2Params
    N, H, W, C_, G = shape
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C_ + silu)
    x = (torch.randn(N, C_, H, W, generator=g) * 1.5 + 0.3).to(dev)
    dy = torch.randn(N, C_, H, W, generator=g).to(dev)
    cc = torch.randn(N, C_, H, W, generator=g).to(dev)
    gamma, beta = (1 + 0.2 * torch.randn(C_, generator=g)).to(dev), (0.1 * torch.randn(C_, generator=g)).to(dev)
    xh, dyh, ch = ops.nchw_to_nhwc(x, dtype), ops.nchw_to_nhwc(dy, dtype), ops.nchw_to_nhwc(cc, dtype)
    lib = _abi.lib()
    sums = torch.zeros(N, G, 2, device=dev)
    yh = torch.empty_like(xh)
    p = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(yh), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                 None, None, None, N, H, W, C_, G, silu, 1e-5, 0)
    _abi.check(lib.dmu_gn_forward(C.byref(p), _stream()))
    gx, gdy = torch.empty_like(xh), torch.empty_like(xh)
    dgam, dbet = torch.zeros(C_, device=dev), torch.zeros(C_, device=dev)
    p2 = GnBwd2Params(ops.t4_nhwc(xh), ops.t4_nhwc(dyh), ops.t4_nhwc(ch), ops.t4_nhwc(gx), ops.t4_nhwc(gdy), sums.data_ptr(), gamma.data_ptr(),
                      beta.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), N, H, W, C_, G, silu, 1e-5, 0)
    _abi.check(lib.dmu_gn_bwd_bwd(C.byref(p2), _stream()))
    xq = xh.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    dq = dyh.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    cq = ch.float().permute(0, 3, 1, 2)
    gq, bq = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xq, G, gq, bq, eps=1e-5)
    y = F.silu(y) if silu else y
    (dx,) = torch.autograd.grad(y, xq, dq, create_graph=True)
    rx, rd, rg, rb = torch.autograd.grad((cq * dx).sum(), [xq, dq, gq, bq], allow_unused=True)
    tol = 1e-4 if dtype == torch.float32 else 1.5e-2
    assert rel_l2(gx.float().permute(0, 3, 1, 2), rx) < tol
    assert rel_l2(gdy.float().permute(0, 3, 1, 2), rd) < tol
    assert rel_l2(dgam, rg) < 1e-3
    if silu:
        assert rel_l2(dbet, rb) < 1e-3
    else:
        assert dbet.abs().max() < 1e-3 * max(1.0, float(dgam.abs().max()))


def test_head_conv_on_halo_kernel():
    """C -> 3 head conv at bench size: impl 0 routes it to the persistent halo kernel (weight-box rows past Cj are TMA
    zeros, fp32 NCHW epilogue); compare with ATen and with the direct SIMT kernel (impl 3)."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    N, H, W, Cm = 80, 32, 32, 64
    a = torch.randn(N, Cm, H, W, generator=g).to(dev)
    ah = ops.nchw_to_nhwc(a, torch.bfloat16)
    wh = (torch.randn(3, Cm, 3, 3, generator=g) / 24).to(dev)
    bh = torch.randn(3, generator=g).to(dev)
    whk = _repack(wh, False, torch.bfloat16)
    outs = []
    for impl in (0, 3):
        out = torch.zeros(N, 3, H, W, device=dev)
        ops.conv2d_raw(ConvParams(ops.t4_nhwc(ah), ops.t4_nchw(out), _null(), whk.data_ptr(), 9 * Cm, 1, Cm, bh.data_ptr(), None, 0,
                                  N, H, W, Cm, H, W, 3, 3, 3, 1, 1, 0, 1, impl, 0))
        outs.append(out)
    ref = F.conv2d(ah.float().permute(0, 3, 1, 2), wh.to(torch.bfloat16).float(), bh, padding=1)
    assert rel_l2(outs[0], ref) < 1e-5 and rel_l2(outs[1], ref) < 1e-5


# ------------------------------------------------------------------------------------------------ GroupNorm in the conv epilogue
GN_EPI_CASES = [
    # N, H, W, Ci, Co, R, stride, G of the norm over the conv's OUTPUT channels, has_res
    (6, 8, 8, 64, 128, 3, 1, 32, True),       # 2 images per 128-pixel tile, cpg 4
    (128, 8, 8, 128, 128, 3, 1, 32, False),   # the bench's 8x8 stage
    (5, 8, 8, 64, 64, 3, 1, 32, True),        # cpg 2, batch not a multiple of the images per tile
    (37, 4, 4, 128, 128, 3, 1, 32, True),     # 8 images per tile, ragged batch
    (128, 2, 2, 256, 256, 3, 1, 32, True),    # cpg 8, 32 images per tile
    (128, 1, 1, 256, 256, 3, 1, 32, False),   # one pixel per image: 8 of 9 taps dead
    (16, 16, 16, 64, 128, 4, 2, 32, False),   # the 4x4 stride-2 downsample feeding the next stage's norm1 (output 8x8)
    (9, 4, 4, 128, 128, 1, 1, 32, True),      # attention final projection (+ x) followed by its post-norm
    (8, 1, 1, 256, 512, 3, 1, 32, False),     # cpg 16
    (4, 8, 8, 64, 64, 3, 1, 2, False),        # cpg 32
    (128, 16, 16, 64, 64, 3, 1, 32, True),    # the bench's 16x16 stage: one 256-pixel box (two accumulators) per image, cpg 2
    (5, 16, 16, 64, 128, 3, 1, 32, True),     # the same with two channel tiles, cpg 4
    (3, 16, 16, 128, 256, 3, 1, 32, False),   # cpg 8
]


def _gn_epi_setup(case, g, dev):
    ops, _abi = _mods()
    N, H, W, Ci, Co, R, stride, G, has_res = case
    pad = 1 if R > 1 else 0
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    w = (torch.randn(Co, Ci, R, R, generator=g) / math.sqrt(Ci * R * R)).to(dev)
    return N, H, W, Ci, Co, R, stride, pad, G, has_res, Ho, Wo, x, w


@pytest.mark.parametrize("silu", [1, 0])
@pytest.mark.parametrize("case", GN_EPI_CASES)
def test_conv_gn_epilogue_forward(case, silu):
    """dmu_conv_params.gn_fuse mode 1: y = conv(x) + bias + temb + res and a = act(GroupNorm(y)) from ONE launch, against ATen
    on the bf16-rounded operands and against this library's own conv launch followed by dmu_gn_forward."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, GnParams
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(sum(case[:8]))
    N, H, W, Ci, Co, R, stride, pad, G, has_res, Ho, Wo, x, w = _gn_epi_setup(case, g, dev)
    bias = torch.randn(Co, generator=g).to(dev)
    temb = torch.randn(N, Co, generator=g).to(dev)
    res = torch.randn(N, Ho, Wo, Co, generator=g).to(dev).to(dtype) if has_res else None
    gamma, beta = (1 + 0.2 * torch.randn(Co, generator=g)).to(dev), (0.1 * torch.randn(Co, generator=g)).to(dev)
    xh = ops.nchw_to_nhwc(x, dtype)
    wk = _repack(w, False, dtype)
    code = ops.dtype_code(xh)
    lib = _abi.lib()

    def run(fused):
        y = torch.zeros(N, Ho, Wo, Co, device=dev, dtype=dtype)
        a = torch.zeros(N, Ho, Wo, Co + 8, device=dev, dtype=dtype)       # normalised output into a channel slice
        sums = torch.full((N, G, 2), float("nan"), device=dev) if fused else torch.zeros(N, G, 2, device=dev)
        gp = GnParams(ops.t4_nhwc(y), ops.t4_nhwc(a, 8, Co), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      None, None, None, N, Ho, Wo, Co, G, silu, 1e-5, 0)
        p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y), ops.t4_nhwc(res) if has_res else _null(), wk.data_ptr(), R * R * Ci, 1, Ci,
                       bias.data_ptr(), temb.data_ptr(), Co, N, H, W, Ci, Ho, Wo, Co, R, R, stride, pad, 0, code, 0, 0, None, 0)
        if fused:
            p.gn_fuse = C.cast(C.pointer(gp), C.c_void_p)
            p.gn_fuse_mode = 1
            tiles = lib.dmu_conv2d_gn_fuse_supported(C.byref(p))
            assert tiles > 0, "the library refused a shape the plan builder relies on"
        ops.conv2d_raw(p)
        if not fused:
            _abi.check(lib.dmu_gn_forward(C.byref(gp), _stream()))
        torch.cuda.synchronize()
        assert a[..., :8].abs().max() == 0
        return y, a[..., 8:], sums
    y1, a1, s1 = run(True)
    y0, a0, s0 = run(False)
    ref = F.conv2d(xh.float().permute(0, 3, 1, 2), w.to(dtype).float(), bias, stride=stride, padding=pad) + temb[:, :, None, None]
    if has_res:
        ref = ref + res.float().permute(0, 3, 1, 2)
    assert rel_l2(y1.float().permute(0, 3, 1, 2), ref) < TOL[dtype], "conv output"
    yq = y1.float().permute(0, 3, 1, 2)           # the norm sees the stored (bf16) tensor
    aref = F.group_norm(yq, G, gamma, beta, eps=1e-5)
    aref = F.silu(aref) if silu else aref
    assert rel_l2(a1.float().permute(0, 3, 1, 2), aref) < TOL[dtype], "normalised output vs ATen"
    assert rel_l2(y1, y0) < 1e-3, "the fused launch must store the same conv output"
    assert rel_l2(a1, a0) < 2e-3, "fused vs stand-alone GroupNorm"
    assert rel_l2(s1, s0) < 1e-5, "raw sums handed to the backward"


@pytest.mark.parametrize("silu", [1, 0])
@pytest.mark.parametrize("adds", [0, 2])
@pytest.mark.parametrize("case", [c for c in GN_EPI_CASES if c[5] != 4])
def test_conv_gn_epilogue_backward(case, adds, silu):
    """dmu_conv_params.gn_fuse mode 2: a layer y = conv(act(GroupNorm(x))).  One launch computes dgrad(dy) and, in its epilogue,
    the GroupNorm(+SiLU) backward (+ addends), against autograd of the ATen composition and against this library's own
    dgrad launch followed by dmu_gn_backward; the per-tile channel sums fold to dgamma / dbeta."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, GnParams, GnPgDesc
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(sum(case[:8]) + 1)
    # here Ci is the channel count of the NORM (the conv's input), Co the conv's output; the dgrad maps Co -> Ci
    N, H, W, Co, Ci, R, stride, pad, G, _, _, _, _, _ = _gn_epi_setup(case, g, dev)
    x = (torch.randn(N, Ci, H, W, generator=g) * 1.5 + 0.3).to(dev)
    w = (torch.randn(Co, Ci, R, R, generator=g) / math.sqrt(Ci * R * R)).to(dev)
    gamma, beta = (1 + 0.2 * torch.randn(Ci, generator=g)).to(dev), (0.1 * torch.randn(Ci, generator=g)).to(dev)
    dy = torch.randn(N, Co, H, W, generator=g).to(dev)
    add_t = [torch.randn(N, H, W, Ci, generator=g).to(dev).to(dtype) for _ in range(adds)]
    xh, dyh = ops.nchw_to_nhwc(x, dtype), ops.nchw_to_nhwc(dy, dtype)
    wkt = _repack(w, False, dtype, dgrad=True)
    code = ops.dtype_code(xh)
    lib = _abi.lib()
    # forward statistics from the library's own norm
    ah = torch.empty_like(xh)
    sums = torch.zeros(N, G, 2, device=dev)
    gf = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(ah), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  None, None, None, N, H, W, Ci, G, silu, 1e-5, 0)
    _abi.check(lib.dmu_gn_forward(C.byref(gf), _stream()))

    def run(fused):
        da = torch.zeros(N, H, W, Ci, device=dev, dtype=dtype)
        dx = torch.zeros(N, H, W, Ci, device=dev, dtype=dtype)
        red = torch.zeros(N, Ci, 2, device=dev)
        dgam, dbet = torch.zeros(Ci, device=dev), torch.zeros(Ci, device=dev)
        t4s = [ops.t4_nhwc(t) for t in add_t] + [_null(), _null()]
        gp = GnParams(ops.t4_nhwc(xh), ops.t4_nhwc(da), ops.t4_nhwc(dx), t4s[0], t4s[1], sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      red.data_ptr(), None, None, N, H, W, Ci, G, silu, 1e-5, 0)
        p = ConvParams(ops.t4_nhwc(dyh), ops.t4_nhwc(da), _null(), wkt.data_ptr(), R * R * Co, 1, Co, None, None, 0,
                       N, H, W, Co, H, W, Ci, R, R, 1, pad, 1, code, 0, 0, None, 0)
        tiles = 0
        if fused:
            p.gn_fuse = C.cast(C.pointer(gp), C.c_void_p)
            p.gn_fuse_mode = 2
            tiles = lib.dmu_conv2d_gn_fuse_supported(C.byref(p))
            assert tiles > 0, "the library refused a shape the plan builder relies on"
        ops.conv2d_raw(p)
        if not fused:
            _abi.check(lib.dmu_gn_backward(C.byref(gp), _stream()))
        d = GnPgDesc(red.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), Ci, tiles)
        tab = torch.frombuffer(bytearray(bytes(d)), dtype=torch.uint8).cuda()
        _abi.check(lib.dmu_gn_param_grads(tab.data_ptr(), 1, Ci, N, _stream()))
        torch.cuda.synchronize()
        return dx, dgam, dbet
    dx1, dg1, db1 = run(True)
    dx0, dg0, db0 = run(False)
    xq = xh.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gq, bq = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a = F.group_norm(xq, G, gq, bq, eps=1e-5)
    a = F.silu(a) if silu else a
    y = F.conv2d(a, w.to(dtype).float(), padding=pad)
    y.backward(dyh.float().permute(0, 3, 1, 2))
    want = xq.grad + sum(t.float().permute(0, 3, 1, 2) for t in add_t) if adds else xq.grad
    assert rel_l2(dx1.float().permute(0, 3, 1, 2), want) < 1.2e-2, "dx vs autograd (bf16 dy rounding included)"
    assert rel_l2(dx1, dx0) < 4e-3, "fused vs dgrad + stand-alone GroupNorm backward"
    assert rel_l2(dg1, gq.grad) < 1e-2 and rel_l2(db1, bq.grad) < 1e-2
    # the fused path forms du from the fp32 accumulator, the stand-alone norm from the bf16-rounded dgrad output: 2e-3 apart at most
    assert rel_l2(dg1, dg0) < 4e-3 and rel_l2(db1, db0) < 4e-3


# ------------------------------------------------------------------------------------------------ GroupNorm statistics in the halo conv's epilogue
@pytest.mark.parametrize("case", [(128, 32, 32, 64, 64, True, True), (32, 64, 64, 64, 64, False, True), (128, 32, 32, 64, 128, True, False),
                                  (129, 32, 32, 64, 64, True, False)])
def test_conv_gn_statistics_epilogue(case):
    """dmu_conv_params.gn_fuse mode 3: the persistent 3x3 kernel adds the raw GroupNorm sums of its stored output to gn->sums; the
    conv output is unchanged and dmu_gn_apply on those sums equals dmu_gn_forward (statistics pass + apply) and ATen."""
    ops, _abi = _mods()
    from diffusion_model_universal_b200._abi import ConvParams, GnParams
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16
    N, H, W, Ci, Co, has_res, has_temb = case
    G = 32
    g = torch.Generator().manual_seed(N + H + Co)
    x = torch.randn(N, Ci, H, W, generator=g).to(dev)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / math.sqrt(Ci * 9)).to(dev)
    bias = torch.randn(Co, generator=g).to(dev)
    temb = torch.randn(N, Co, generator=g).to(dev) if has_temb else None
    res = torch.randn(N, H, W, Co, generator=g).to(dev).to(dtype) if has_res else None
    gamma, beta = (1 + 0.2 * torch.randn(Co, generator=g)).to(dev), (0.1 * torch.randn(Co, generator=g)).to(dev)
    xh = ops.nchw_to_nhwc(x, dtype)
    wk = _repack(w, False, dtype)
    code = ops.dtype_code(xh)
    lib = _abi.lib()

    def run(fused):
        y = torch.zeros(N, H, W, Co, device=dev, dtype=dtype)
        a = torch.zeros(N, H, W, Co, device=dev, dtype=dtype)
        sums = torch.zeros(N, G, 2, device=dev)
        gp = GnParams(ops.t4_nhwc(y), ops.t4_nhwc(a), _null(), _null(), _null(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      None, None, None, N, H, W, Co, G, 1, 1e-5, 0)
        p = ConvParams(ops.t4_nhwc(xh), ops.t4_nhwc(y), ops.t4_nhwc(res) if has_res else _null(), wk.data_ptr(), 9 * Ci, 1, Ci,
                       bias.data_ptr(), temb.data_ptr() if has_temb else None, Co if has_temb else 0, N, H, W, Ci, H, W, Co, 3, 3, 1, 1, 0, code, 0, 0, None, 0)
        if fused:
            p.gn_fuse = C.cast(C.pointer(gp), C.c_void_p)
            p.gn_fuse_mode = 3
            if lib.dmu_conv2d_gn_fuse_supported(C.byref(p)) <= 0:
                pytest.skip("the persistent 3x3 kernel does not take this layer by its own heuristics")
        ops.conv2d_raw(p)
        _abi.check((lib.dmu_gn_apply if fused else lib.dmu_gn_forward)(C.byref(gp), _stream()))
        torch.cuda.synchronize()
        return y, a, sums
    y1, a1, s1 = run(True)
    y0, a0, s0 = run(False)
    assert torch.equal(y1, y0), "the conv output must not change"
    assert rel_l2(s1, s0) < 1e-5, "raw sums"
    yq = y1.float().permute(0, 3, 1, 2)
    aref = F.silu(F.group_norm(yq, G, gamma, beta, eps=1e-5))
    assert rel_l2(a1.float().permute(0, 3, 1, 2), aref) < TOL[dtype]
    assert rel_l2(a1, a0) < 2e-3
