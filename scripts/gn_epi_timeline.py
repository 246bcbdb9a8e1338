"""Development aid: clock64 stamps of conv_tc_kernel with and without the GroupNorm epilogue (cycles since CTA start)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, GnParams, Tensor4

lib = _abi.lib()
h = C.CDLL(_abi.LIB_PATH)
h.dmu_debug_set_buffer.argtypes = [C.c_void_p]
dev = torch.device("cuda:0")
dbg = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
null = Tensor4(None, 0, 0, 0, 0, 0, 0)


def run(N, H, Ci, Co, R, mode):
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
    w = (torch.randn(Co, R, R, Ci, device=dev) / math.sqrt(Ci * R * R)).bfloat16()
    y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
    a = torch.empty_like(y); xin = torch.randn_like(y); dx = torch.empty_like(y)
    G = 32
    sums = torch.rand(N, G, 2, device=dev) + 1; sums[..., 1] += 100
    red = torch.zeros(N, Co, 2, device=dev)
    gam, bet = torch.ones(Co, device=dev), torch.zeros(Co, device=dev)
    b = torch.zeros(Co, device=dev)
    gp = GnParams(ops.t4_nhwc(xin), ops.t4_nhwc(a), ops.t4_nhwc(dx), null, null, sums.data_ptr(), gam.data_ptr(), bet.data_ptr(), red.data_ptr(), None, None,
                  N, H, H, Co, G, 1, 1e-5, 0)
    p = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, w.data_ptr(), R * R * Ci, 1, Ci, b.data_ptr() if mode != 2 else None, None, 0, N, H, H, Ci, H, H, Co,
                   R, R, 1, R // 2, 0, 1, 4, 0, None, 0)
    if mode:
        p.gn_fuse = C.cast(C.pointer(gp), C.c_void_p); p.gn_fuse_mode = mode
        assert lib.dmu_conv2d_gn_fuse_supported(C.byref(p)) > 0
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        dbg.zero_()
        h.dmu_debug_set_buffer(dbg.data_ptr())
        _abi.check(lib.dmu_conv2d(C.byref(p), s))
        torch.cuda.synchronize()
    h.dmu_debug_set_buffer(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.dmu_conv2d(C.byref(p), s)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    d = dbg.view(-1, 8).cpu(); d = d[d[:, 0] != 0]
    rel = (d[:, 1:8] - d[:, 0:1]).float().mean(0)
    tag = ["plain", "gn-fwd", "gn-bwd"][mode]
    extra = f"tile-parked {rel[5]:.0f} reductions {rel[6]:.0f}" if mode else ""
    print(f"N={N} H={H} {Ci}->{Co} k{R} {tag:6s} ctas={len(d)} setup {rel[0]:.0f} first-stage {rel[1]:.0f} last-mma {rel[2]:.0f} acc-ready {rel[3]:.0f} {extra} done {rel[4]:.0f} | back-to-back {us:.1f} us", flush=True)


for shape in ((128, 1, 256, 256, 3), (128, 2, 256, 256, 3), (128, 4, 128, 128, 3), (128, 8, 128, 128, 3), (128, 8, 64, 64, 3)):
    for mode in (0, 1, 2):
        run(*shape, mode)
