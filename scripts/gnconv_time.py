"""Development aid: GroupNorm+SiLU followed by a 3x3 conv, unfused (dmu_gn_forward + halo conv) against fused
(dmu_gn_stats + dmu_gn_coef + halo conv with gn_coef), back-to-back launches, CUDA events."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, GnParams, Tensor4

lib = _abi.lib()
dev = torch.device("cuda:0")
null = Tensor4(None, 0, 0, 0, 0, 0, 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for N, H, Ci, Co, train in [(128, 32, 64, 64, 1), (128, 32, 64, 64, 0), (256, 64, 64, 64, 0), (256, 32, 64, 64, 0), (256, 32, 128, 64, 0)]:
    G = 32
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
    a = torch.empty_like(x)
    w = (torch.randn(Co, 3, 3, Ci, device=dev) / math.sqrt(9 * Ci)).bfloat16()
    y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(Co, device=dev)
    sums = torch.zeros(N * G * 2, device=dev); coef = torch.empty(N * Ci * 2, device=dev)
    gamma = torch.ones(Ci, device=dev); beta = torch.zeros(Ci, device=dev)
    pg = GnParams(ops.t4_nhwc(x), ops.t4_nhwc(a), null, null, null, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None, None,
                  N, H, H, Ci, G, 1, 1e-5, 0)
    pc = ConvParams(ops.t4_nhwc(a), ops.t4_nhwc(y), null, w.data_ptr(), 9 * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co,
                    3, 3, 1, 1, 0, 1, 5, 0, None, 0)
    pf = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, w.data_ptr(), 9 * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co,
                    3, 3, 1, 1, 0, 1, 5, 0, None, 0, coef.data_ptr(), 1, 0, ops.t4_nhwc(a) if train else null)

    def unfused():
        sums.zero_()
        lib.dmu_gn_forward(C.byref(pg), s)
        lib.dmu_conv2d(C.byref(pc), s)

    def fused():
        sums.zero_()
        lib.dmu_gn_stats(C.byref(pg), s)
        lib.dmu_gn_coef(C.byref(pg), coef.data_ptr(), s)
        lib.dmu_conv2d(C.byref(pf), s)

    def conv_only():
        lib.dmu_conv2d(C.byref(pc), s)

    def fconv_only():
        lib.dmu_conv2d(C.byref(pf), s)

    def variant(v):
        pf.gn_silu = v
        t = timeit(fconv_only)
        pf.gn_silu = 1
        return t

    print(f"N={N} H={H} {Ci}->{Co} train={train}: unfused {timeit(unfused):7.1f} us   fused {timeit(fused):7.1f} us   (plain conv {timeit(conv_only):6.1f}, fused conv alone {timeit(fconv_only):6.1f}; no act {variant(0):6.1f}, exp form {variant(2):6.1f})", flush=True)
