"""Development aid: back-to-back timing of the GroupNorm forward / backward, column-sum and repack entry points on the
tensor shapes of the B=128, 32x32 UNet (warm L2, CUDA events) with the bytes each must move."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import GnParams, Tensor4

lib = _abi.lib()
dev = torch.device("cuda:0")
null = Tensor4(None, 0, 0, 0, 0, 0, 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for N, H, Cc in [(128, 32, 64), (128, 16, 64), (128, 16, 128), (128, 8, 128), (128, 8, 256), (128, 4, 128), (128, 2, 256), (128, 1, 256), (128, 1, 512),
                 (256, 64, 64), (256, 32, 64), (256, 32, 128), (256, 16, 128), (256, 16, 256), (256, 8, 256)]:
    G = 32
    x = torch.randn(N, H, H, Cc, device=dev).bfloat16()
    y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x); a0 = torch.randn_like(x)
    sums = torch.zeros(N * G * 2, device=dev); red = torch.zeros(N * Cc * 2, device=dev)
    gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
    pf = GnParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, null, null, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None, None,
                  N, H, H, Cc, G, 1, 1e-5, 0)
    pb = GnParams(ops.t4_nhwc(x), ops.t4_nhwc(dy), ops.t4_nhwc(dx), ops.t4_nhwc(a0), null, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  red.data_ptr(), None, None, N, H, H, Cc, G, 1, 1e-5, 0)
    def fwd():
        sums.zero_()
        _abi.check(lib.dmu_gn_forward(C.byref(pf), s))
    def bwd():
        red.zero_()
        _abi.check(lib.dmu_gn_backward(C.byref(pb), s))
    def zero():
        sums.zero_()
    def fwd2():
        sums.zero_()
        _abi.check(lib.dmu_gn_stats(C.byref(pf), s))
        _abi.check(lib.dmu_gn_apply(C.byref(pf), s))
    def bwd2():
        red.zero_()
        _abi.check(lib.dmu_gn_bwd_reduce(C.byref(pb), s))
        _abi.check(lib.dmu_gn_bwd_apply(C.byref(pb), s))
    tz = timeit(zero)
    tf, tb = timeit(fwd) - tz, timeit(bwd) - tz
    tf2, tb2 = timeit(fwd2) - tz, timeit(bwd2) - tz
    mb = x.numel() * 2 / 1e6
    print(f"N={N} {H}x{H}x{Cc} ({mb:6.1f} MB): gn fwd {tf:6.1f} us ({2 * mb / tf * 1e3 / 1e3:5.2f} TB/s of 2 passes)   gn bwd(+add) {tb:6.1f} us ({4 * mb / tb * 1e3 / 1e3:5.2f} TB/s of 4 passes)   two-pass fwd {tf2:6.1f} bwd {tb2:6.1f}", flush=True)
