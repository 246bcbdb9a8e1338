"""Development aid: back-to-back timing of the 3-channel boundary convolutions (stem fprop, head dgrad; tensor-core kernel vs the
SIMT edge kernel) against the bytes of the wide tensor they write."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, Tensor4

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


for N, H, Cm in [(128, 32, 64), (256, 64, 64)]:
    x = torch.randn(N, 3, H, H, device=dev)
    w = (torch.randn(Cm, 3, 3, 3, device=dev) / 5).bfloat16()      # [O][R][S][I]
    b = torch.zeros(Cm, device=dev)
    y = torch.empty(N, H, H, Cm, device=dev, dtype=torch.bfloat16)
    for impl, name in [(2, "tcgen05 stem"), (3, "SIMT edge")]:
        p = ConvParams(ops.t4_nchw(x), ops.t4_nhwc(y), null, w.data_ptr(), 27, 1, 3, b.data_ptr(), None, 0, N, H, H, 3, H, H, Cm, 3, 3, 1, 1, 0, 1, impl, 0)
        t = timeit(lambda: lib.dmu_conv2d(C.byref(p), s))
        mb = y.numel() * 2 / 1e6
        print(f"N={N} {H}x{H} 3->{Cm} {name:13s}: {t:7.1f} us  ({mb / t * 1e-3 * 1e3:5.2f} TB/s of the {mb:.0f} MB output)", flush=True)
