"""Development aid for ncu: a few GroupNorm forward / backward launches on the 32x32x64 and 16x16x128 tensors (small process, fast replay)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import GnParams, Tensor4

lib = _abi.lib()
dev = torch.device("cuda:0")
null = Tensor4(None, 0, 0, 0, 0, 0, 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for N, H, Cc in [(128, 32, 64), (128, 16, 128), (128, 8, 128)]:
    G = 32
    x = torch.randn(N, H, H, Cc, device=dev).bfloat16()
    y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x); a0 = torch.randn_like(x)
    sums = torch.zeros(N * G * 2, device=dev); red = torch.zeros(N * Cc * 2, device=dev)
    gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
    pf = GnParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, null, null, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, None, None,
                  N, H, H, Cc, G, 1, 1e-5, 0)
    pb = GnParams(ops.t4_nhwc(x), ops.t4_nhwc(dy), ops.t4_nhwc(dx), ops.t4_nhwc(a0), null, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  red.data_ptr(), None, None, N, H, H, Cc, G, 1, 1e-5, 0)
    for _ in range(2):
        _abi.check(lib.dmu_gn_forward(C.byref(pf), s))
        _abi.check(lib.dmu_gn_backward(C.byref(pb), s))
    torch.cuda.synchronize()
