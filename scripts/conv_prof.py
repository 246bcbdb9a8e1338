"""Development aid for ncu: one launch each of the halo kernel, the per-tap kernel and the wgrad kernel on the largest layer of
the B=128 32x32 UNet (64->64 3x3), small process for fast replay."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, WgradParams, Tensor4

lib = _abi.lib()
dev = torch.device("cuda:0")
null = Tensor4(None, 0, 0, 0, 0, 0, 0)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
N, H, Ci, Co, R = 128, 32, 64, 64, 3
x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
w = (torch.randn(Co, R, R, Ci, device=dev) / math.sqrt(Ci * R * R)).bfloat16()
y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
dy = torch.randn(N, H, H, Co, device=dev).bfloat16()
b = torch.zeros(Co, device=dev)
dw = torch.zeros(Co, R, R, Ci, device=dev)
ws = torch.zeros(int(lib.dmu_conv2d_workspace_bytes()), dtype=torch.uint8, device=dev)
for impl in (5, 4):
    p = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, w.data_ptr(), R * R * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co,
                   R, R, 1, 1, 0, 1, impl, 0, ws.data_ptr(), ws.numel())
    for _ in range(2):
        _abi.check(lib.dmu_conv2d(C.byref(p), s))
wp = WgradParams(ops.t4_nhwc(dy), ops.t4_nhwc(x), dw.data_ptr(), R * R * Ci, 1, Ci, None, N, H, H, Co, H, H, Ci, R, R, 1, 1, 2)
for _ in range(2):
    _abi.check(lib.dmu_conv2d_wgrad(C.byref(wp), s))
torch.cuda.synchronize()
print("ok")
