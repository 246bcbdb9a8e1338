"""Development aid: per-tap kernel (impl 4) against the persistent halo kernel (impl 5) on the 3x3 stride-1 layer shapes of
the UNet, back-to-back launches timed with CUDA events."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, Tensor4

lib = _abi.lib()
dev = torch.device("cuda:0")
null = Tensor4(None, 0, 0, 0, 0, 0, 0)


def run(N, H, Ci, Co, impl, gather=0, reps=30):
    R = 3
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
    w = (torch.randn(Co, R, R, Ci, device=dev) / math.sqrt(Ci * R * R)).bfloat16()
    y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(Co, device=dev)
    ws = torch.zeros(int(lib.dmu_conv2d_workspace_bytes()), dtype=torch.uint8, device=dev)
    p = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, w.data_ptr(), R * R * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co,
                   R, R, 1, 1, gather, 1, impl, 0, ws.data_ptr(), ws.numel())
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        _abi.check(lib.dmu_conv2d(C.byref(p), s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.dmu_conv2d(C.byref(p), s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, y


shapes = [(128, 32, 64, 64), (128, 32, 128, 64), (128, 16, 64, 64), (128, 16, 64, 128), (128, 16, 128, 64), (128, 16, 192, 64),
          (128, 8, 128, 128), (128, 8, 64, 128), (128, 8, 256, 128), (128, 8, 192, 64), (256, 64, 64, 64), (256, 64, 128, 64), (256, 32, 64, 64),
          (256, 32, 192, 64), (256, 16, 128, 128), (256, 16, 256, 128)]
for N, H, Ci, Co in shapes:
    fl = 2.0 * N * H * H * Ci * Co * 9
    t4, y4 = run(N, H, Ci, Co, 4)
    t5, y5 = run(N, H, Ci, Co, 5)
    err = (y4.float() - y5.float()).norm() / y4.float().norm()
    print(f"N={N} H={H} {Ci}->{Co}: per-tap {t4:6.1f} us ({fl / t4 / 1e6:5.0f} TF)   halo {t5:6.1f} us ({fl / t5 / 1e6:5.0f} TF)   rel diff {err:.1e}", flush=True)
