"""Development aid: clock64 stamps of the phases of conv_tc_kernel for a few layer shapes (cycles since CTA start)."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, Tensor4, RepackDesc

lib = _abi.lib()
h = C.CDLL(_abi.LIB_PATH)
h.dmu_debug_set_buffer.argtypes = [C.c_void_p]
dev = torch.device("cuda:0")
dbg = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
null = Tensor4(None, 0, 0, 0, 0, 0, 0)

def run(N, H, Ci, Co, R, ws=None, reps=3):
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16()
    w = (torch.randn(Co, R, R, Ci, device=dev) / math.sqrt(Ci * R * R)).bfloat16()
    y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(Co, device=dev)
    p = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), null, w.data_ptr(), R * R * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co,
                   R, R, 1, R // 2, 0, 1, 2, 0, ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(reps):
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
    flops = 2.0 * N * H * H * Ci * Co * R * R
    d = dbg.view(-1, 8).cpu()
    d = d[d[:, 0] != 0]
    if len(d) == 0:     # kernel without phase stamps (the persistent halo kernel)
        print(f"N={N} H={H} {Ci}->{Co} k{R}  back-to-back {us:.1f} us/launch  {flops / us / 1e6:.0f} TFLOP/s", flush=True)
        return
    rel = (d[:, 1:6] - d[:, 0:1]).float()
    print(f"N={N} H={H} {Ci}->{Co} k{R}  ctas={len(d)} kblocks={int(d[0,6])}  avg cycles since start: setup {rel[:,0].mean():.0f}  first-stage {rel[:,1].mean():.0f}  "
          f"last-mma-issued {rel[:,2].mean():.0f}  acc-ready {rel[:,3].mean():.0f}  epilogue-done {rel[:,4].mean():.0f}  issuer-waited {d[:,7].float().mean():.0f}   | back-to-back {us:.1f} us/launch  {flops / us / 1e6:.0f} TFLOP/s", flush=True)

ws = torch.zeros(int(lib.dmu_conv2d_workspace_bytes()), dtype=torch.uint8, device=dev)
run(128, 32, 64, 64, 3)
run(128, 32, 128, 64, 3)
run(256, 64, 64, 64, 3)
run(128, 16, 64, 64, 3)
run(128, 16, 64, 128, 3)
run(128, 16, 192, 64, 3)
run(128, 8, 128, 128, 3)
run(128, 4, 128, 128, 3)
run(128, 2, 256, 256, 3)
run(128, 2, 384, 128, 3)
run(128, 2, 384, 128, 3, ws)
run(128, 1, 256, 256, 3)
run(128, 4, 128, 128, 1)
