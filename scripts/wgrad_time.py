"""Stand-alone launch time of the weight-gradient kernels (development aid): per-tap kernel (impl 2 below the halo threshold is
forced by DMU_WGRAD_HALO=0 in a second process) against the halo kernel (impl 5), back to back on one stream, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops
from diffusion_model_universal_b200._abi import WgradParams

CASES = [(128, 32, 32, 64, 64), (128, 32, 32, 128, 64), (128, 16, 16, 64, 64), (128, 16, 16, 128, 64), (128, 8, 8, 128, 128), (256, 64, 64, 64, 64)]


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for N, H, W, Ci, Co in CASES:
        xh = torch.randn(N, H, W, Ci, device=dev).to(torch.bfloat16)
        dyh = torch.randn(N, H, W, Co, device=dev).to(torch.bfloat16)
        dw = torch.zeros(Co, 3, 3, Ci, device=dev)
        line = f"{N}x{H}x{W} {Ci}->{Co}:"
        for impl in (2, 5):
            p = WgradParams(ops.t4_nhwc(dyh), ops.t4_nhwc(xh), dw.data_ptr(), 9 * Ci, 1, Ci, None, N, H, W, Co, H, W, Ci, 3, 3, 1, 1, impl)
            for _ in range(3): ops.wgrad_raw(p)
            torch.cuda.synchronize()
            ts = []
            for rep in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ops.wgrad_raw(p); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.wgrad_raw(p)
            e1.record(); torch.cuda.synchronize()
            b2b = e0.elapsed_time(e1) * 1e3 / 20
            fl = 2.0 * N * H * W * Ci * Co * 9
            line += f"  impl {impl}: cold {sorted(ts)[len(ts)//2]:7.1f} us, back-to-back {b2b:7.1f} us = {fl / b2b / 1e6:6.1f} TF/s"
        print(line, flush=True)


main()
