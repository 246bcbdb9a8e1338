"""Stand-alone launch time of the 4x4 stride-2 layers (development aid): per-tap kernel (impl 4) against the halo kernels (impl 5),
transposed gather (ConvTranspose2d forward / strided-conv dgrad) and strided conv, cold (L2 flushed) and back to back."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops
from diffusion_model_universal_b200._abi import ConvParams, Tensor4

null = Tensor4(None, 0, 0, 0, 0, 0, 0)
CASES = [(256, 32, 32), (256, 16, 16), (128, 16, 16), (128, 8, 8)]      # N, small-grid H, W (64 -> 64 channels)


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for N, h, w_ in CASES:
        small = torch.randn(N, h, w_, 64, device=dev).to(torch.bfloat16)
        big = torch.randn(N, 2 * h, 2 * w_, 64, device=dev).to(torch.bfloat16)
        wk = (torch.randn(64, 16, 64, device=dev) / 32).to(torch.bfloat16)
        bias = torch.zeros(64, device=dev)
        for name, gather in (("transposed", 1), ("strided", 0)):
            line = f"{N}x{h}x{w_} {name:10s}:"
            for impl in (4, 5):
                if gather == 1:
                    p = ConvParams(ops.t4_nhwc(small), ops.t4_nhwc(big), null, wk.data_ptr(), 16 * 64, 1, 64, bias.data_ptr(), None, 0,
                                   N, h, w_, 64, 2 * h, 2 * w_, 64, 4, 4, 2, 1, 1, 1, impl, 0, None, 0)
                else:
                    p = ConvParams(ops.t4_nhwc(big), ops.t4_nhwc(small), null, wk.data_ptr(), 16 * 64, 1, 64, bias.data_ptr(), None, 0,
                                   N, 2 * h, 2 * w_, 64, h, w_, 64, 4, 4, 2, 1, 0, 1, impl, 0, None, 0)
                for _ in range(3): ops.conv2d_raw(p)
                torch.cuda.synchronize()
                ts = []
                for rep in range(10):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); ops.conv2d_raw(p); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): ops.conv2d_raw(p)
                e1.record(); torch.cuda.synchronize()
                b2b = e0.elapsed_time(e1) * 1e3 / 20
                fl = 2.0 * N * h * w_ * 64 * 64 * 16
                line += f"  impl {impl}: cold {sorted(ts)[len(ts)//2]:7.1f} us, back-to-back {b2b:7.1f} us = {fl / b2b / 1e6:6.1f} TF/s"
            print(line, flush=True)


main()
