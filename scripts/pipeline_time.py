"""Times the ingest / sample-grid launches back-to-back with CUDA events and prints achieved HBM GB/s.
usage: python scripts/pipeline_time.py [B] [R]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
img = torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8, device=dev)
mean = torch.tensor([0.5, 0.5, 0.5], device=dev)
noise = torch.randn(B, 3, R, R, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
acp = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0).to(dev)
x0 = ops.ingest_u8(img, mean, mean, "NHWC")[0]
stack = torch.rand(11, 8, 3, R, R, device=dev)


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


n = B * 3 * R * R
rows = [
    ("ingest_u8 NHWC -> x0", lambda: ops.ingest_u8(img, mean, mean, "NHWC"), n * 5),
    ("ingest_u8 NHWC + q_sample -> xt", lambda: ops.ingest_u8(img, mean, mean, "NHWC", t, noise, acp, want_x0=False), n * 9),
    ("q_sample (fp32 in)", lambda: ops.q_sample(x0, t, noise, acp), n * 12),
    ("image_grid_u8 8 x 11 cells", lambda: ops.image_grid_u8(stack, nrow=11, transpose=True), stack.numel() * 5),
]
for name, fn, nbytes in rows:
    us = timeit(fn)
    print(f"{name:36s} {us:8.2f} us   {nbytes / us / 1e3:8.1f} GB/s algorithmic")
