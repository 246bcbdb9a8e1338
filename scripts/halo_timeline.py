"""Development aid: per-tile clock64 stamps of CTA 0 of the persistent halo conv kernel."""
import ctypes as C, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_model_universal_b200 import ops, _abi
from diffusion_model_universal_b200._abi import ConvParams, Tensor4
lib = _abi.lib(); h = C.CDLL(_abi.LIB_PATH); h.dmu_debug_set_buffer.argtypes = [C.c_void_p]
dev = torch.device("cuda:0"); dbg = torch.zeros(2 * 8 * 64, dtype=torch.int64, device=dev); null = Tensor4(None, 0, 0, 0, 0, 0, 0)
def run(N, H, Ci, Co, res=False, gn=False):
    x = torch.randn(N, H, H, Ci, device=dev).bfloat16(); w = (torch.randn(Co, 3, 3, Ci, device=dev) / math.sqrt(Ci * 9)).bfloat16()
    y = torch.empty(N, H, H, Co, device=dev, dtype=torch.bfloat16); b = torch.zeros(Co, device=dev)
    r = torch.randn(N, H, H, Co, device=dev).bfloat16()
    coef = torch.ones(N * Ci * 2, device=dev)
    p = ConvParams(ops.t4_nhwc(x), ops.t4_nhwc(y), ops.t4_nhwc(r) if res else null, w.data_ptr(), 9 * Ci, 1, Ci, b.data_ptr(), None, 0, N, H, H, Ci, H, H, Co, 3, 3, 1, 1, 0, 1, 5, 0, None, 0,
                   coef.data_ptr() if gn else None, 1, 0, null)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        dbg.zero_(); h.dmu_debug_set_buffer(dbg.data_ptr()); _abi.check(lib.dmu_conv2d(C.byref(p), s)); torch.cuda.synchronize()
    h.dmu_debug_set_buffer(None)
    dall = dbg.view(-1, 8).cpu(); d = dall[:64]; tr = dall[64:]; d = d[d[:, 0] != 0]; tr = tr[tr[:, 0] != 0]; t0 = int(d[0, 0])
    print(f"--- N={N} H={H} {Ci}->{Co} res={res}: tiles of CTA 0 = {len(d)}")
    print("tile  mma:start  acc_empty_ok  a_full_ok  mmas_issued | epi:start_wait  acc_full_ok  stores_done")
    for i in range(len(d)):
        print(f"{i:4d} " + " ".join(f"{int(v) - t0:10d}" for v in d[i, :7]) + ("   | mma:before_a_full_wait  tma:wait_a_empty a_empty_ok loads_issued " + " ".join(f"{int(v) - t0:8d}" for v in tr[i, :4]) if i < len(tr) else ""))
cases = sys.argv[1] if len(sys.argv) > 1 else "gn"
if cases == "gn":
    run(128, 32, 64, 64); run(128, 32, 64, 64, gn=True); run(256, 64, 64, 64, gn=True)
else:
    run(128, 32, 64, 64); run(256, 64, 64, 64); run(256, 64, 64, 64, res=True)
