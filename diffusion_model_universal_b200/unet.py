"""UNet denoiser: parameter containers + the CUDA execution engine.

Host-side mirror of the reference network (models/ddpm.py:32-135,
models/layers/{residual,attention,embeddings}.py).  The ``nn.Module`` tree
below only *holds parameters* — same names, shapes, registration order and
default initialisation as the reference, so ``state_dict()`` is interchangeable
 — its ``forward`` never touches ATen: it hands raw pointers to
a recorded plan of libdmu_b200.so launches (include/dmu_b200.h).

Data layout in HBM:
  * activations: NHWC, compute dtype (fp32 or bf16), one arena per plan;
    ``torch.cat`` of the up path never happens — producers write straight into
    channel slices of the consumer's buffer (row pitch != C);
  * parameters: one flat fp32 arena (the nn.Parameters are views into it),
    conv filters additionally repacked to [O][R][S][I] (fprop) and [I][R][S][O]
    (dgrad) in the compute dtype;
  * gradients: flat fp32 arena with the same offsets (``param.grad`` views).
"""

import ctypes as C
import math
import os
import warnings
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _abi, ops
from ._abi import F32, BF16, Tensor4, ConvParams, WgradParams, GnParams, AttnParams, RepackDesc, GnPgDesc, ColsumDesc


AUX, AUXK = 1000, 16      # lane id of micro-batch k's auxiliary chain j (sub-plan tag 2 + j): AUX + j * AUXK + k (Engine._run_forked)


def gn_groups(channels: int, num_groups: int = 32) -> int:
    """residual.py:22-29 group-count rule."""
    g = min(num_groups, channels)
    while channels % g != 0 and g > 1:
        g -= 1
    return g


# ============================================================================
# Parameter containers (never executed by ATen)
# ============================================================================
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: executed by the dmu_b200 engine, not by ATen")


class TransformerPositionalEmbedding(_Holder):
    """embeddings.py:11-39 (no parameters)."""

    def __init__(self, dimension: int):
        super().__init__()
        self.dimension = dimension


class TimeEmbedding(_Holder):
    """embeddings.py:41-64."""

    def __init__(self, base_dim: int, output_dim: int):
        super().__init__()
        self.positional_encoding = nn.Sequential(
            TransformerPositionalEmbedding(base_dim), nn.Linear(base_dim, output_dim), nn.GELU(), nn.Linear(output_dim, output_dim))
        for layer in self.positional_encoding:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.zeros_(layer.bias)


class ResidualBlock(_Holder):
    """residual.py:11-52."""

    def __init__(self, in_channels, out_channels, time_emb_channels, num_groups=32):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(gn_groups(in_channels, num_groups), in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.time_mlp = nn.Linear(time_emb_channels, out_channels)
        self.norm2 = nn.GroupNorm(gn_groups(out_channels, num_groups), out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else nn.Identity()
        nn.init.zeros_(self.time_mlp.weight)
        nn.init.zeros_(self.time_mlp.bias)
        nn.init.zeros_(self.conv2.weight)
        nn.init.zeros_(self.conv2.bias)


class SelfAttentionBlock(_Holder):
    """attention.py:10-27."""

    def __init__(self, in_channels, embedding_dim, num_heads=4, num_groups=32):
        super().__init__()
        self.num_heads = num_heads
        self.query_projection = nn.Linear(in_channels, embedding_dim)
        self.key_projection = nn.Linear(in_channels, embedding_dim)
        self.value_projection = nn.Linear(in_channels, embedding_dim)
        self.final_projection = nn.Linear(embedding_dim, embedding_dim)
        self.norm = nn.GroupNorm(num_groups, embedding_dim)


class _Stage(_Holder):
    """ConvDownBlock / ConvUpBlock / AttentionDownBlock / AttentionUpBlock (residual.py:70-255)."""

    def __init__(self, in_channels, out_channels, temb, attn: bool, up: bool):
        super().__init__()
        self.res_blocks = nn.ModuleList(
            [ResidualBlock(in_channels if i == 0 else out_channels, out_channels, temb) for i in range(2)])
        if attn:
            self.attention_blocks = nn.ModuleList([SelfAttentionBlock(out_channels, out_channels, 4) for _ in range(2)])
        if up:
            self.upsample = nn.ConvTranspose2d(out_channels, out_channels, kernel_size=4, stride=2, padding=1)
        else:
            self.downsample = nn.Conv2d(out_channels, out_channels, kernel_size=4, stride=2, padding=1)
        self.has_attn = attn


def down_plan(Cm):
    return [(False, Cm, Cm), (False, Cm, Cm), (False, Cm, 2 * Cm), (True, 2 * Cm, 2 * Cm), (False, 2 * Cm, 4 * Cm)]


def up_plan(Cm):
    return [(False, 8 * Cm, 4 * Cm), (True, 6 * Cm, 2 * Cm), (False, 4 * Cm, 2 * Cm), (False, 3 * Cm, Cm), (False, 2 * Cm, Cm)]


class UNet(nn.Module):
    """models/ddpm.py:32-135 — same constructor, same state_dict, CUDA-only forward."""

    def __init__(self, in_channels: int, model_channels: int, out_channels: int, precision: str = "fp32", sigma_embed: bool = False):
        super().__init__()
        Cm = model_channels
        self.in_channels, self.model_channels, self.out_channels = in_channels, Cm, out_channels
        self.initial_conv = nn.Conv2d(in_channels, Cm, kernel_size=3, padding="same")
        nn.init.kaiming_normal_(self.initial_conv.weight)
        self.time_embedding = TimeEmbedding(Cm, Cm * 4)
        self.down_blocks = nn.ModuleList([_Stage(ci, co, 4 * Cm, a, up=False) for a, ci, co in down_plan(Cm)])
        self.bottleneck = nn.Sequential(ResidualBlock(4 * Cm, 4 * Cm, 4 * Cm), SelfAttentionBlock(4 * Cm, 4 * Cm, 4), ResidualBlock(4 * Cm, 4 * Cm, 4 * Cm))
        self.up_blocks = nn.ModuleList([_Stage(ci, co, 4 * Cm, a, up=True) for a, ci, co in up_plan(Cm)])
        self.output_conv = nn.Sequential(nn.GroupNorm(32, Cm), nn.SiLU(), nn.Conv2d(Cm, out_channels, kernel_size=3, padding=1))
        if sigma_embed:  # models/score_based.py:57-61
            self.time_embed = nn.Sequential(nn.Linear(1, Cm), nn.SiLU(), nn.Linear(Cm, Cm * 4))
        self.sigma_embed = sigma_embed
        self.precision = precision
        self._engine = None

    @property
    def engine(self) -> "Engine":
        if self._engine is None:
            object.__setattr__(self, "_engine", Engine(self))
        return self._engine

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """eps_theta(x, t): x fp32 [B,C,H,W] (NCHW), t [B] -> fp32 [B,C,H,W]."""
        return self.engine.forward(x, t)


# ============================================================================
# Engine
# ============================================================================
class Buf:
    """NHWC activation view: channels [0,C) of rows with `pitch` elements."""
    __slots__ = ("addr", "N", "H", "W", "C", "pitch", "code", "esize", "grad", "root", "_written")

    def __init__(self, addr, N, H, W, C, pitch, code, root=None):
        self.addr, self.N, self.H, self.W, self.C, self.pitch, self.code = addr, N, H, W, C, pitch, code
        self.esize = 2 if code == BF16 else 4
        self.grad = None
        self.root = root if root is not None else self   # channel slices share their parent's state
        self._written = False

    # "has some consumer already written this tensor's gradient buffer?" (shared by all slices of a buffer)
    @property
    def grad_written(self):
        return self.grad is not None and self.grad.root._written

    @grad_written.setter
    def grad_written(self, v):
        self.grad.root._written = v

    def t4(self) -> Tensor4:
        return Tensor4(self.addr, self.H * self.W * self.pitch, self.W * self.pitch, self.pitch, 1, self.code, 0)

    def slice(self, c0, c) -> "Buf":
        b = Buf(self.addr + c0 * self.esize, self.N, self.H, self.W, c, self.pitch, self.code, root=self.root)
        if self.grad is not None:
            b.grad = self.grad.slice(c0, c)
        return b


def _null_t4():
    return Tensor4(None, 0, 0, 0, 0, 0, 0)


def _nchw_t4(addr, Cc, H, W):
    return Tensor4(addr, Cc * H * W, W, 1, H * W, F32, 0)


def _rows_t4(addr, pitch, code=F32):
    return Tensor4(addr, pitch, pitch, pitch, 1, code, 0)


class _Bump:
    def __init__(self, base=0):
        self.base, self.off = base, 0

    def take(self, nbytes, align=256):
        self.off = (self.off + align - 1) // align * align
        a = self.base + self.off
        self.off += nbytes
        return a


class Plan:
    """A recorded launch sequence for one (N, H, W, mode) problem."""

    def __init__(self):
        self.fwd, self.bwd = [], []
        self.bwd_parts = []     # bwd cut in three (head + up | bottleneck + down 4,3 | rest), for overlapping the DP all-reduce
        self.bwd_tails = []     # per part: GroupNorm parameter-gradient fold + staging unpack (what completes the part's arena range)
        self.ranges = []        # gradient-arena range that is final after each part
        self.arena = None       # torch uint8 tensor keeping all plan buffers alive
        # static I/O buffers: every pointer in the recorded launches is fixed, so a plan can be replayed as a CUDA graph
        self.x_in = None        # fp32 [N,Cin,H,W] (NCHW) network input, read by the stem conv (and its wgrad)
        self.t_in = None        # fp32 [N] timesteps (or sigmas)
        self.out = None         # fp32 [N,Cout,H,W] written by the head conv
        self.dout = None        # fp32 [N,Cout,H,W] upstream gradient (training plans)
        self.ws = None          # split-K scratch of the tensor-core convs (zero between launches)
        self.zero_ops = []      # memsets of everything a backward accumulates into (Engine.zero_backward_buffers)
        self.fwd_z = None       # training plans: fwd with zero_ops riding on an auxiliary chain (run_forward(zero_backward=True))
        self.zero_red = None    # (sub-plan) the GroupNorm / time-projection accumulators' memset
        self.graphs = {}        # "fwd"/"bwd" -> torch.cuda.CUDAGraph
        self.runs = {}          # "fwd"/"bwd" -> eager executions so far (the first one is the warm-up before capture)
        self.nlaunch = {}       # "fwd"/"bwd" -> kernels per replay
        self.busy = False       # activations saved for a pending backward


class Engine:
    """Builds and runs launch plans for one UNet instance."""
    _warned_fp32 = False

    def __init__(self, net: UNet):
        self.net = net
        self.code = BF16 if net.precision == "bf16" else F32
        if net.precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {net.precision!r}")
        if net.precision == "fp32" and not Engine._warned_fp32:
            # measured on B200 (bench.py extra.fp32_mode_train): ~4.0k img/s against ~44k img/s in bf16 mode
            Engine._warned_fp32 = True
            warnings.warn("diffusion_model_universal_b200: precision 'fp32' (the default, reference numerics: eps rel-L2 <= 1e-3) runs the "
                          "convolutions on fp32 SIMT kernels; add `precision: bf16` to model_config for the tcgen05 tensor-core path "
                          "(~11x faster, eps rel-L2 <= 2e-2).", stacklevel=3)
        self.tdtype = torch.bfloat16 if self.code == BF16 else torch.float32
        self.esize = 2 if self.code == BF16 else 4
        self.flat = None
        self.gflat = None
        self.plans = {}
        self.frozen = False     # weights known unchanged: skip the repack launch
        self.impl = 0           # conv implementation selector forwarded to the kernels (0 auto)
        self.use_graphs = True  # replay recorded plans as CUDA graphs (second execution onwards)
        # Independent sub-plans per batch, run as parallel graph branches.  Measured on B200 (scripts/phase_times.py): 1.43 /
        # 1.48 / 1.48 ms forward at 1 / 2 / 4 lanes for 128x32x32 - a chain of latency-bound kernels is as long for half the
        # batch, so lanes over the batch buy nothing; the default stays 1 (the machinery also carries the dgrad | wgrad lanes).
        self.micro_batches = int(os.environ.get("DMU_MICROBATCHES", "1"))
        # GroupNorm applied inside the halo conv kernel (dmu_conv_params.gn_coef).  Correct and tested, but measured slower than
        # GroupNorm launch + plain halo conv on B200 so far (scripts/gnconv_time.py: 39.5 vs 37.8 us at 128x32x32, 268 vs 192 us
        # at 256x64x64: the in-place transform is a latency chain in front of every tile's MMAs), so it is opt-in.
        self.fuse_gn = os.environ.get("DMU_GN_FUSE", "0") == "1"
        # GroupNorm in the EPILOGUE of the conv that produces its input (forward) / of the dgrad that produces its upstream
        # gradient (backward): the <= 8x8 stages, where one output tile holds whole images (dmu_conv_params.gn_fuse).  The library
        # decides per layer (dmu_conv2d_gn_fuse_supported); DMU_GN_EPI=0 keeps every GroupNorm a launch of its own (A/B aid).
        self.fuse_gn_epi = os.environ.get("DMU_GN_EPI", "1") != "0"
        # ResBlock shortcut convolutions (forward and dgrad) on an auxiliary graph branch instead of inside the main chain
        # (DMU_AUX_LANES=0: A/B aid)
        self.aux_lanes = os.environ.get("DMU_AUX_LANES", "1") != "0"
        self.fixed_sums = os.environ.get("DMU_GN_FIXED_SUMS", "1") != "0"     # order-independent GroupNorm statistics (0: float atomics, A/B aid)
        # GroupNorm statistics of the large layers accumulated by the producing conv (dmu_conv_params.gn_fuse_mode 3); DMU_GN_STATS=0: A/B aid
        self.fuse_gn_stats = os.environ.get("DMU_GN_STATS", "1") != "0"
        self._tail_stream = None
        self._lib = None

    # ------------------------------------------------------------------ parameters
    def _arena_order(self):
        """Arena order: time_mlp weights (block order) | time_mlp biases | per attention block q,k,v weights then
        q,k,v biases | everything else in registration order.  Contiguity lets the 22 time projections run as one
        GEMM and Q/K/V as one [3C, C] projection."""
        named = OrderedDict(self.net.named_parameters())
        res_prefixes = self.res_block_prefixes()
        order = [p + "time_mlp.weight" for p in res_prefixes] + [p + "time_mlp.bias" for p in res_prefixes]
        for ap in self.attn_block_prefixes():
            order += [ap + n + ".weight" for n in ("query_projection", "key_projection", "value_projection")]
            order += [ap + n + ".bias" for n in ("query_projection", "key_projection", "value_projection")]
        seen = set(order)
        order += [k for k in named if k not in seen]
        return named, order

    def res_block_prefixes(self):
        out = []
        for i in range(5):
            out += [f"down_blocks.{i}.res_blocks.{j}." for j in range(2)]
        out += ["bottleneck.0.", "bottleneck.2."]
        for i in range(5):
            out += [f"up_blocks.{i}.res_blocks.{j}." for j in range(2)]
        return out

    def attn_block_prefixes(self):
        return ([f"down_blocks.3.attention_blocks.{j}." for j in range(2)] + ["bottleneck.1."] +
                [f"up_blocks.1.attention_blocks.{j}." for j in range(2)])

    def _flatten(self, device):
        named, order = self._arena_order()
        offs, total = {}, 0
        for k in order:
            n = named[k].numel()
            offs[k] = (total, n)
            total += (n + 3) // 4 * 4
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        with torch.no_grad():
            for k in order:
                o, n = offs[k]
                flat[o:o + n].copy_(named[k].detach().reshape(-1))
                named[k].data = flat[o:o + n].view(named[k].shape)
        self.flat, self.offs, self.total = flat, offs, total
        self.named = named
        self.gflat = torch.zeros(total, device=device, dtype=torch.float32)
        # Derived filter caches in the compute dtype (refreshed by ONE repack launch per step).  Every conv filter twice:
        # [O][R][S][I] for fprop and [I][R][S][O] for dgrad, so that the contraction axis is contiguous in both directions
        # (what the TMA/UMMA K-major operand needs).  In bf16 mode the attention projections (q|k|v stacked, final) get
        # the same treatment so that they run on the tensor-core path too.
        entries = []   # (key, src tensor, rows, cols, R, S, is ConvTranspose2d)
        for k, p in named.items():
            if p.dim() == 4:
                entries.append((k, p, p.shape[0], p.shape[1], p.shape[2], p.shape[3], k.endswith("upsample.weight")))
        if self.code == BF16:
            for ap in self.attn_block_prefixes():
                q = named[ap + "query_projection.weight"]
                Cc = q.shape[0]
                entries.append((ap + "qkv", q, 3 * Cc, Cc, 1, 1, False))          # q, k, v weights are adjacent in the arena
                f = named[ap + "final_projection.weight"]
                entries.append((ap + "final_projection.weight", f, Cc, Cc, 1, 1, False))
        self.wc_off, wc_total, descs = {}, 0, []
        for key, p, a, b, R, S, is_t in entries:
            self.wc_off[key] = wc_total
            wc_total += (a * b * R * S + 63) // 64 * 64
        self.wc_half = max(wc_total, 64)
        self.wcache = torch.zeros(2 * self.wc_half, device=device, dtype=self.tdtype)
        # Filter-gradient staging, fp32, [first dim][R][S][second dim] per filter (same offsets as the cache): the wgrad
        # kernels' atomics are contiguous across a warp in this layout; one unpack launch per backward writes the
        # stored layout ([first][second][R][S]) into the gradient arena.
        self.gstage = torch.zeros(self.wc_half, device=device, dtype=torch.float32)
        undescs = []
        for key, p, a, b, R, S, is_t in entries:
            if p.dim() == 4:
                undescs.append(RepackDesc(self.gstage.data_ptr() + self.wc_off[key] * 4, self.gflat.data_ptr() + offs[key][0] * 4,
                                          a, b, R, S, 3, F32))
        # The backward is cut in three parts (head + up path | bottleneck + down 4, 3 | down 2..0 + stem + time embedding);
        # the filters whose gradients are complete after each part get their own unpack table, and in the gradient arena
        # (registration order) they are the ranges [cuts[0], total), [cuts[1], cuts[0]) and [0, cuts[1]).
        def part_of(k):
            if k.startswith("up_blocks.") or k.startswith("output_conv."):
                return 0
            if k.startswith("bottleneck.") or k.startswith("down_blocks.3.") or k.startswith("down_blocks.4."):
                return 1
            return 2
        conv_entries = [e_ for e_ in entries if e_[1].dim() == 4]
        self.unpack_tables = []
        for part in range(3):
            und = [d for d, e_ in zip(undescs, conv_entries) if part_of(e_[0]) == part]
            uarr = (RepackDesc * max(len(und), 1))(*und)
            self.unpack_tables.append((torch.frombuffer(bytearray(bytes(uarr)), dtype=torch.uint8).to(device), len(und)))
        rest_lo = offs["initial_conv.weight"][0]
        self.cuts = [min(offs[k][0] for k in order if offs[k][0] >= rest_lo and part_of(k) == part) for part in (0, 1)]
        self.tail_lo = self.cuts[0]
        max_numel = 1
        for key, p, a, b, R, S, is_t in entries:
            off = self.wc_off[key]
            fwd = self.wcache.data_ptr() + off * self.esize
            bwd = self.wcache.data_ptr() + (self.wc_half + off) * self.esize
            if is_t:   # IOHW
                descs.append(RepackDesc(p.data_ptr(), fwd, b, a, R, S, 1, self.code))   # [O][R][S][I]
                descs.append(RepackDesc(p.data_ptr(), bwd, a, b, R, S, 0, self.code))   # [I][R][S][O]
            else:      # OIHW (or a Linear's [O][I])
                descs.append(RepackDesc(p.data_ptr(), fwd, a, b, R, S, 0, self.code))   # [O][R][S][I]
                descs.append(RepackDesc(p.data_ptr(), bwd, b, a, R, S, 1, self.code))   # [I][R][S][O]
            max_numel = max(max_numel, a * b * R * S)
        arr = (RepackDesc * len(descs))(*descs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.repack_table = host.to(device)
        self.repack_n, self.repack_max = len(descs), max_numel
        self.plans = {}
        self.device = device

    def _params_in_arena(self):
        if self.flat is None:
            return False
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * 4
        p0 = self.net.initial_conv.weight
        return p0.device == self.flat.device and lo <= p0.data_ptr() < hi and lo <= self.net.output_conv[2].bias.data_ptr() < hi

    def prepare(self, device):
        if self._lib is None:
            self._lib = _abi.lib()
        if not self._params_in_arena():
            self._flatten(device)

    def paddr(self, name):  # fp32 address of a parameter inside the flat arena
        return self.flat.data_ptr() + self.offs[name][0] * 4

    def gaddr(self, name):
        return self.gflat.data_ptr() + self.offs[name][0] * 4

    def waddr(self, name):  # repacked filter, fprop layout [O][R][S][I]
        return self.wcache.data_ptr() + self.wc_off[name] * self.esize

    def gsaddr(self, name):  # fp32 gradient staging of a filter, [first dim][R][S][second dim]
        return self.gstage.data_ptr() + self.wc_off[name] * 4

    def waddr_t(self, name):  # repacked filter, dgrad layout [I][R][S][O]
        return self.wcache.data_ptr() + (self.wc_half + self.wc_off[name]) * self.esize

    def repack(self, stream):
        ops.LAUNCHES += 1
        _abi.check(self._lib.dmu_repack_weights(self.repack_table.data_ptr(), self.repack_n, self.repack_max, stream), "repack_weights")

    # ------------------------------------------------------------------ public entry points
    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        ops._need_cuda(x, t)
        if x.dim() != 4 or x.shape[1] != self.net.in_channels:
            raise ValueError(f"expected x of shape [B,{self.net.in_channels},H,W], got {tuple(x.shape)}")
        if x.shape[2] % 32 != 0 or x.shape[3] % 32 != 0:
            raise ValueError("UNet has five stride-2 stages: H and W must be multiples of 32")
        if t.dim() != 1 or t.shape[0] != x.shape[0]:
            raise ValueError("t must have shape [B]")
        self.prepare(x.device)
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("the gradient with respect to the network INPUT is not computed on this path (the recorded backward "
                                      "stops at the stem's weight gradient, like the reference's training loop needs); detach x")
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.net.parameters())
        x = x.contiguous().float()
        if need_grad:
            params = [self.named[k] for k in self.offs]
            return _UNetFn.apply(self, x, t, *params)
        return self.run_forward(x, t, self.get_plan(x.shape, False))

    def get_plan(self, shape, train: bool) -> Plan:
        key = (tuple(shape), train)
        lst = self.plans.setdefault(key, [])
        for p in lst:
            if not p.busy:
                return p
        p = self._build(shape[0], shape[2], shape[3], train)
        lst.append(p)
        return p

    def _run(self, oplist, stream):
        """Eager execution: one stream, list order (side-lane tags and join markers are ignored)."""
        for op in oplist:
            if op[0] is None:
                continue
            ops.LAUNCHES += 1
            rc = op[0](*op[1], stream)
            if rc != 0:
                _abi.check(rc, op[0].__name__)

    def _run_forked(self, oplist):
        """Capture-time execution over lanes.  Lane 0 is the capturing stream; lane 2k is the main chain of micro-batch k,
        lane 2k+1 its side chain (weight gradients, bias / time-projection column sums: they only read finished tensors and
        accumulate into their own outputs), lane AUX+k its auxiliary chain (an independent branch of the layer graph that
        rejoins: a ResBlock's 1x1 shortcut convolution next to its conv1 -> norm2 chain).  Markers: (None, (), -2) = fork
        point (main lanes start after everything issued on lane 0 so far), (None, (), 2k) = side lane 2k+1 joins main lane 2k,
        (None, (), AUX+k) = the auxiliary lane joins main lane 2k, (None, (), -1) = every lane joins lane 0.  An op on a side
        or auxiliary lane depends on everything issued on its main lane before it."""
        # (Measured in round 2: capturing the main chain on a high-priority stream, so that its CTAs are placed before the weight-
        # gradient lane's, LOSES 100 us per step - 2.78 vs 2.68 ms: the starved lane becomes the critical path at the part ends.)
        main = torch.cuda.current_stream()
        streams, ptrs, started, dirty = {0: main}, {}, {0}, set()
        fork = None

        def main_of(lane):
            if lane >= AUX:
                return 2 * ((lane - AUX) % AUXK)
            return lane - 1 if lane % 2 == 1 else lane

        def lane_stream(lane):
            if lane not in streams:
                streams[lane] = torch.cuda.Stream(device=self.device)
            if lane not in started:
                started.add(lane)
                if main_of(lane) == lane:
                    if fork is not None:
                        streams[lane].wait_event(fork)
                    else:
                        streams[lane].wait_stream(main)
            if lane not in ptrs:
                ptrs[lane] = C.c_void_p(streams[lane].cuda_stream)
            return streams[lane]

        for op in oplist:
            lane = op[2] if len(op) > 2 else 0
            if op[0] is None:
                if lane == -2:
                    fork = torch.cuda.Event()
                    fork.record(main)
                elif lane == -1:
                    for l in sorted(dirty, reverse=True):      # sides / auxiliaries into their mains first, then mains into lane 0
                        tgt = main_of(l) if main_of(l) != l else 0
                        if l != 0:
                            lane_stream(tgt).wait_stream(streams[l])
                            if tgt != 0:
                                dirty.add(tgt)
                    dirty.clear()
                elif lane >= AUX:
                    if lane in dirty:
                        lane_stream(main_of(lane)).wait_stream(streams[lane])
                        dirty.discard(lane)
                        dirty.add(main_of(lane))
                else:
                    if lane + 1 in dirty:
                        lane_stream(lane).wait_stream(streams[lane + 1])
                        dirty.discard(lane + 1)
                        dirty.add(lane)
                continue
            st = lane_stream(lane)
            if main_of(lane) != lane:
                st.wait_stream(lane_stream(main_of(lane)))
            ops.LAUNCHES += 1
            rc = op[0](*op[1], ptrs[lane])
            if rc != 0:
                _abi.check(rc, op[0].__name__)
            if lane != 0:
                dirty.add(lane)
        for l in sorted(dirty, reverse=True):
            if l != 0:
                main.wait_stream(streams[l])

    def _execute(self, plan: Plan, which: str):
        """Run plan.fwd / plan.bwd: eagerly the first time (warms every lazy one-time initialisation), then captured once
        into a CUDA graph and replayed — ~250 launches (and their tensor-map encodes) become one host call."""
        oplist = (plan.fwd if which == "fwd" else plan.fwd_z if which == "fwd_z" else plan.bwd if which == "bwd" else
                  plan.bwd_tails[int(which[4:])] if which.startswith("tail") else plan.bwd_parts[int(which[3:])])
        if self.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            # the caller is capturing (TrainStep's whole-step graph): record the launches, lanes included, into ITS graph
            self._run_forked(oplist)
            return
        g = plan.graphs.get(which)
        if g is not None:
            g.replay()
            ops.LAUNCHES += plan.nlaunch[which]
            return
        n = plan.runs.get(which, 0)
        plan.runs[which] = n + 1
        graphable = self.use_graphs and self.device.type == "cuda" and n >= 1 and not torch.cuda.is_current_stream_capturing()
        if not graphable:
            self._run(oplist, ops._stream())
            return
        g = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        with torch.cuda.graph(g):
            self._run_forked(oplist)             # recorded on the capture stream (+ forked lanes), not executed
        plan.nlaunch[which] = ops.LAUNCHES - n0
        ops.LAUNCHES = n0
        plan.graphs[which] = g
        g.replay()
        ops.LAUNCHES += plan.nlaunch[which]

    def run_forward(self, x, t, plan: Plan, repacked: bool = False, clone: bool = True, zero_backward: bool = False) -> torch.Tensor:
        """x = None: the caller has already written the network input into plan.x_in (TrainStep's q_sample does).
        repacked: the filter caches are already current (refreshed on a side stream).  clone=False returns the plan's static
        output buffer itself (overwritten by the next execution).  zero_backward: also run zero_backward_buffers(plan), beside the
        latency-bound stages of this forward (the caller then passes prezeroed=True to run_backward; only for a caller that owns
        the whole step: the memsets wipe the gradient arena)."""
        if not (self.frozen or repacked):
            self.repack(ops._stream())
        if x is not None:
            plan.x_in.copy_(x)
        plan.t_in.copy_(t)       # int64 timesteps are cast to fp32 here exactly like embeddings.py:35 promotes them
        if zero_backward and plan.fwd_z is None:
            self.zero_backward_buffers(plan)
        self._execute(plan, "fwd_z" if (zero_backward and plan.fwd_z is not None) else "fwd")
        return plan.out.clone() if clone else plan.out

    def zero_backward_buffers(self, plan: Plan, stream=None):
        """The memsets a backward of `plan` needs beforehand (gradient arena, filter-gradient staging, GroupNorm / time-projection
        accumulators); anything between the previous optimizer update and this plan's backward is early enough."""
        self._run(plan.zero_ops, stream if stream is not None else ops._stream())

    def run_backward(self, plan: Plan, dout, between=None, prezeroed=False):
        """Fills the gradient arena; returns it (flat fp32, same offsets as the parameter arena).  ``between(lo, hi)`` is
        called after each of the three parts of the backward with the arena range that just became final.  prezeroed: the caller
        has already run zero_backward_buffers(plan) since the last update."""
        g = self.gflat
        # param.grad tensors handed out by an earlier backward are views of this arena.  If any is still installed
        # (gradient accumulation, zero_grad(set_to_none=False)) detach it first so autograd's `grad += new` stays correct.
        lo, hi = g.data_ptr(), g.data_ptr() + g.numel() * 4
        for p in self.named.values():
            if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                p.grad = p.grad.clone()
        if dout is not None:      # None: the caller wrote the upstream gradient straight into plan.dout
            plan.dout.copy_(dout)
        if not prezeroed:
            self.zero_backward_buffers(plan)
        if between is None:
            self._execute(plan, "bwd")
        elif self.device.type != "cuda":
            for i in range(len(plan.bwd_parts)):
                self._execute(plan, "bwd%d" % i)
                if plan.bwd_tails[i]:
                    self._execute(plan, "tail%d" % i)
                between(*plan.ranges[i])      # gflat[lo:hi] is final: e.g. start its all-reduce while the next part runs
        else:
            # the tail of part i (GroupNorm parameter-gradient fold, staging unpack: ~45 us) and the caller's hook (the all-reduce of the
            # range that tail completes) go to a second stream; the main stream continues with part i + 1 at once
            cur = torch.cuda.current_stream(self.device)
            if self._tail_stream is None:
                self._tail_stream = torch.cuda.Stream(device=self.device)
            ts = self._tail_stream
            for i in range(len(plan.bwd_parts)):
                self._execute(plan, "bwd%d" % i)
                ts.wait_stream(cur)
                with torch.cuda.stream(ts):
                    if plan.bwd_tails[i]:
                        self._execute(plan, "tail%d" % i)
                    between(*plan.ranges[i])      # gflat[lo:hi] is final: e.g. start its all-reduce while the next part runs
            cur.wait_stream(ts)
        return g

    # ------------------------------------------------------------------ plan construction
    def _build(self, N, H, W, train) -> Plan:
        """A plan is K independent sub-plans over N/K images each ("micro-batch lanes").  Per-image arithmetic is untouched
        (GroupNorm and attention are per image, weight gradients accumulate atomically), but in the captured graph the K
        chains run side by side: the launches of the <= 8x8 stages are latency-bound and leave most SMs idle, so two of
        them overlap almost for free."""
        K = self.micro_batches if (self.device.type == "cuda" and N % max(self.micro_batches, 1) == 0 and
                                   N // max(self.micro_batches, 1) >= 8) else 1
        Ns = N // K
        net = self.net
        plan = Plan()
        plan.keep = []
        plan.x_in = torch.zeros(N, net.in_channels, H, W, device=self.device, dtype=torch.float32)
        plan.t_in = torch.zeros(N, device=self.device, dtype=torch.float32)
        plan.out = torch.zeros(N, net.out_channels, H, W, device=self.device, dtype=torch.float32)
        if train:
            plan.dout = torch.zeros(N, net.out_channels, H, W, device=self.device, dtype=torch.float32)
        dry = _PlanBuilder(self, Ns, H, W, train, base=(0, 0, 0))
        dry.build()
        rnd = lambda v: (v + 255) // 256 * 256
        n_main, n_stats, n_red = rnd(dry.bump.off), rnd(dry.stats.off + 4), rnd(dry.red.off + 4)
        per = n_main + n_stats + n_red
        nbytes = K * per + 256
        arena = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
        base = rnd(arena.data_ptr())
        subs, gn_pg = [], []
        for k in range(K):
            b0 = base + k * per
            sub = _PlanBuilder(self, Ns, H, W, train, base=(b0, b0 + n_main, b0 + n_main + n_stats))
            sp = sub.plan
            sp.x_in, sp.t_in, sp.out = plan.x_in[k * Ns:(k + 1) * Ns], plan.t_in[k * Ns:(k + 1) * Ns], plan.out[k * Ns:(k + 1) * Ns]
            if train:
                sp.dout = plan.dout[k * Ns:(k + 1) * Ns]
            if self.code == BF16:   # split-K scratch: one per lane (the lanes run concurrently)
                sp.ws = torch.zeros(int(self._lib.dmu_conv2d_workspace_bytes()), device=self.device, dtype=torch.uint8)
            sub.build()
            subs.append(sub)
            plan.keep.append(sp)
            gn_pg += sub.gn_pg

        def retag(ops_, k):
            """sub-plan tags (none = main chain, 1 = side chain, 2 + j = auxiliary chain j) -> lane ids of micro-batch k"""
            out = []
            for op in ops_:
                tag = op[2] if len(op) == 3 else 0
                if op[0] is None:
                    out.append((None, (), AUX + (tag - 2) * AUXK + k if tag >= 2 else 2 * k))   # join an auxiliary / the side lane into the main lane
                else:
                    out.append((op[0], op[1], 2 * k if tag == 0 else 2 * k + 1 if tag == 1 else AUX + (tag - 2) * AUXK + k))
            return out

        lib = self._lib
        plan.fwd = [(None, (), -2)]                                      # fork point
        for k, sub in enumerate(subs):
            plan.fwd += retag(sub.plan.fwd, k)
        plan.fwd.append((None, (), -1))                                  # join all lanes
        fwd_marked = plan.fwd
        plan.fwd = [op for op in fwd_marked if op[0] != "zero_here"]
        if train:
            # one zeroing of the gradient arena / staging for all lanes, then the lanes, then the batch-folded tails
            # Three parts (see Engine._flatten): after part i the arena range ranges[i] is final, so a data-parallel step can
            # all-reduce it while the next part still runs.
            parts = ([], [], [])
            pg = ([], [], [])
            for k, sub in enumerate(subs):
                ops_k = retag(sub.plan.bwd, k)
                cuts = [i for i, op in enumerate(ops_k) if op[0] == "split"]
                assert len(cuts) == 2
                parts[0].extend(ops_k[:cuts[0]])
                parts[1].extend(ops_k[cuts[0] + 1:cuts[1]])
                parts[2].extend(ops_k[cuts[1] + 1:])
                g0, g1 = sub.gn_pg_split
                pg[0].extend(sub.gn_pg[:g0]); pg[1].extend(sub.gn_pg[g0:g1]); pg[2].extend(sub.gn_pg[g1:])
            plan.gn_pg_tables = []
            # what a backward accumulates into (filter-gradient staging, gradient arena, GroupNorm / time-projection sums):
            # zeroed by Engine.zero_backward_buffers - at the start of run_backward, or earlier by the caller (TrainStep does it
            # on a branch next to the forward)
            plan.zero_ops = [(lib.dmu_zero, (self.gstage.data_ptr(), self.gstage.numel() * 4)),
                             (lib.dmu_zero, (self.gflat.data_ptr(), self.gflat.numel() * 4))] + [sub.plan.zero_red for sub in subs]
            plan.bwd_parts = [[], [], []]
            # forward of a training step that zeroes them on the way: the memsets (2 x 64 MB, ~25 us of the whole GPU) hang off the
            # first lane's chain where the <= 4x4 stages begin - those launches are latency-bound and leave most SMs idle - on an
            # auxiliary chain of their own, joined with everything else at the end of the forward
            plan.fwd_z, placed = [], False
            for op in fwd_marked:
                if op[0] == "zero_here":
                    if not placed:
                        plan.fwd_z += [(fn, args, AUX + 2 * AUXK) for fn, args in plan.zero_ops]
                        placed = True
                else:
                    plan.fwd_z.append(op)
            if not placed:
                plan.fwd_z = None
            tails = ([], [], [])       # per part: batch fold of the GroupNorm parameter gradients + staging unpack
            for h, lst in enumerate(plan.bwd_parts):
                lst.append((None, (), -2))
                lst.extend(parts[h])
                lst.append((None, (), -1))
                if pg[h]:
                    arr = (GnPgDesc * len(pg[h]))(*pg[h])
                    tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
                    plan.gn_pg_tables.append(tab)
                    tails[h].append((lib.dmu_gn_param_grads, (tab.data_ptr(), len(pg[h]), max(d.C for d in pg[h]), Ns), 0))
                # filter gradients: staging layout -> the parameters' own layout inside the gradient arena
                tab, n_un = self.unpack_tables[h]
                if n_un:
                    tails[h].append((lib.dmu_repack_weights, (tab.data_ptr(), n_un, self.repack_max), 0))
            # the parts form (data parallel): the tail of part h runs on a stream of its own, in front of that part's all-reduce,
            # while the main stream already replays part h + 1 (Engine.run_backward)
            plan.bwd_tails = [list(t) for t in tails]
            total = self.gflat.numel()
            plan.ranges = [(self.cuts[0], total), (self.cuts[1], self.cuts[0]), (0, self.cuts[1])]
            # One-graph form (no all-reduce between the parts): the tail of part h only touches part h's gradients, so it rides
            # on the side lane at the start of part h + 1 instead of standing between the two parts on lane 0.
            plan.bwd = []
            for h in range(3):
                body = plan.bwd_parts[h]
                if h == 0:
                    plan.bwd += body
                else:
                    i = body.index((None, (), -2)) + 1
                    plan.bwd += body[:i] + [(fn, args, 1) for fn, args, _ in tails[h - 1]] + body[i:]
            # the last tail runs next to the time-embedding backward: on the side lane, behind the stem's weight gradient (the unpack
            # reads what that writes) - where the sub-plan left its "tails" marker
            marks = [i for i, op in enumerate(plan.bwd) if op[0] == "tails"]
            if K == 1 and len(marks) == 1:
                i = marks[0]
                plan.bwd[i:i + 1] = [(fn, args, 1) for fn, args, _ in tails[2]]
            else:
                plan.bwd = [op for op in plan.bwd if op[0] != "tails"] + tails[2]
            # parts form: the last part's tail rides on its side lane in the same place (next to the embedding backward) instead of
            # following the part on the tail stream; the other parts' tails stay in bwd_tails
            marks = [i for i, op in enumerate(plan.bwd_parts[2]) if op[0] == "tails"]
            if K == 1 and len(marks) == 1:
                i = marks[0]
                plan.bwd_parts[2][i:i + 1] = [(fn, args, 1) for fn, args, _ in tails[2]]
                plan.bwd_tails[2] = []
            plan.bwd_parts = [[op for op in lst if op[0] != "tails"] for lst in plan.bwd_parts]
        plan.arena = arena
        plan.nbytes = nbytes
        plan.lanes = K
        plan.gn_fused = [sum(sb.n_gn_fused[i] for sb in subs) for i in (0, 1)]   # GroupNorms riding in conv epilogues (fwd, bwd)
        return plan


class _PlanLease:
    """Marks a training plan busy (its activation arena holds what a pending backward needs) until the backward has run OR
    the autograd node that owns the lease is collected - a grad-enabled forward whose graph is dropped (an evaluation loss
    outside no_grad, an exception) must not strand the plan: get_plan would then allocate a new arena on every such call."""

    def __init__(self, plan):
        self.plan = plan
        plan.busy = True

    def release(self):
        if self.plan is not None:
            self.plan.busy = False
            self.plan = None

    def __del__(self):
        self.release()


class _UNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: Engine, x, t, *params):
        plan = eng.get_plan(x.shape, True)
        ctx.lease = _PlanLease(plan)
        out = eng.run_forward(x, t, plan)
        ctx.eng, ctx.plan = eng, plan
        return out

    @staticmethod
    def backward(ctx, dout):
        eng, plan = ctx.eng, ctx.plan
        try:
            g = eng.run_backward(plan, dout)
        finally:
            ctx.lease.release()
        grads = []
        for k in eng.offs:
            o, n = eng.offs[k]
            p = eng.named[k]
            grads.append(g[o:o + n].view(p.shape) if p.requires_grad else None)
        return (None, None, None) + tuple(grads)


class _PlanBuilder:
    def __init__(self, eng: Engine, N, H, W, train, base):
        self.e, self.N, self.H, self.W, self.train = eng, N, H, W, train
        main_base, stats_base, red_base = base
        self.bump = _Bump(main_base)        # activations, gradients of activations, fp32 rows
        self.stats = _Bump(stats_base)      # GroupNorm (sum, sumsq) accumulators: zeroed by one launch per forward
        self.red = _Bump(red_base)          # GroupNorm backward accumulators: zeroed by one launch per backward
        self.plan = Plan()
        self.lib = eng._lib
        self.code, self.esize = eng.code, eng.esize
        self.tape = []   # backward emitters, run in reverse
        self.gn_pg = []         # (red, dgamma, dbeta, C) of every GroupNorm backward: folded by one launch at the end
        self.gn_pg_split = []   # len(gn_pg) at each split marker of the backward
        self.side_lane = True   # weight-gradient / column-sum launches of the backward go to the graph's second branch
        self.leaf_tag = 0
        self.cs_pending = []    # column sums (bias gradients, time-projection sums) batched into one launch per backward part
        self.temb_join_pending = False
        self.prod = {}          # output address -> ConvParams of the forward conv that writes it (candidates for a fused GroupNorm)
        self.n_gn_fused = [0, 0]   # GroupNorms that ride in a conv epilogue: [forward, backward]
        self.n_gn_stats = 0        # GroupNorms whose statistics the producing conv accumulates (the apply pass stays a launch)

    # ---- allocation helpers
    def act(self, H, W, Cc, want_grad=True) -> Buf:
        b = Buf(self.bump.take(self.N * H * W * Cc * self.esize), self.N, H, W, Cc, Cc, self.code)
        if self.train and want_grad:
            b.grad = Buf(self.bump.take(self.N * H * W * Cc * self.esize), self.N, H, W, Cc, Cc, self.code)
        return b

    def tmp(self, H, W, Cc) -> Buf:
        return Buf(self.bump.take(self.N * H * W * Cc * self.esize), self.N, H, W, Cc, Cc, self.code)

    def f32(self, n):
        return self.bump.take(n * 4)

    # ---- op emitters
    def emit(self, lst, fn, *args):
        lst.append((fn, args))

    def gp(self, name):
        """Gradient-arena address of a parameter."""
        return self.e.gaddr(name)

    def colsum(self, t4: Tensor4, N, H, W, Cc, out_nc, pitch, out_c):
        """Per-image (out_nc) and / or total (out_c) channel sums of an NHWC tensor.  While the side lane is open the request is
        only recorded: flush_colsums() turns every request of a backward part into ONE dmu_colsum_multi launch."""
        if not (self.side_lane and t4.sc == 1 and t4.dtype == self.code):
            self.plan.keep.append(t4)
            op = (self.lib.dmu_colsum, (C.byref(t4), N, H, W, Cc, out_nc, pitch, out_c, 1.0))
            self.plan.bwd.append(op + (1,) if self.side_lane else op)
            return
        if out_nc is None and t4.sh == W * t4.sw and t4.sn == H * t4.sh:    # only the total: one flat pixel range
            t4 = Tensor4(t4.ptr, N * H * W * t4.sw, N * H * W * t4.sw, t4.sw, 1, t4.dtype, 0)
            N, H, W = 1, 1, N * H * W
        chunks = max(1, min(148 if N == 1 else 8, (H * W * Cc) // 65536))
        self.cs_pending.append(ColsumDesc(t4, out_nc, pitch, out_c, N, H, W, Cc, chunks, 1.0, 0, 0))

    def flush_colsums(self):
        if not self.cs_pending:
            return
        cta = 0
        for d in self.cs_pending:
            d.cta0 = cta
            cta += d.N * d.chunks
        arr = (ColsumDesc * len(self.cs_pending))(*self.cs_pending)
        dev = self.e.device
        tab = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self.plan.keep.append(tab)
        # its own auxiliary chain (not the weight-gradient lane's queue): the sums only need finished main-lane tensors, and the
        # embedding backward at the very end of the step waits for the last batch of them
        self.plan.bwd.append((self.lib.dmu_colsum_multi, (tab.data_ptr(), len(self.cs_pending), cta, self.code), 3 if self.e.aux_lanes else 1))
        self.cs_pending = []

    def conv(self, lst, x: Tensor4, y: Tensor4, w, w_strides, w_code, dims, geom, bias=None, temb=None, temb_pitch=0, res=None, gather=0,
             gn_coef=None, gn_silu=0, a_out=None, emit=True, lane=0):
        N, Hi, Wi, Ck, Ho, Wo, Cj = dims
        R, S, stride, pad = geom
        p = ConvParams(x, y, res if res is not None else _null_t4(), w, w_strides[0], w_strides[1], w_strides[2], bias, temb, temb_pitch,
                       N, Hi, Wi, Ck, Ho, Wo, Cj, R, S, stride, pad, gather, w_code, self.e.impl, 0,
                       self.plan.ws.data_ptr() if self.plan.ws is not None else None, self.plan.ws.numel() if self.plan.ws is not None else 0,
                       gn_coef, gn_silu, 0, a_out if a_out is not None else _null_t4())
        if emit:
            if temb is not None and self.temb_join_pending and lst is self.plan.fwd:
                lst.append((None, ()))       # the time-embedding lane joins before the first launch that reads its projections
                self.temb_join_pending = False
            lst.append((self.lib.dmu_conv2d, (C.byref(p),), lane) if lane else (self.lib.dmu_conv2d, (C.byref(p),)))
            self.plan.keep.append(p)
            if lst is self.plan.fwd and gn_coef is None and not lane:
                self.prod[(y.ptr, y.sw, Cj)] = p
        return p

    def _try_fuse(self, conv_p, gn_p, mode) -> int:
        """Ask the library whether `conv_p` can take the GroupNorm `gn_p` in its epilogue (mode 1 forward, 2 backward); leaves
        the conv parameters armed when it can.  Returns the launch's pixel-tile count (0 = no)."""
        if not (self.e.fuse_gn_epi and hasattr(self.lib, "dmu_conv2d_gn_fuse_supported")):
            return 0
        conv_p.gn_fuse = C.cast(C.pointer(gn_p), C.c_void_p)
        conv_p.gn_fuse_mode = mode
        tiles = int(self.lib.dmu_conv2d_gn_fuse_supported(C.byref(conv_p)))
        if tiles <= 0:
            conv_p.gn_fuse, conv_p.gn_fuse_mode = None, 0
            return 0
        return tiles

    def gn_conv(self, x: Buf, G, gamma_name, beta_name, silu, y4: Tensor4, w, w_strides, dims, geom, bias=None, temb=None, temb_pitch=0, res=None):
        """act(GroupNorm(x)) followed by a 3x3 convolution (residual.py:57-58,63-64; ddpm.py:88-90).  Where the persistent halo
        kernel takes the layer, the normalisation is applied to its operand tile in shared memory (statistics pass +
        coefficients + conv: the activation is neither written nor re-read, except as the weight-gradient operand when
        training); elsewhere a GroupNorm launch writes the activation first.  Returns (a, gn-record) like gn()."""
        fuse = False
        if self.e.fuse_gn and self.code == BF16 and hasattr(self.lib, "dmu_conv2d_gn_supported"):
            # the probe must not depend on this pass's addresses (the sizing pass runs at base 0): dummy aligned pointers
            d4 = lambda t: Tensor4(1 << 20, t.sn, t.sh, t.sw, t.sc, t.dtype, 0)
            probe = self.conv(None, d4(x.t4()), d4(y4), 1 << 20, w_strides, self.code, dims, geom, bias=(1 << 20) if bias else None,
                              temb=(1 << 20) if temb else None, temb_pitch=temb_pitch, res=d4(res) if res is not None else None, emit=False)
            fuse = self.lib.dmu_conv2d_gn_supported(C.byref(probe)) == 1
        if not fuse:
            a, rec = self.gn(x, G, gamma_name, beta_name, silu)
            self.conv(self.plan.fwd, a.t4(), y4, w, w_strides, self.code, dims, geom, bias=bias, temb=temb, temb_pitch=temb_pitch, res=res)
            return a, rec
        sums = self.stats.take(self.N * G * 2 * 4, 16)
        coef = self.f32(self.N * x.C * 2)
        gp = GnParams(x.t4(), _null_t4(), _null_t4(), _null_t4(), _null_t4(), sums, self.e.paddr(gamma_name), self.e.paddr(beta_name),
                      None, None, None, self.N, x.H, x.W, x.C, G, 1 if silu else 0, 1e-5, 0)
        self.plan.keep.append(gp)
        self.plan.fwd.append((self.lib.dmu_gn_stats, (C.byref(gp),)))
        self.plan.fwd.append((self.lib.dmu_gn_coef, (C.byref(gp), coef)))
        a = self.act(x.H, x.W, x.C) if self.train else None      # only the backward reads the activation itself
        self.conv(self.plan.fwd, x.t4(), y4, w, w_strides, self.code, dims, geom, bias=bias, temb=temb, temb_pitch=temb_pitch, res=res,
                  gn_coef=coef, gn_silu=1 if silu else 0, a_out=a.t4() if a is not None else None)
        return a, (x, a, sums, G, gamma_name, beta_name, silu)

    def wgrad(self, p4: Tensor4, q4: Tensor4, dw, dw_strides, dbias, dims, geom):
        N, Hp, Wp, Ca, Hq, Wq, Cb = dims
        R, S, stride, pad = geom
        batched = dbias is not None and self.side_lane and p4.sc == 1 and p4.dtype == self.code
        p = WgradParams(p4, q4, dw, dw_strides[0], dw_strides[1], dw_strides[2], None if batched else dbias, N, Hp, Wp, Ca, Hq, Wq, Cb, R, S, stride, pad,
                        self.e.impl)
        tag = 1 if self.side_lane else self.leaf_tag      # leaf_tag: where weight gradients go once the side lane is closed
        self.plan.bwd.append((self.lib.dmu_conv2d_wgrad, (C.byref(p),), tag) if tag else (self.lib.dmu_conv2d_wgrad, (C.byref(p),)))
        self.plan.keep.append(p)
        if batched:
            self.colsum(p4, N, Hp, Wp, Ca, None, 0, dbias)
        return p

    def gn(self, x: Buf, G, gamma_name, beta_name, silu: bool, out: Buf = None):
        """Forward GroupNorm(+SiLU): returns (y, gn-record)."""
        # [N][G][2] float sums, followed by the int64 fixed-point accumulators of the launches that add statistics with atomics
        # (GN_FIXED_SUMS: integer atomics commute, the bf16 forward is bit-identical from run to run)
        fixed = self.e.fixed_sums
        sums = self.stats.take(self.N * G * 2 * 4 * (3 if fixed else 1), 16)
        y = out if out is not None else self.act(x.H, x.W, x.C)
        p = GnParams(x.t4(), y.t4(), _null_t4(), _null_t4(), _null_t4(), sums, self.e.paddr(gamma_name), self.e.paddr(beta_name),
                     None, None, None, self.N, x.H, x.W, x.C, G, 1 if silu else 0, 1e-5, _abi.GN_FIXED_SUMS if fixed else 0)
        self.plan.keep.append(p)
        # the conv that wrote x (and nothing else has normalised yet) may carry this norm in its epilogue: no launch then
        prod = self.prod.pop((x.addr, x.pitch, x.C), None)
        if prod is not None and self._try_fuse(prod, p, 1):
            self.n_gn_fused[0] += 1
        elif prod is not None and self.e.fuse_gn_stats and self._try_fuse(prod, p, 3):
            # the large layers (persistent 3x3 kernel): the producer adds the raw sums in its epilogue, the norm is the apply pass alone
            self.n_gn_stats += 1
            self.plan.fwd.append((self.lib.dmu_gn_apply, (C.byref(p),)))
        else:
            self.plan.fwd.append((self.lib.dmu_gn_forward, (C.byref(p),)))
        return y, (x, y, sums, G, gamma_name, beta_name, silu)

    def gn_bwd(self, rec, dx: Buf, add0: Buf = None, add1: Buf = None, dgrad=None, after_dgrad=None):
        """Backward of gn(): reads rec.y.grad, writes dx (+ addends).  `dgrad` = emitter of the dgrad conv that PRODUCES rec.y.grad
        (called with no argument, returns its ConvParams without emitting): where the library can, the norm's backward runs in that
        conv's epilogue and rec.y.grad is never stored."""
        x, y, sums, G, gname, bname, silu = rec
        red = self.red.take(self.N * x.C * 2 * 4, 16)
        # dgamma/dbeta are left out here: one dmu_gn_param_grads launch folds every layer's per-image sums at the end
        p = GnParams(x.t4(), y.grad.t4(), dx.t4(), add0.t4() if add0 is not None else _null_t4(), add1.t4() if add1 is not None else _null_t4(),
                     sums, self.e.paddr(gname), self.e.paddr(bname), red, None, None,
                     self.N, x.H, x.W, x.C, G, 1 if silu else 0, 1e-5, 0)
        self.plan.keep.append(p)
        tiles = 0
        if dgrad is not None:
            cp = dgrad()
            tiles = self._try_fuse(cp, p, 2)
            self.plan.bwd.append((self.lib.dmu_conv2d, (C.byref(cp),)))
            self.plan.keep.append(cp)
        if after_dgrad is not None:
            after_dgrad()          # the weight gradient keeps its place right behind the dgrad (side lane)
        # fused: red holds per-tile channel sums ([tiles][C][2]) instead of per-image ones
        self.gn_pg.append(GnPgDesc(red, self.gp(gname), self.gp(bname), x.C, tiles))
        if tiles:
            self.n_gn_fused[1] += 1
        else:
            self.plan.bwd.append((self.lib.dmu_gn_backward, (C.byref(p),)))

    # ---- conv layer helpers (filters repacked [O][R][S][I])
    def conv_layer(self, x: Buf, y: Buf, wname, bname, R, stride, pad, temb=None, temb_pitch=0, res: Buf = None, lane=0):
        Ci, Co = x.C, y.C
        self.conv(self.plan.fwd, x.t4(), y.t4(), self.e.waddr(wname), (R * R * Ci, 1, Ci), self.code,
                  (self.N, x.H, x.W, Ci, y.H, y.W, Co), (R, R, stride, pad), bias=self.e.paddr(bname), temb=temb, temb_pitch=temb_pitch,
                  res=res.t4() if res is not None else None, lane=lane)

    def conv_layer_bwd(self, x: Buf, y: Buf, wname, bname, R, stride, pad, dx: Buf, need_dx=True, gn=None, lane=0):
        """dgrad into dx (plain write) and wgrad/dbias into the arena.  dy = y.grad.  gn = (rec, out, add0, add1): x is the output
        of the GroupNorm `rec`, whose backward follows immediately (into `out`) - in the dgrad's epilogue where possible."""
        Ci, Co = x.C, y.C
        dy = y.grad
        # dx[n,hi,wi,ci] = sum dy[n,(hi+pad-r)/s,..,co] w[co][r][s][ci]
        dgrad = lambda emit: self.conv(self.plan.bwd, dy.t4(), dx.t4(), self.e.waddr_t(wname), (R * R * Co, 1, Co), self.code,
                                       (self.N, y.H, y.W, Co, x.H, x.W, Ci), (R, R, stride, pad), gather=1, emit=emit, lane=lane)
        wgrad = lambda: self.wgrad(dy.t4(), x.t4(), self.e.gsaddr(wname), (R * R * Ci, 1, Ci), self.gp(bname),
                                   (self.N, y.H, y.W, Co, x.H, x.W, Ci), (R, R, stride, pad))
        if gn is not None:
            rec, out, add0, add1 = gn
            self.gn_bwd(rec, out, add0=add0, add1=add1, dgrad=lambda: dgrad(False), after_dgrad=wgrad)
            return
        if need_dx:
            dgrad(True)
        wgrad()

    def linear(self, lst, x4: Tensor4, y4: Tensor4, M, I, O, w_addr, b_addr, res4=None, w_code=F32):
        self.conv(lst, x4, y4, w_addr, (I, 1, 0), w_code, (M, 1, 1, I, 1, 1, O), (1, 1, 1, 0), bias=b_addr, res=res4)

    def linear_bwd(self, x4, dy4, dx4, M, I, O, wname_addr, gw_addr, gb_addr, res4=None, need_dx=True, w_t=None):
        """w_t: address of the [I][O] copy in the compute dtype (contraction axis contiguous: tensor-core path);
        None = contract the fp32 [O][I] parameter through its strides."""
        if need_dx and w_t is not None:
            self.conv(self.plan.bwd, dy4, dx4, w_t, (O, 1, 0), self.code, (M, 1, 1, O, 1, 1, I), (1, 1, 1, 0), res=res4)
        elif need_dx:
            self.conv(self.plan.bwd, dy4, dx4, wname_addr, (1, I, 0), F32, (M, 1, 1, O, 1, 1, I), (1, 1, 1, 0), res=res4)
        self.wgrad(dy4, x4, gw_addr, (I, 1, 0), gb_addr, (M, 1, 1, O, 1, 1, I), (1, 1, 1, 0))

    # ---- blocks
    def res_block(self, pfx, x: Buf, y: Buf, tp_addr, tp_pitch, extra_add: Buf = None):
        """Forward ops of residual.py:54-68 writing into y; registers the backward emitter.
        extra_add: an already-written gradient of x from another consumer (skip connection)."""
        e = self.e
        Ci, Co = x.C, y.C
        h = self.act(x.H, x.W, Co)
        has_sc = Ci != Co
        aux = 2 if (has_sc and e.aux_lanes) else 0
        if aux:     # the 1x1 shortcut only needs x: it runs on the auxiliary lane next to norm1 -> conv1 -> norm2 and rejoins before conv2
            sc = self.tmp(x.H, x.W, Co)
            self.conv_layer(x, sc, pfx + "shortcut.weight", pfx + "shortcut.bias", 1, 1, 0, lane=aux)
        a1, rec1 = self.gn_conv(x, gn_groups(Ci), pfx + "norm1.weight", pfx + "norm1.bias", True, h.t4(), e.waddr(pfx + "conv1.weight"),
                                (9 * Ci, 1, Ci), (self.N, x.H, x.W, Ci, h.H, h.W, Co), (3, 3, 1, 1), bias=e.paddr(pfx + "conv1.bias"),
                                temb=tp_addr, temb_pitch=tp_pitch)
        if has_sc and not aux:
            sc = self.tmp(x.H, x.W, Co)
            self.conv_layer(x, sc, pfx + "shortcut.weight", pfx + "shortcut.bias", 1, 1, 0)
        if aux:
            self.plan.fwd.append((None, (), 2))
        a2, rec2 = self.gn_conv(h, gn_groups(Co), pfx + "norm2.weight", pfx + "norm2.bias", True, y.t4(), e.waddr(pfx + "conv2.weight"),
                                (9 * Co, 1, Co), (self.N, h.H, h.W, Co, y.H, y.W, Co), (3, 3, 1, 1), bias=e.paddr(pfx + "conv2.bias"),
                                res=(sc if has_sc else x).t4())
        if not self.train:
            return

        def bwd():
            # shortcut: its input gradient is an addend of norm1's backward; the dgrad only needs dy, so it runs on the auxiliary
            # lane next to conv2 / norm2 / conv1 and rejoins before norm1's backward
            if has_sc:
                dxs = self.tmp(x.H, x.W, Ci)
                scb = Buf(0, self.N, x.H, x.W, Co, Co, self.code)
                scb.grad = y.grad
                if aux:
                    self.conv_layer_bwd(x, scb, pfx + "shortcut.weight", pfx + "shortcut.bias", 1, 1, 0, dxs, lane=aux)
            # conv2, then norm2 + silu -> dh (in the dgrad's epilogue where one tile holds whole images)
            self.conv_layer_bwd(a2, y, pfx + "conv2.weight", pfx + "conv2.bias", 3, 1, 1, a2.grad, gn=(rec2, h.grad, None, None))
            # time projection: per-image channel sums of dh
            self.colsum(h.grad.t4(), self.N, h.H, h.W, Co, self.dtproj + self.tp_off[pfx] * 4, self.tp_total, None)
            if has_sc and not aux:
                self.conv_layer_bwd(x, scb, pfx + "shortcut.weight", pfx + "shortcut.bias", 1, 1, 0, dxs)
            if not has_sc:
                dxs = y.grad
            if aux:
                self.plan.bwd.append((None, (), 2))
            prev = x.grad if x.grad_written else None
            # conv1, then norm1 + silu (+ shortcut / skip gradients) -> dx
            self.conv_layer_bwd(a1, h, pfx + "conv1.weight", pfx + "conv1.bias", 3, 1, 1, a1.grad, gn=(rec1, x.grad, dxs, prev))
            x.grad_written = True
        self.tape.append(bwd)

    def attn_block(self, pfx, x: Buf, y: Buf, heads=4):
        """attention.py:36-68 writing GN(proj + x) into y."""
        e = self.e
        Cc, S = x.C, x.H * x.W
        M = self.N * S
        qkv = self.act(x.H, x.W, 3 * Cc)
        bq = e.paddr(pfx + "query_projection.bias")
        tc = self.code == BF16      # projections through the filter cache (tensor-core path) in bf16 mode, fp32 arena otherwise
        wq, wq_t, wcode = (e.waddr(pfx + "qkv"), e.waddr_t(pfx + "qkv"), BF16) if tc else (e.paddr(pfx + "query_projection.weight"), None, F32)
        self.linear(self.plan.fwd, _rows_t4(x.addr, x.pitch, self.code), _rows_t4(qkv.addr, 3 * Cc, self.code), M, Cc, 3 * Cc, wq, bq, w_code=wcode)
        o = self.act(x.H, x.W, Cc)
        lse = self.f32(self.N * heads * S)
        ap = AttnParams(qkv.addr, 3 * Cc, o.addr, Cc, None, 0, None, 0, lse, self.N, S, Cc, heads, self.code, 0)
        self.plan.keep.append(ap)
        self.plan.fwd.append((self.lib.dmu_attn_fwd, (C.byref(ap),)))
        z = self.act(x.H, x.W, Cc)
        bf = e.paddr(pfx + "final_projection.bias")
        wf, wf_t = (e.waddr(pfx + "final_projection.weight"), e.waddr_t(pfx + "final_projection.weight")) if tc else (e.paddr(pfx + "final_projection.weight"), None)
        # final projection (+ x) as a 1x1 convolution over (N, H, W) - the same rows as the [M, C] Linear - so that an output tile
        # holds whole images and the post-norm of attention.py:68 can ride in its epilogue
        self.conv(self.plan.fwd, o.t4(), z.t4(), wf, (Cc, 1, 0), wcode, (self.N, x.H, x.W, Cc, x.H, x.W, Cc), (1, 1, 1, 0), bias=bf, res=x.t4())
        _, rec = self.gn(z, gn_groups(Cc), pfx + "norm.weight", pfx + "norm.bias", False, out=y)
        if not self.train:
            return

        def bwd():
            self.gn_bwd(rec, z.grad)
            dz4 = _rows_t4(z.grad.addr, Cc, self.code)
            # final projection: do = dz Wf ; dWf += dz^T o
            self.linear_bwd(_rows_t4(o.addr, Cc, self.code), dz4, _rows_t4(o.grad.addr, Cc, self.code), M, Cc, Cc, wf,
                            self.gp(pfx + "final_projection.weight"), self.gp(pfx + "final_projection.bias"), w_t=wf_t)
            bp = AttnParams(qkv.addr, 3 * Cc, o.addr, Cc, o.grad.addr, Cc, qkv.grad.addr, 3 * Cc, lse, self.N, S, Cc, heads, self.code, 0)
            self.plan.keep.append(bp)
            self.plan.bwd.append((self.lib.dmu_attn_bwd, (C.byref(bp),)))
            # qkv projection: dx = dqkv Wqkv + dz ; dWqkv += dqkv^T x
            self.linear_bwd(_rows_t4(x.addr, x.pitch, self.code), _rows_t4(qkv.grad.addr, 3 * Cc, self.code),
                            _rows_t4(x.grad.addr, x.grad.pitch, self.code), M, Cc, 3 * Cc, wq,
                            self.gp(pfx + "query_projection.weight"), self.gp(pfx + "query_projection.bias"), res4=dz4, w_t=wq_t)
            x.grad_written = True
        self.tape.append(bwd)

    def stage(self, pfx, has_attn, x: Buf, Co) -> Buf:
        """Two (ResBlock [, Attention]) pairs; returns the stage output before down/up-sampling."""
        h = x
        for j in range(2):
            y = self.act(x.H, x.W, Co)
            self.res_block(f"{pfx}res_blocks.{j}.", h, y, self.tproj + self.tp_off[f"{pfx}res_blocks.{j}."] * 4, self.tp_total)
            h = y
            if has_attn:
                y2 = self.act(x.H, x.W, Co)
                self.attn_block(f"{pfx}attention_blocks.{j}.", h, y2)
                h = y2
        return h

    # ---- whole network
    def build(self) -> Plan:
        e, net, N, H, W = self.e, self.e.net, self.N, self.H, self.W
        Cm = net.model_channels
        plan = self.plan
        plan.keep = []
        lib = self.lib
        T4 = 4 * Cm
        # time-projection column offsets (arena order == res_block_prefixes order)
        self.tp_off, off = {}, 0
        for p in e.res_block_prefixes():
            self.tp_off[p] = off
            off += e.named[p + "time_mlp.weight"].shape[0]
        self.tp_total = off

        # -------- time embedding
        t_ptr = plan.t_in.data_ptr() if plan.t_in is not None else 0   # dry (sizing) pass has no buffers yet
        if not net.sigma_embed:
            emb = self.f32(N * Cm)
            plan.fwd.append((lib.dmu_sinusoidal_embedding, (t_ptr, 1, emb, N, Cm)))
            h1 = self.f32(N * T4)
            te = "time_embedding.positional_encoding."
            self.linear(plan.fwd, _rows_t4(emb, Cm), _rows_t4(h1, T4), N, Cm, T4, e.paddr(te + "1.weight"), e.paddr(te + "1.bias"))
            g1 = self.f32(N * T4)
            plan.fwd.append((lib.dmu_act_fwd, (h1, g1, N * T4, 0)))
            temb = self.f32(N * T4)
            self.linear(plan.fwd, _rows_t4(g1, T4), _rows_t4(temb, T4), N, T4, T4, e.paddr(te + "3.weight"), e.paddr(te + "3.bias"))
        else:
            # score_based.py:57-61,82-83: Linear(1,C) -> SiLU -> Linear(C,4C) on log(sigma)
            ls = self.f32(N)
            plan.fwd.append((lib.dmu_act_fwd, (t_ptr, ls, N, 2)))
            h1 = self.f32(N * Cm)
            self.linear(plan.fwd, _rows_t4(ls, 1), _rows_t4(h1, Cm), N, 1, Cm, e.paddr("time_embed.0.weight"), e.paddr("time_embed.0.bias"))
            g1 = self.f32(N * Cm)
            plan.fwd.append((lib.dmu_act_fwd, (h1, g1, N * Cm, 1)))
            temb = self.f32(N * T4)
            self.linear(plan.fwd, _rows_t4(g1, Cm), _rows_t4(temb, T4), N, Cm, T4, e.paddr("time_embed.2.weight"), e.paddr("time_embed.2.bias"))
        self.tproj = self.f32(N * self.tp_total)
        first = e.res_block_prefixes()[0]
        self.linear(plan.fwd, _rows_t4(temb, T4), _rows_t4(self.tproj, self.tp_total), N, T4, self.tp_total,
                    e.paddr(first + "time_mlp.weight"), e.paddr(first + "time_mlp.bias"))
        # Everything emitted so far (embedding MLP + the 22-way projection: ~5 small launches) only feeds the `+ temb` of the
        # first ResBlock's conv1: it runs on the side lane of the forward graph, next to the stem conv and the first GroupNorm.
        plan.fwd[:] = [op + (1,) for op in plan.fwd]
        self.temb_join_pending = True
        # per-image channel sums of dh (time-projection gradient): accumulated with atomics -> lives in the zeroed region
        self.dtproj = self.red.take(N * self.tp_total * 4, 16) if self.train else 0

        # -------- stem (NCHW fp32 -> NHWC)
        h0 = self.act(H, W, Cm)
        x_ptr = plan.x_in.data_ptr() if plan.x_in is not None else 0
        out_ptr = plan.out.data_ptr() if plan.out is not None else 0
        dout_ptr = plan.dout.data_ptr() if plan.dout is not None else 0
        self.conv(plan.fwd, _nchw_t4(x_ptr, net.in_channels, H, W), h0.t4(), e.waddr("initial_conv.weight"),
                              (9 * net.in_channels, 1, net.in_channels), self.code, (N, H, W, net.in_channels, H, W, Cm), (3, 3, 1, 1),
                              bias=e.paddr("initial_conv.bias"))

        # -------- concat buffers of the up path: cat_k = [h part | skip part]
        dplan, uplan = down_plan(Cm), up_plan(Cm)
        cat = []
        for k, (_, cin, _) in enumerate(uplan):
            r = 1 << k  # spatial size factor: cat_0 at H/32
            cat.append(self.act(H // 32 * r, W // 32 * r, cin))
        # -------- down path
        x = h0
        skips = []
        down3_tape_start = 0
        for i, (attn, ci, co) in enumerate(dplan):
            pfx = f"down_blocks.{i}."
            if i == 3:
                down3_tape_start = len(self.tape)
                plan.fwd.append(("zero_here", ()))     # Engine._build: where a training step may hang the backward's memsets (latency-bound stages follow)
            y = self.stage(pfx, attn, x, co)
            kcat = 4 - i
            hpart = uplan[kcat][1] - co
            d = cat[kcat].slice(hpart, co)
            self.conv_layer(y, d, pfx + "downsample.weight", pfx + "downsample.bias", 4, 2, 1)
            if self.train:
                def down_bwd(y=y, d=d, pfx=pfx):
                    self.conv_layer_bwd(y, d, pfx + "downsample.weight", pfx + "downsample.bias", 4, 2, 1, y.grad)
                    y.grad_written = True
                self.tape.append(down_bwd)
            skips.append(d)
            x = d
        # -------- bottleneck (writes into the h part of cat_0)
        b0 = self.act(x.H, x.W, 4 * Cm)
        self.res_block("bottleneck.0.", x, b0, self.tproj + self.tp_off["bottleneck.0."] * 4, self.tp_total)
        b1 = self.act(x.H, x.W, 4 * Cm)
        self.attn_block("bottleneck.1.", b0, b1)
        b2 = cat[0].slice(0, 4 * Cm)
        self.res_block("bottleneck.2.", b1, b2, self.tproj + self.tp_off["bottleneck.2."] * 4, self.tp_total)
        # -------- up path
        up_tape_start = len(self.tape)
        for k, (attn, ci, co) in enumerate(uplan):
            pfx = f"up_blocks.{k}."
            y = self.stage(pfx, attn, cat[k], co)
            if k < 4:
                u = cat[k + 1].slice(0, co)
            else:
                u = self.act(H, W, co)
            # ConvTranspose2d 4x4 s2 p1 (residual.py:121,242): transposed gather, filters repacked [O][R][S][I]
            self.conv(plan.fwd, y.t4(), u.t4(), e.waddr(pfx + "upsample.weight"), (16 * co, 1, co), self.code,
                      (N, y.H, y.W, co, u.H, u.W, co), (4, 4, 2, 1), bias=e.paddr(pfx + "upsample.bias"), gather=1)
            if self.train:
                def up_bwd(y=y, u=u, pfx=pfx, co=co):
                    du = u.grad
                    # dgrad of a transposed conv is a strided conv over du
                    self.conv(plan.bwd, du.t4(), y.grad.t4(), e.waddr_t(pfx + "upsample.weight"), (16 * co, 1, co), self.code,
                              (N, u.H, u.W, co, y.H, y.W, co), (4, 4, 2, 1), gather=0)
                    y.grad_written = True
                    # dW[ci][co][r][s] (IOHW) += x[.., ci] * du[gathered, co]
                    self.wgrad(y.t4(), du.t4(), e.gsaddr(pfx + "upsample.weight"), (16 * co, 1, co), None,
                               (N, y.H, y.W, co, u.H, u.W, co), (4, 4, 2, 1))
                    self.colsum(du.t4(), N, u.H, u.W, co, None, 0, self.gp(pfx + "upsample.bias"))
                self.tape.append(up_bwd)
            x = u
        # -------- head: GroupNorm -> SiLU -> conv3x3 -> NCHW fp32
        a, rec = self.gn_conv(x, 32, "output_conv.0.weight", "output_conv.0.bias", True, _nchw_t4(out_ptr, net.out_channels, H, W),
                              e.waddr("output_conv.2.weight"), (9 * Cm, 1, Cm), (N, H, W, Cm, H, W, net.out_channels), (3, 3, 1, 1),
                              bias=e.paddr("output_conv.2.bias"))
        # one launch zeroes every GroupNorm statistics accumulator of the forward
        plan.fwd.insert(0, (lib.dmu_zero, (self.stats.base, max(self.stats.off, 4))))

        if not self.train:
            return plan

        # ======================= backward =======================
        Co = net.out_channels
        dout4 = _nchw_t4(dout_ptr, Co, H, W)
        # head conv: da = dgrad(dout), dW, db
        self.conv(plan.bwd, dout4, a.grad.t4(), e.waddr("output_conv.2.weight"), (1, 9 * Cm, Cm), self.code,
                  (N, H, W, Co, H, W, Cm), (3, 3, 1, 1), gather=1)
        self.wgrad(_nchw_t4(dout_ptr, Co, H, W), a.t4(), e.gsaddr("output_conv.2.weight"), (9 * Cm, 1, Cm), self.gp("output_conv.2.bias"),
                   (N, H, W, Co, H, W, Cm), (3, 3, 1, 1))
        self.gn_bwd(rec, x.grad)
        x.grad_written = True
        self.gn_pg_split = []
        for i in range(len(self.tape) - 1, -1, -1):
            if i == up_tape_start - 1 or i == down3_tape_start - 1:
                # Everything emitted so far is the backward of (head + up path) or of (... + bottleneck + down 4, 3): the
                # parameter gradients of those layers are final here (except their time projections and q/k/v, which live at
                # the head of the arena), which is where a data-parallel step can start all-reducing them.
                self.flush_colsums()
                plan.bwd.append(("split", ()))
                self.gn_pg_split.append(len(self.gn_pg))
            self.tape[i]()
        # time projections (one GEMM for all 22 blocks), then the embedding MLP: they consume the per-image column sums the
        # side lane produced, so the lanes join here - BEFORE the stem's weight gradient (wgrad only: the network input needs
        # no gradient on this path) is queued on the side lane: it runs next to the embedding backward instead of in front of it
        self.colsum(h0.grad.t4(), N, H, W, Cm, None, 0, self.gp("initial_conv.bias"))
        self.flush_colsums()
        plan.bwd.append((None, (), 3) if e.aux_lanes else (None, ()))      # the column sums are in: join their chain
        self.wgrad(h0.grad.t4(), _nchw_t4(x_ptr, net.in_channels, H, W), e.gsaddr("initial_conv.weight"),
                   (9 * net.in_channels, 1, net.in_channels), None, (N, H, W, Cm, H, W, net.in_channels), (3, 3, 1, 1))
        plan.bwd.append(("tails", ()))      # Engine._build puts the last part's tail (parameter-gradient fold, staging unpack) here, on the side lane
        self.side_lane = False
        # the embedding backward is a chain dtemb -> dg1 -> dh1 with three weight gradients hanging off it: those leaves go to the
        # (now idle) column-sum chain instead of standing in the chain
        self.leaf_tag = 3 if e.aux_lanes else 0
        dtemb = self.f32(N * T4)
        self.linear_bwd(_rows_t4(temb, T4), _rows_t4(self.dtproj, self.tp_total), _rows_t4(dtemb, T4), N, T4, self.tp_total,
                        e.paddr(first + "time_mlp.weight"), self.gp(first + "time_mlp.weight"), self.gp(first + "time_mlp.bias"))
        if not net.sigma_embed:
            te = "time_embedding.positional_encoding."
            dg1 = self.f32(N * T4)
            self.linear_bwd(_rows_t4(g1, T4), _rows_t4(dtemb, T4), _rows_t4(dg1, T4), N, T4, T4, e.paddr(te + "3.weight"),
                            self.gp(te + "3.weight"), self.gp(te + "3.bias"))
            dh1 = self.f32(N * T4)
            plan.bwd.append((lib.dmu_act_bwd, (h1, dg1, dh1, N * T4, 0)))
            self.linear_bwd(_rows_t4(emb, Cm), _rows_t4(dh1, T4), None, N, Cm, T4, e.paddr(te + "1.weight"),
                            self.gp(te + "1.weight"), self.gp(te + "1.bias"), need_dx=False)
        else:
            dg1 = self.f32(N * Cm)
            self.linear_bwd(_rows_t4(g1, Cm), _rows_t4(dtemb, T4), _rows_t4(dg1, Cm), N, Cm, T4, e.paddr("time_embed.2.weight"),
                            self.gp("time_embed.2.weight"), self.gp("time_embed.2.bias"))
            dh1 = self.f32(N * Cm)
            plan.bwd.append((lib.dmu_act_bwd, (h1, dg1, dh1, N * Cm, 1)))
            self.linear_bwd(_rows_t4(ls, 1), _rows_t4(dh1, Cm), None, N, 1, Cm, e.paddr("time_embed.0.weight"),
                            self.gp("time_embed.0.weight"), self.gp("time_embed.0.bias"), need_dx=False)
        # one launch zeroes every GroupNorm backward accumulator (issued with the gradient-arena memsets: Engine._build)
        plan.zero_red = (lib.dmu_zero, (self.red.base, max(self.red.off, 4)))
        plan.bwd.append((None, ()))      # side lane joins
        # (zeroing of the gradient arena / staging, the GroupNorm parameter-gradient fold and the staging unpack are
        #  emitted once for all micro-batch lanes by Engine._build)
        return plan
