"""One optimisation step of the reference trainer's hot loop on the CUDA engine.

Mirror of trainers/ddpm_trainer.py:539-555 (``images.to(device)`` ->
``zero_grad`` -> ``loss_function`` -> ``backward`` -> ``optimizer.step`` ->
``_update_ema_model``) with the three host-side costs removed:
  * gradients live in one flat arena, averaged across ranks by bucketed
    asynchronous all-reduces (the reference's DDP wrapper never reduces:
    trainers/ddpm_trainer.py:130-136 vs :543-547);
  * Adam over 314 tensors + the 314-iteration Python EMA loop become one
    ``dmu_adam_ema`` launch over the arena;
  * the loss stays on the device — callers read it when they want it.
"""

from collections import OrderedDict
from typing import Optional

import os

import torch
import torch.distributed as dist

from . import ops
from .losses import DiffusionLoss
from .optim import FusedAdamEMA
from .parallel import GradAllReducer


class _FlatParams:
    """Flat fp32 parameter / gradient arenas for a network that has no launch-plan engine of its own (EnergyNet,
    models/energy_based.py:51-85): the parameters become views of ``flat`` and their ``.grad`` views of ``gflat`` (autograd
    accumulates into an installed ``.grad`` in place), which is all FusedAdamEMA and GradAllReducer need of an engine."""

    def __init__(self, module):
        self.module = module
        self.flat = self.gflat = None
        self.named = OrderedDict(module.named_parameters())
        self.offs = {}

    def _in_arena(self):
        if self.flat is None:
            return False
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * 4
        return all(p.device == self.flat.device and lo <= p.data_ptr() < hi for p in self.named.values())

    def prepare(self, device=None):
        if self._in_arena():
            return
        device = device if device is not None else next(iter(self.named.values())).device
        offs, total = {}, 0
        for k, p in self.named.items():
            offs[k] = (total, p.numel())
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        with torch.no_grad():
            for k, p in self.named.items():
                o, n = offs[k]
                flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = flat[o:o + n].view(p.shape)
        self.flat, self.offs = flat, offs
        self.gflat = torch.zeros_like(flat)

    def install_grads(self):
        """Zero the gradient arena and make every ``param.grad`` a view of it."""
        self.prepare()
        self.gflat.zero_()
        for k, p in self.named.items():
            o, n = self.offs[k]
            p.grad = self.gflat[o:o + n].view(p.shape)


class _EngineShim:
    """What FusedAdamEMA / GradAllReducer take as ``unet``: an object with ``.engine`` and the module's parameter iterators."""

    def __init__(self, module):
        self.engine = _FlatParams(module)
        self.named_parameters = module.named_parameters
        self.parameters = module.parameters


class TrainStep:
    def __init__(self, model, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_decay: Optional[float] = 0.9999, bucket_mb: float = 16.0, group=None,
                 input_norm=None, input_layout: str = "NCHW"):
        """input_norm = (mean, std) per channel and input_layout ("NCHW" | "NHWC") describe uint8 batches: ``step`` then
        takes the decoded image bytes and runs ToTensor + Normalize (datasets/dataset_utils.py:58-61) on the device, fused
        into q_sample.  fp32 batches are used as they are."""
        self.model = model
        if input_layout not in ("NCHW", "NHWC"):
            raise ValueError(f"input_layout must be 'NCHW' or 'NHWC', got {input_layout!r}")
        self.input_layout = input_layout
        self._norm_host = None if input_norm is None else tuple(torch.as_tensor(v, dtype=torch.float32).flatten() for v in input_norm)
        self._norm_dev = (None, None)
        # the UNet-based models (DDPM / DDIM / score) own an Engine with flat arenas; any other network (EnergyNet) gets them here
        self._net = model.model if hasattr(model.model, "engine") else _EngineShim(model.model)
        self._own_arena = not hasattr(model.model, "engine")
        self.opt = FusedAdamEMA(self._net, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay)
        self.reducer = GradAllReducer(self._net, bucket_mb=bucket_mb, group=group)
        self.group = group
        self._synced = False
        self._stage = None
        self._works = None
        # Single-GPU DDPM steps replay ONE CUDA graph of everything between the input batch and the gradient arena (RNG draws,
        # time weights, q_sample, forward, loss, three-part backward): ~30 small eager launches and four graph launches per step
        # become one.  DMU_STEP_GRAPH=0 keeps the piecewise path.
        self._use_step_graph = os.environ.get("DMU_STEP_GRAPH", "1") != "0"
        # (Capturing the NCCL all-reduces into the step graph was measured at 2 GPUs in round 2: 84.3k vs 84.5k img/s, and the
        # process hung in teardown - the piecewise form below is what runs with several ranks.)
        self._graph = None
        self._g_in = self._g_loss = self._g_dpred = self._g_plan = None
        self._g_warm = 0
        self._g_launches = 0
        self._g_arena = 0

    def step(self, images: torch.Tensor) -> torch.Tensor:
        """images: fp32 [B,C,H,W] on the device, or a (pinned) host tensor which is
        copied asynchronously first (trainers/ddpm_trainer.py:539); or the uint8 image bytes
        ([B,C,H,W] / [B,H,W,C] per ``input_layout``), normalised on the device.  Returns the
        0-dim device loss tensor of this rank (no host sync)."""
        if images.dtype not in (torch.float32, torch.uint8):
            raise TypeError(f"images must be float32 or uint8, got {images.dtype}")
        if not images.is_cuda:
            dev = next(self.model.parameters()).device
            if self._stage is None or self._stage.shape != images.shape or self._stage.dtype != images.dtype:
                self._stage = torch.empty(images.shape, device=dev, dtype=images.dtype)
            self._stage.copy_(images, non_blocking=True)
            images = self._stage
        m = self.model
        ddpm_like = isinstance(getattr(m, "loss_fn", None), DiffusionLoss) and hasattr(m, "alphas_cumprod") and not self._own_arena
        if not self._synced:
            self.sync_parameters(images.device if images.is_cuda else next(m.parameters()).device)
        if images.dtype == torch.uint8:
            if self._norm_host is not None and (self._norm_dev[0] is None or self._norm_dev[0].device != images.device):
                self._norm_dev = tuple(v.to(images.device) for v in self._norm_host)
            if not ddpm_like:      # score / energy variants take the normalised batch
                images, _ = ops.ingest_u8(images, self._norm_dev[0], self._norm_dev[1], self.input_layout)
        if ddpm_like:     # gradient exchange and optimizer update run inside, range by range, while the backward continues
            if self._use_step_graph and images.is_cuda and m.model.engine.use_graphs:
                return self._ddpm_step_graphed(images).detach()
            return self._ddpm_step(images).detach()
        elif self._own_arena:   # EnergyNet: autograd accumulates straight into the views of the gradient arena
            self._net.engine.install_grads()
            loss = m.loss_function(images)
            loss.backward()
        else:   # generic route through autograd (score variant): _UNetFn.backward fills the engine's gradient arena
            loss = m.loss_function(images)
            loss.backward()
            for p in m.parameters():      # the arena holds this step's gradients; views must not accumulate into the next
                p.grad = None
        if self._works is None:
            self._works = self.reducer.launch()
        scale = self.reducer.finish(self._works)
        self._works = None
        self.opt.step(grad_scale=scale)
        return loss.detach()

    def sync_parameters(self, device=None):
        """What DDP does at construction (trainers/ddpm_trainer.py:130-136): every rank starts from rank 0's parameters; here
        also the optimizer's moment and EMA arenas when they exist (after ``load_checkpoint``).  No-op on one rank."""
        eng = self._net.engine
        eng.prepare(device if device is not None else next(self.model.parameters()).device)
        self._synced = True
        if self.reducer.world == 1:
            return
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        for t in (eng.flat, self.opt.m, self.opt.v, self.opt.ema):
            if t is not None:
                dist.broadcast(t, src=src, group=self.group)
        steps = torch.tensor([self.opt.step_count], device=eng.flat.device, dtype=torch.int64)
        dist.broadcast(steps, src=src, group=self.group)
        self.opt.step_count = int(steps.item())

    def _ddpm_step_graphed(self, images: torch.Tensor) -> torch.Tensor:
        eng = self.model.model.engine
        arena = eng.flat.data_ptr() if eng.flat is not None else 0
        if self._g_in is None or self._g_in.shape != images.shape or self._g_in.dtype != images.dtype \
                or self._g_in.device != images.device or arena != self._g_arena:
            # new batch shape / device, or the engine rebuilt its parameter arena (model.to(), first call): the captured graph
            # holds the old addresses
            # a host batch has already been staged into self._stage: that buffer is the graph's static input (no second copy)
            self._g_in = images if images is self._stage else torch.empty_like(images)
            self._graph, self._g_warm, self._g_arena = None, 0, arena
        if images is not self._g_in:
            self._g_in.copy_(images)
        if self._graph is None:
            if self._g_warm < 3:               # eager steps first: plans, lazy one-time initialisation, allocator warm-up
                self._g_warm += 1
                return self._ddpm_step(self._g_in)
            g = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(g):
                if self.reducer.world == 1:    # front, backward and the three range updates: one graph
                    self._g_loss = self._ddpm_step(self._g_in, capturing=True)
                else:      # data parallel: the graph ends at dL/d(eps); the three backward graphs alternate with the all-reduces
                    self._g_loss, self._g_dpred, self._g_plan = self._ddpm_front(self._g_in)
            self._g_launches = ops.LAUNCHES - n0
            ops.LAUNCHES = n0
            self._graph = g
        self._graph.replay()
        ops.LAUNCHES += self._g_launches
        if self.reducer.world > 1:
            self.opt.begin_step()
            self._ddpm_back(self._g_plan)
        else:
            self.opt.step_count += 1       # host mirror of the device step count the replayed graph advanced
        # the graph's static output is overwritten by the next replay: hand out a copy (one 4-byte device copy, no sync)
        return self._g_loss.clone()

    def _ddpm_step(self, images: torch.Tensor, capturing: bool = False) -> torch.Tensor:
        """``DDPM.loss_function`` + ``backward`` (models/ddpm.py:207-235) straight on the engine: same RNG calls in the same
        order and the same launches, without the autograd graph (314 AccumulateGrad nodes cost more host time than the
        GPU needs for the whole backward pass).  Gradients land in the engine's flat arena; every third of it is exchanged and
        fed to the fused Adam + EMA as soon as its part of the backward has finished."""
        loss, _, plan = self._ddpm_front(images)
        self.opt.begin_step(count_host=not capturing)
        self._ddpm_back(plan)
        return loss

    def _ddpm_front(self, images: torch.Tensor):
        """RNG draws, time weights, q_sample, forward, loss and dL/d(eps).  No staging copies: q_sample writes the network's
        static input buffer, the loss reads its static output and writes the static upstream-gradient buffer; the refresh of
        the bf16 filter caches (dmu_repack_weights, ~60 us) comes first; the memsets the backward needs ride beside the forward's
        latency-bound stages (Engine.run_forward(zero_backward=True))."""
        m = self.model
        eng = m.model.engine
        eng.prepare(images.device)
        if images.dtype == torch.uint8:
            b, d1, d2, d3 = images.shape
            shape = (b, d3, d1, d2) if self.input_layout == "NHWC" else (b, d1, d2, d3)
        else:
            shape = tuple(images.shape)
        plan = eng.get_plan(shape, True)
        # The refresh of the bf16 filter caches leads the step on the same stream as everything else.  (Measured in round 2 inside the
        # step graph, 2.645 - 2.657 ms whichever way: the refresh on a side branch next to the draws, the draws on a high-priority
        # branch next to the refresh - the 9344-CTA permutation fills the machine, a concurrent branch only queues behind it; and the
        # refresh split by consumer - the forward's [O][R][S][I] copies first, the backward's [I][R][S][O] copies on a second stream
        # beside the forward - LOSES 1.1 %: 48.04k against 48.58k img/s, the co-running half slows the forward by more than it saves.)
        if not eng.frozen:
            eng.repack(ops._stream())
        t = torch.randint(0, m.num_timesteps, (shape[0],), device=images.device)
        noise = torch.randn(shape, device=images.device, dtype=torch.float32) if images.dtype == torch.uint8 else torch.randn_like(images)
        w = m.loss_fn.time_weights(t)          # [B]-sized, issued before the forward so nothing waits on it later
        if images.dtype == torch.uint8:        # decoded bytes: ToTensor + Normalize + q_sample in one launch
            ops.ingest_u8(images, self._norm_dev[0], self._norm_dev[1], self.input_layout, t, noise, m.alphas_cumprod,
                          want_x0=False, xt_out=plan.x_in)
        else:
            ops.q_sample(images.contiguous(), t, noise, m.alphas_cumprod, out=plan.x_in)
        eps = eng.run_forward(None, t, plan, repacked=True, clone=False, zero_backward=True)
        wm, wl, wh = m.loss_fn.coefficients()
        loss, dpred = ops.diffusion_loss(eps, noise, w, wm, wl, wh, float(m.loss_fn.huber_delta), True, dpred_out=plan.dout)
        return loss, None, plan       # dL/d(eps) already sits in plan.dout

    def _ddpm_back(self, plan):
        """The backward (dL/d(eps) already sits in plan.dout), the gradient exchange and the fused Adam + EMA update.
        One rank: the whole backward as one launch list, then one update of the whole arena - inside the step graph (the bias
        corrections come from the device-resident step count).  Several ranks: the backward runs in three parts and the
        all-reduce of each third of the gradient arena starts as soon as its part has finished.
        (Measured in round 2 on one GPU: updating each third of the arena on a second stream while the rest of the backward
        runs LOSES - 2.93 ms against 2.865 ms per step, 2.89 ms with the Adam grid capped at 2 CTAs per SM: any co-running
        kernel takes SM slots from the backward's two lanes - so the update stays behind the backward.)"""
        eng = self.model.model.engine
        if self.reducer.world == 1:
            eng.run_backward(plan, None, prezeroed=True)
            self.opt.step_range(0, None, 1.0)
            return
        # The all-reduces of the first two ranges are long done when the backward ends; the last range (the gradients the backward
        # completes last, ~9.5 MB) is still on the wire for ~50 us.  The update therefore runs range by range in arrival order: Adam + EMA
        # of the first two ranges (85 % of the arena, HBM-bound) overlaps the last all-reduce (NVLink-bound) instead of waiting for it.
        parts = []
        eng.run_backward(plan, None, between=lambda lo, hi: parts.append((lo, hi, self.reducer.launch(lo, hi))), prezeroed=True)
        for lo, hi, works in parts:
            self.opt.step_range(lo, hi, self.reducer.finish(works))

    # ------------------------------------------------------------------ checkpoint interop
    def checkpoint(self, epoch: int, config=None, best_val_loss: float = float("inf")) -> dict:
        """The dict ``DDPMTrainer.save_checkpoint`` writes (trainers/ddpm_trainer.py:869-877): model and EMA-model
        ``state_dict``s under the reference's names and torch.optim.Adam's state layout, so ``torch.save`` of it resumes in
        either implementation."""
        m = self.model
        ema_sd = None
        if self.opt.ema_decay is not None and self._net.engine.flat is not None:
            ema_sd = {k: v.detach().clone() for k, v in m.state_dict().items()}          # buffers as they are
            ema_sd.update(self.opt.ema_state_dict(prefix="model."))
        return {"epoch": epoch, "model_state_dict": {k: v.detach().clone() for k, v in m.state_dict().items()},
                "ema_model_state_dict": ema_sd, "optimizer_state_dict": self.opt.state_dict(), "config": config,
                "best_val_loss": best_val_loss, "scheduler_state_dict": None}

    def load_checkpoint(self, ckpt: dict) -> int:
        """``DDPMTrainer.load_checkpoint`` (trainers/ddpm_trainer.py:897-925) for a dict read with ``torch.load``:
        parameters, Adam moments and step count, EMA weights.  Returns the epoch."""
        m = self.model
        m.load_state_dict(ckpt["model_state_dict"])
        dev = next(m.parameters()).device
        self._net.engine.prepare(dev)
        self.opt.load_state_dict(ckpt["optimizer_state_dict"], device=dev)
        if ckpt.get("ema_model_state_dict") is not None:
            self.opt.load_ema_state_dict(ckpt["ema_model_state_dict"], prefix="model.")
        self.sync_parameters(dev)      # ranks that loaded different files (or only rank 0 did) continue from rank 0's state
        return int(ckpt["epoch"])
