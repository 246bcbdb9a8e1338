#!/usr/bin/env python
"""Benchmark of the B200-native denoiser hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|ddim]

Workload (default): BASELINE.json configs[1] — DDPM CIFAR-10-shape 32x32 UNet training, bf16, batch
128 per GPU, synthetic data, data-parallel.  A step = q_sample + UNet fwd + loss + UNet bwd + gradient
all-reduce (N>1) + fused Adam/EMA.  `value` = images/s of the whole job with the batch resident in HBM;
`e2e` = the same through TrainStep.step() from pinned HOST batches with the loss read back every step.

`--impl reference` times the CPU restatement of the reference path (oracle/, kind "port": the Python
reference cannot travel to the GPU box) on the host cores with the same metric/unit.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

# algorithmic work, SURVEY.md §8(d): 2*MACs per image
FLOP_FWD = {32: 0.7756e9, 64: 3.1006e9}
FLOP_FWD_BWD = {32: 2.3232e9, 64: 9.2875e9}

LOSS_CFG = {"mse_weight": 1.0, "use_time_weighting": True, "time_weight_type": "snr",
            "time_weight_params": {"min_weight": 0.1, "max_weight": 1.0}}


def model_config(image_size, precision):
    # configs/ddpm_config.yaml model_config as the reference actually reads it
    return {"beta_start": 1e-4, "beta_end": 0.02, "image_size": image_size, "image_channels": 3, "in_channels": 3,
            "model_channels": 64, "num_timesteps": 1000, "loss_type": "mse", "loss_config": LOSS_CFG, "precision": precision,
            "ddim_sampling_steps": 50, "eta": 0.0}


def reseed_zero_init(module, seed):
    """SURVEY.md §4 pitfall: conv2/time_mlp are zero-initialised, which makes most of the network dead; give
    every all-zero weight tensor N(0, 0.02) values so the benchmark exercises real arithmetic."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() > 1 and float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tf_sustained": d["bf16_tflops_sustained"], "tf_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"tf_sustained": 1400.0, "tf_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_train_sample(budget_s=15.0, batch=16, max_steps=50):
    """BASELINE config 1 on the host cores: zero_grad -> loss_function(x).backward() -> Adam(2e-4).step(), fp32,
    batch 16, all intra-op threads.  Runs until ~budget_s of CPU work; returns (img/s, cores, description)."""
    from oracle import weights as W, unet as U, process as P, losses as L
    torch.manual_seed(0)
    sd = W.make_state_dict(W.unet_param_spec(64, 3, "model."), 1)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=2e-4)
    _, _, acp = P.linear_schedule(1e-4, 0.02, 1000)
    x = torch.randn(batch, 3, 32, 32)

    def step():
        opt.zero_grad()
        t = torch.randint(0, 1000, (batch,))
        noise = torch.randn_like(x)
        eps = U.unet_forward(params, P.q_sample(x, t, noise, acp), t)
        loss = L.diffusion_loss(eps, noise, t, "mse", LOSS_CFG)
        loss.backward()
        opt.step()
        return float(loss.detach())

    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_steps and (time.perf_counter() - t_start < budget_s or len(times) < 3):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return batch / med, torch.get_num_threads(), f"{len(times)} train steps of batch {batch} (fp32, 32x32, C=64), median step {med * 1e3:.1f} ms"


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    per_step = max(1, args.steps)
    # each "step" of this arm is one bounded CPU train step at batch 16; keep the whole run within a few minutes
    budget = min(120.0, 4.0 * (args.steps + args.warmup))
    v, cores, sample = cpu_train_sample(budget_s=budget, batch=32, max_steps=max(3, args.steps))
    line = {
        "impl": "reference", "metric": "ddpm_unet_train_images_per_s", "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 32.0 / v * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ddpm_train_32x32_C64 (BASELINE configs[1])", "global_batch": 32},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def conv_flops(p):
    """2*MACs of one dmu_conv2d / dmu_conv2d_wgrad launch from its parameter struct."""
    if hasattr(p, "Ck"):
        taps = p.R * p.S
        if p.gather == 1 and p.stride > 1:      # transposed gather touches 1/stride^2 of the taps per output pixel
            taps = taps / (p.stride * p.stride)
        return 2.0 * p.N * p.Ho * p.Wo * p.Cj * p.Ck * taps
    return 2.0 * p.N * p.Hp * p.Wp * p.Ca * p.Cb * p.R * p.S


def profile_plan(eng, plan):
    """Replay the recorded launches of one training step eagerly with a CUDA event after every launch and return
    {entry point: (launches, ms, flops)}.  The stream is first held busy (torch.cuda._sleep) so that the host runs ahead and
    the kernels execute back to back: the event deltas are then device durations, not host launch gaps."""
    from diffusion_model_universal_b200 import ops
    stream = ops._stream()
    out = {}
    big = None      # the single largest conv launch (most FLOPs)
    for lst in (plan.fwd, plan.bwd):
        lst = [(op[0], op[1]) for op in lst if op[0] is not None]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(lst) + 1)]
        torch.cuda._sleep(int(60e6))          # ~30 ms of spinning: enough for the host to enqueue everything below
        evs[0].record()
        for i, (fn, a) in enumerate(lst):
            fn(*a, stream)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, (fn, a) in enumerate(lst):
            ms = evs[i].elapsed_time(evs[i + 1])
            fl = 0.0
            if fn.__name__ in ("dmu_conv2d", "dmu_conv2d_wgrad"):
                fl = conv_flops(a[0]._obj)
                if fn.__name__ == "dmu_conv2d" and (big is None or fl > big[0] or (fl == big[0] and ms < big[1])):
                    big = (fl, ms)
            n, m, f = out.get(fn.__name__, (0, 0.0, 0.0))
            out[fn.__name__] = (n + 1, m + ms, f + fl)
    return out, big


def ddim_sample_rate(dev, batch=256, size=64):
    """BASELINE configs[2]: DDIM 50-step deterministic sampling at 64x64, batch 256 per GPU (no communication)."""
    import diffusion_model_universal_b200 as D
    m = D.DDIM(model_config(size, "bf16"))
    reseed_zero_init(m, 7)
    m.to(dev)
    with torch.no_grad():
        m.generate_samples(batch, dev)          # builds the plan, captures the graph
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x = m.generate_samples(batch, dev)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = bool(torch.isfinite(x).all())
    del m
    torch.cuda.empty_cache()
    return {"img_per_s": batch / (ms * 1e-3), "ms_per_batch": ms, "batch": batch, "image": [3, size, size], "steps": 50, "eta": 0.0,
            "unet_evals_per_s": 50 * batch / (ms * 1e-3), "model_tflops": 50 * batch * FLOP_FWD[size] / (ms * 1e-3) / 1e12, "finite": ok}


def run_ours(args):
    import torch.distributed as dist
    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import diffusion_model_universal_b200 as D
    from diffusion_model_universal_b200 import ops, _abi
    from diffusion_model_universal_b200.trainer import TrainStep
    _abi.lib()   # fail loudly if the CUDA extension is missing

    B, R = args.batch, 32
    cfg = model_config(R, "bf16")
    torch.manual_seed(1234 + rank)
    model = D.DDPM(cfg)
    reseed_zero_init(model, 7)
    model.to(dev)
    ts = TrainStep(model, lr=2e-4, ema_decay=0.9999)

    n_batches = 4
    g = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(B, 3, R, R, generator=g).pin_memory() for _ in range(n_batches)]
    devb = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.LAUNCHES
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.LAUNCHES - l0
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches

    losses = []
    # one-time setup, outside the warm-up / timed steps: launch-plan construction, tensor-map encodes, CUDA-graph capture
    # (TrainStep captures its step graph on its fourth call)
    for i in range(6):
        ts.step(devb[i % n_batches])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(lambda i: ts.step(devb[i % n_batches]), args.steps, args.warmup)
    sampler.stop_flag = True

    def e2e_step(i):
        losses.append(float(ts.step(host[i % n_batches]).item()))   # D2H read of the loss every step
    ms_e2e, _ = timed(e2e_step, args.steps, max(3, args.warmup // 2))

    # ---- roofline of the dominant kernel family (implicit-GEMM conv), live CUDA events over one replayed step
    roof = None
    cpu = None
    extra = {}
    if rank == 0:
        eng = model.model.engine
        plan = eng.get_plan(devb[0].shape, True)
        prof, big = profile_plan(eng, plan)
        pk = peaks()
        conv_ms = sum(prof[k][1] for k in ("dmu_conv2d", "dmu_conv2d_wgrad") if k in prof)
        conv_fl = sum(prof[k][2] for k in ("dmu_conv2d", "dmu_conv2d_wgrad") if k in prof)
        conv_n = sum(prof[k][0] for k in ("dmu_conv2d", "dmu_conv2d_wgrad") if k in prof)
        total_ms = sum(v[1] for v in prof.values())
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv family (conv_tc_kernel / conv3x3_halo_kernel / wgrad_tc_kernel behind dmu_conv2d + dmu_conv2d_wgrad; all launches of one step, incl. the latency-bound <= 8x8 layers)",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                "peak_source": pk["src"] + " (sustained cuBLAS bf16; kernel timed inside a long step)",
                "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_note": traffic["note"] if traffic else None,
                "launches_per_step": conv_n, "flops_per_step": conv_fl,
                "avg_launch_us": conv_ms * 1e3 / max(conv_n, 1), "share_of_step": conv_ms / total_ms if total_ms else None,
                "largest_launch": {"what": "64->64 3x3 at 32x32 x batch (fprop/dgrad)", "flops": big[0], "us": big[1] * 1e3,
                                   "tflops": big[0] / (big[1] * 1e-3) / 1e12, "frac": big[0] / (big[1] * 1e-3) / 1e12 / pk["tf_sustained"]} if big else None,
                "by_entry_point_ms": {k: round(v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_train_sample(budget_s=args.cpu_budget)
            cpu = {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample, "host_cpus": os.cpu_count()}
        if world == 1 and not args.no_ddim:
            extra["ddim50_64x64"] = ddim_sample_rate(dev)

    if rank == 0:
        gb = B * world
        value = gb * args.steps / (ms_dev * 1e-3)
        e2e = gb * args.steps / (ms_e2e * 1e-3)
        act_mb = sum(p.nbytes for lst in model.model.engine.plans.values() for p in lst) / 2 ** 20
        line = {
            "metric": "ddpm_unet_train_images_per_s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "ddpm_train_32x32_C64 (BASELINE configs[1])", "global_batch": gb, "per_gpu_batch": B, "image": [3, R, R],
                       "parallelism": f"dp{world}", "optimizer": "fused Adam(2e-4)+EMA(0.9999)", "loss": "mse x snr time-weights",
                       "l2": f"no explicit flush: the step streams a {act_mb:.0f} MiB activation/gradient arena + 190 MiB of weights/optimizer state, larger than the 126 MB L2; 4 input batches rotate"},
            "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": B * 3 * R * R * 4, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "TrainStep.step(pinned host batch) -> loss.item()"},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "flops_per_image": FLOP_FWD_BWD[R], "model_tflops": value * FLOP_FWD_BWD[R] / 1e12,
            "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(), "extra": extra,
            "last_loss": losses[-1] if losses else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch (BASELINE configs[1]: 128)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ddim", action="store_true", help="skip the DDIM-50 64x64 sampling rate reported under extra")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
