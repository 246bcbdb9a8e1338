#!/usr/bin/env python
"""Benchmark of the B200-native denoiser hot path (BASELINE.json metric: UNet train img/s and DDIM-50 / DDPM-1000 sample img/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload train|ddim|ddpm_sample|score|energy]

Workloads (one JSON line each, same schema; `--workload train` is the default the driver runs):
  train        BASELINE configs[1]: DDPM CIFAR-10-shape 32x32 UNet training, bf16, batch 128 per GPU, data-parallel.  A step =
               RNG draws + q_sample + UNet fwd + loss + UNet bwd + gradient all-reduce (N > 1) + fused Adam/EMA.
  ddim         BASELINE configs[2]: DDIM 50-step deterministic sampling at 64x64, batch 256 per GPU, sample batch sharded over
               the ranks with no communication.  A step = one full 50-evaluation sample batch.
  ddpm_sample  DDPM 1000-step ancestral sampling at 32x32, batch 128 per GPU (the metric's third number).  A step = one batch.
  score        BASELINE configs[3]: annealed Langevin sampling at 32x32 (sigma 50 -> 0.01), reduced 10 x 10 ladder
               (SURVEY.md §8d), batch 128 per GPU; unit = UNet evaluations x images / s.
  energy       BASELINE configs[4]: EnergyNet training at 32x32 with the gradient penalty (double backward), batch 64 per GPU,
               data-parallel.
`value` = whole-job throughput with the inputs resident in HBM; `e2e` = the same through the public API from pinned HOST
buffers with the result read back every step.  `--impl reference` times the CPU restatement of the reference path (oracle/,
kind "port": the Python reference cannot travel to the GPU box) on all host cores with the same metric / unit / config.
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
FLOP_ENERGY_STEP = 24.27e9          # EnergyNet train step per image (10 Langevin steps + CD + gradient penalty), SURVEY §8 a12

LOSS_CFG = {"mse_weight": 1.0, "use_time_weighting": True, "time_weight_type": "snr",
            "time_weight_params": {"min_weight": 0.1, "max_weight": 1.0}}

WORKLOADS = {
    "train": {"metric": "ddpm_unet_train_images_per_s", "unit": "img/s", "name": "ddpm_train_32x32_C64 (BASELINE configs[1])",
              "batch": 128, "size": 32},
    "ddim": {"metric": "ddim50_sample_images_per_s", "unit": "img/s", "name": "ddim50_eta0_sample_64x64_C64 (BASELINE configs[2])",
             "batch": 256, "size": 64, "evals": 50},
    "ddpm_sample": {"metric": "ddpm1000_sample_images_per_s", "unit": "img/s", "name": "ddpm1000_sample_32x32_C64 (BASELINE metric, third number)",
                    "batch": 128, "size": 32, "evals": 1000},
    "score": {"metric": "score_langevin_unet_evals_x_images_per_s", "unit": "eval*img/s",
              "name": "score_annealed_langevin_32x32_C64, 10 scales x 10 steps (BASELINE configs[3], reduced ladder of SURVEY 8d)",
              "batch": 128, "size": 32, "evals": 100},
    "energy": {"metric": "energy_train_images_per_s", "unit": "img/s", "name": "energy_train_gp_32x32_C64 (BASELINE configs[4])",
               "batch": 64, "size": 32},
}


def model_config(image_size, precision):
    # configs/ddpm_config.yaml model_config as the reference actually reads it
    return {"beta_start": 1e-4, "beta_end": 0.02, "image_size": image_size, "image_channels": 3, "in_channels": 3,
            "model_channels": 64, "num_timesteps": 1000, "loss_type": "mse", "loss_config": LOSS_CFG, "precision": precision,
            "ddim_sampling_steps": 50, "eta": 0.0}


def score_config(precision, num_scales=10, langevin_steps=10):
    # configs/score_based_config.yaml as read by models/score_based.py:150-170 (+ in_channels)
    return {"sigma_min": 0.01, "sigma_max": 50.0, "num_scales": num_scales, "beta": 1.0, "in_channels": 3, "model_channels": 64,
            "image_size": 32, "image_channels": 3, "loss_type": "score_matching", "langevin_steps": langevin_steps, "precision": precision}


def energy_config(precision):
    # configs/energy_based_config.yaml:7-9 (langevin_steps 10, step 0.01, lambda 0.01) with the §8c repairs
    return {"num_timesteps": 1000, "beta_start": 1e-4, "beta_end": 0.02, "use_time_conditioning": False, "in_channels": 3,
            "model_channels": 64, "image_size": 32, "image_channels": 3, "loss_type": "energy_based", "energy_scale": 1.0,
            "regularization_weight": 0.01, "langevin_steps": 10, "langevin_step_size": 0.01, "precision": precision}


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


def config_of(workload, world):
    """The `config` object of the JSON line: identical for both arms (the reference arm times a bounded sample of it)."""
    w = WORKLOADS[workload]
    c = {"workload": w["name"], "global_batch": w["batch"] * world, "per_gpu_batch": w["batch"], "image": [3, w["size"], w["size"]],
         "parallelism": f"dp{world}" if workload in ("train", "energy") else f"sample batch sharded over {world} rank(s), no communication"}
    if workload == "train":
        c.update({"optimizer": "fused Adam(2e-4)+EMA(0.9999)", "loss": "mse x snr time-weights"})
    elif workload == "energy":
        c.update({"optimizer": "fused Adam(2e-4)", "loss": "contrastive divergence + 0.01 x gradient penalty, 10 Langevin steps"})
    else:
        c.update({"unet_evals_per_step": w["evals"]})
    # timing rule: flush L2 between iterations or use inputs larger than L2 - here the working set itself is larger
    c["l2"] = {"train": "no explicit flush: every step streams its activation/gradient arena (~640 MiB, extra.arena_mib) + 190 MiB of weights/optimizer "
                        "state, larger than the 126 MB L2; 4 input batches rotate",
               "energy": "no explicit flush: 4 input batches rotate; every evaluation allocates fresh activation tensors (eager launches)"
               }.get(workload, "no explicit flush: every UNet evaluation streams its activation arena (extra.arena_mib; 2.1 GiB at 256x64x64) + 32 MB of "
                               "bf16 filters, larger than the 126 MB L2")
    return c


# ----------------------------------------------------------------------------------------------- CPU reference arm
def use_all_host_cores():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU program and gets every host
    core whatever launched it (round 1's N > 1 reference numbers ran on one thread)."""
    n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def _median_time(step, budget_s, max_steps, min_steps=3):
    step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_steps and (time.perf_counter() - t_start < budget_s or len(times) < min_steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    return times[len(times) // 2], len(times)


def cpu_sample(workload, budget_s=15.0, max_steps=50, batch=None):
    """The reference path of `workload` (oracle/ restatement, fp32) on the host cores, bounded to ~budget_s of CPU work.
    Returns (value in the workload's unit, threads used, description of the sample, median ms of one sample step)."""
    from oracle import weights as W, unet as U, process as P, losses as L
    w = WORKLOADS[workload]
    cores = use_all_host_cores()
    torch.manual_seed(0)
    R = w["size"]
    if workload == "train":
        # BASELINE config 1 arithmetic: zero_grad -> loss_function(x).backward() -> Adam(2e-4).step()
        batch = batch or 16
        sd = W.make_state_dict(W.unet_param_spec(64, 3, "model."), 1)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(params.values()), lr=2e-4)
        _, _, acp = P.linear_schedule(1e-4, 0.02, 1000)
        x = torch.randn(batch, 3, R, R)

        def step():
            opt.zero_grad()
            t = torch.randint(0, 1000, (batch,))
            noise = torch.randn_like(x)
            eps = U.unet_forward(params, P.q_sample(x, t, noise, acp), t)
            loss = L.diffusion_loss(eps, noise, t, "mse", LOSS_CFG)
            loss.backward()
            opt.step()
        med, n = _median_time(step, budget_s, max_steps)
        return batch / med, cores, f"{n} train steps of batch {batch} (fp32, {R}x{R}, C=64), median step {med * 1e3:.1f} ms", med * 1e3
    if workload in ("ddim", "ddpm_sample", "score"):
        # a sampling chain is `evals` identical (UNet evaluation + update) iterations: time a few of them at a small batch
        batch = batch or (4 if R == 64 else 16)
        spec = W.scorenet_param_spec if workload == "score" else W.unet_param_spec
        sd = W.make_state_dict(spec(64, 3, "model."), 1)
        betas, alphas, acp = P.linear_schedule(1e-4, 0.02, 1000)
        x0 = torch.randn(batch, 3, R, R)
        state = {"x": x0}
        if workload == "ddim":
            ts = P.ddim_timesteps(1000, 50)
            tables = P.ddim_tables(acp, ts, 0.0)
        sig = P.score_sigma_ladder(0.01, 50.0, 10) if workload == "score" else None

        def step():
            with torch.no_grad():
                x = state["x"]
                if workload == "ddim":
                    eps = U.unet_forward(sd, x, torch.full((batch,), int(ts[25])))
                    x = P.ddim_step(x, eps, torch.full((batch,), 25), tables, 0.0)
                elif workload == "ddpm_sample":
                    t = torch.full((batch,), 500)
                    eps = U.unet_forward(sd, x, t)
                    x = P.ddpm_reverse_step(x, eps, t, torch.randn_like(x), betas, alphas, acp)
                else:
                    s = U.scorenet_forward(sd, x, sig[5].expand(batch))
                    x = P.score_langevin_step(x, s, torch.randn_like(x), sig[5], 1.0)
                state["x"] = x0          # keep the iterate bounded: every timed evaluation sees the same magnitudes
        med, n = _median_time(step, budget_s, max_steps)
        per_eval = batch / med          # eval*img/s
        value = per_eval if workload == "score" else per_eval / w["evals"]
        return value, cores, (f"{n} (UNet evaluation + update) iterations of batch {batch} (fp32, {R}x{R}, C=64), median {med * 1e3:.1f} ms; "
                              f"a chain is {w['evals']} such iterations, so {w['unit']} = batch / (median x {1 if workload == 'score' else w['evals']})"), med * 1e3
    if workload == "energy":
        batch = batch or 16
        sd = W.make_state_dict(W.energynet_param_spec(64, 3, "model."), 1)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(params.values()), lr=2e-4)
        _, _, acp = P.linear_schedule(1e-4, 0.02, 1000)
        x = torch.randn(batch, 3, R, R)
        efn = lambda z: U.energynet_forward(params, z)

        def step():
            opt.zero_grad()
            t = torch.randint(0, 1000, (batch,))
            xf = P.q_sample(x, t, torch.randn_like(x), acp)
            for _ in range(10):     # models/energy_based.py:264-278
                xf = xf.detach().requires_grad_(True)
                g = torch.autograd.grad(efn(xf).sum(), xf)[0]
                xf = P.energy_langevin_step(xf.detach(), g, torch.randn_like(x), 0.01)
            loss = L.energy_loss(efn, x, xf.detach(), torch.rand(batch, 1, 1, 1), 0.01)
            loss.backward()
            opt.step()
        med, n = _median_time(step, budget_s, max_steps)
        return batch / med, cores, f"{n} EnergyNet train steps of batch {batch} (fp32, {R}x{R}, C=64, 10 Langevin steps + CD + GP), median step {med * 1e3:.1f} ms", med * 1e3
    raise ValueError(workload)


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    # each "step" of this arm is one bounded CPU sample step; keep the whole run within a few minutes
    budget = min(120.0, 4.0 * (args.steps + args.warmup))
    v, cores, sample, med_ms = cpu_sample(args.workload, budget_s=budget, max_steps=max(3, args.steps), batch=32 if args.workload == "train" else None)
    line = {
        "impl": "reference", "metric": w["metric"], "value": v, "unit": w["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(args.workload, max(world, 1)),
        "cpu_baseline": {"value": v, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample, "host_cpus": os.cpu_count(),
                         "note": "one CPU program on all host cores at every N (it does not scale with the GPU count)"},
        "e2e": {"value": v, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm: helpers
def conv_flops(p):
    """2*MACs of one dmu_conv2d / dmu_conv2d_wgrad launch from its parameter struct."""
    if hasattr(p, "Ck"):
        taps = p.R * p.S
        if p.gather == 1 and p.stride > 1:      # transposed gather touches 1/stride^2 of the taps per output pixel
            taps = taps / (p.stride * p.stride)
        return 2.0 * p.N * p.Ho * p.Wo * p.Cj * p.Ck * taps
    return 2.0 * p.N * p.Hp * p.Wp * p.Ca * p.Cb * p.R * p.S


def profile_plan(plan, lists=("fwd", "bwd")):
    """Replay the recorded launches of a plan eagerly with a CUDA event after every launch.  The stream is first held busy
    (torch.cuda._sleep) so that the host runs ahead and the kernels execute back to back: the event deltas are then device
    durations, not host launch gaps.  Returns ({entry point: (launches, ms, flops)}, largest conv launch, {lane: ms})."""
    from diffusion_model_universal_b200 import ops
    stream = ops._stream()
    out, lanes = {}, {}
    big = None      # the single largest conv launch (most FLOPs)
    for which in lists:
        lst = [(op[0], op[1], (op[2] if len(op) > 2 else 0)) for op in getattr(plan, which) if op[0] is not None]
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(lst) + 1)]
        torch.cuda._sleep(int(60e6))          # ~30 ms of spinning: enough for the host to enqueue everything below
        evs[0].record()
        for i, (fn, a, _) in enumerate(lst):
            fn(*a, stream)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i, (fn, a, lane) in enumerate(lst):
            ms = evs[i].elapsed_time(evs[i + 1])
            fl = 0.0
            if fn.__name__ in ("dmu_conv2d", "dmu_conv2d_wgrad"):
                fl = conv_flops(a[0]._obj)
                if fn.__name__ == "dmu_conv2d" and (big is None or fl > big[0] or (fl == big[0] and ms < big[1])):
                    big = (fl, ms)
            n, m, f = out.get(fn.__name__, (0, 0.0, 0.0))
            out[fn.__name__] = (n + 1, m + ms, f + fl)
            key = f"{which}_{'side' if lane % 2 else 'main'}_lane_serial_ms"
            lanes[key] = lanes.get(key, 0.0) + ms
    return out, big, lanes


def conv_roofline(plan, lists, graph_ms, what, step_ms=None, execs_per_step=1):
    """Tensor roofline of the implicit-GEMM conv family.  `achieved` is measured LIVE over the timed region: the algorithmic FLOPs
    of every conv launch of one step / the step time (CUDA events around the K timed steps, `step_ms`).  The convs run inside a
    CUDA graph on two overlapping lanes and share the step with the GroupNorm / attention / optimizer launches, and since round 2
    most GroupNorm(+SiLU) work runs INSIDE the conv launches' epilogues - a per-launch time for "the conv part" no longer
    exists, so the whole step is charged: a lower bound on the family's in-situ rate.  The round-1 figure (each launch replayed
    eagerly, one at a time, with an event after it: no lane overlap, cold-ish caches, fused epilogue work counted as conv time)
    stays as `eager_serialized`."""
    prof, big, lanes = profile_plan(plan, lists)
    pk = peaks()
    keys = [k for k in ("dmu_conv2d", "dmu_conv2d_wgrad") if k in prof]
    conv_ms = sum(prof[k][1] for k in keys)
    conv_fl = sum(prof[k][2] for k in keys)
    conv_n = sum(prof[k][0] for k in keys)
    eager = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    achieved = conv_fl * execs_per_step / (step_ms * 1e-3) / 1e12 if step_ms else eager
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
            break
    return {"bound": "tensor",
            "kernel": "tcgen05 implicit-GEMM conv family (conv_tc_kernel / conv3x3_halo_kernel / wgrad_tc_kernel behind dmu_conv2d + "
                      "dmu_conv2d_wgrad; all conv launches of " + what + ", incl. the latency-bound <= 8x8 layers)",
            "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
            "how": ("conv FLOPs of one step (%d plan executions x %.2f GFLOP) / step time of the timed region (%.4f ms, CUDA events); "
                    "the step graph also holds the non-conv launches" % (execs_per_step, conv_fl / 1e9, step_ms)) if step_ms else "eager per-launch events",
            "eager_serialized": {"achieved": eager, "frac": eager / pk["tf_sustained"], "avg_launch_us": conv_ms * 1e3 / max(conv_n, 1),
                                 "what": "sum of conv FLOPs / sum of per-launch times, every launch replayed eagerly with a CUDA event after it "
                                         "(fused GroupNorm epilogue time included, no lane overlap)"},
            "peak_source": pk["src"] + " (sustained cuBLAS bf16; kernel timed inside a long step)",
            "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_note": traffic["note"] if traffic else None,
            "launches": conv_n, "flops": conv_fl, "avg_launch_us": conv_ms * 1e3 / max(conv_n, 1),
            # critical path: the graph runs the lanes side by side, so shares are given against the replayed graph time, next
            # to the serial (one launch at a time) sum of each lane
            "graph_ms": graph_ms, "lanes_serial_ms": {k: round(v, 4) for k, v in lanes.items()},
            "largest_launch": {"what": "largest conv launch by FLOPs (64->64 3x3 at full resolution x batch)", "flops": big[0], "us": big[1] * 1e3,
                               "tflops": big[0] / (big[1] * 1e-3) / 1e12, "frac": big[0] / (big[1] * 1e-3) / 1e12 / pk["tf_sustained"]} if big else None,
            "by_entry_point_ms": {k: round(v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}


def timeit_cuda(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def gpu_eager_reference(workload, dev):
    """Informational: the reference's own eager path (oracle/ restatement = the same ATen calls -> cuDNN / cuBLAS)
    on this GPU, fp32 with TF32 off and bf16 autocast, outside the timed region.  Not the reference arm; not a target."""
    from oracle import weights as W, unet as U, process as P, losses as L
    w = WORKLOADS[workload]
    B, R = w["batch"], w["size"]
    out = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        if workload == "energy":
            return None
        spec = W.scorenet_param_spec if workload == "score" else W.unet_param_spec
        sd = {k: v.to(dev) for k, v in W.make_state_dict(spec(64, 3, "model."), 1).items()}
        _, _, acp = P.linear_schedule(1e-4, 0.02, 1000, device=dev)
        x = torch.randn(B, 3, R, R, device=dev)
        for mode in ("fp32", "bf16_autocast"):
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else torch.autocast("cuda", enabled=False)
            if workload == "train":
                params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
                opt = torch.optim.Adam(list(params.values()), lr=2e-4)

                def step():
                    opt.zero_grad()
                    t = torch.randint(0, 1000, (B,), device=dev)
                    noise = torch.randn_like(x)
                    with ctx:
                        eps = U.unet_forward(params, P.q_sample(x, t, noise, acp), t)
                    loss = L.diffusion_loss(eps.float(), noise, t, "mse", LOSS_CFG)
                    loss.backward()
                    opt.step()
                ms = timeit_cuda(step, n=5, warm=2)
                out[mode] = {"img_per_s": B / (ms * 1e-3), "ms_per_step": ms}
            else:
                t = torch.full((B,), 500, device=dev)
                sg = torch.full((B,), 1.0, device=dev)

                def step():
                    with torch.no_grad(), ctx:
                        U.scorenet_forward(sd, x, sg) if workload == "score" else U.unet_forward(sd, x, t)
                ms = timeit_cuda(step, n=5, warm=2)
                evals = 1 if workload == "score" else w["evals"]
                out[mode] = {w["unit"]: B / (ms * 1e-3) / evals, "ms_per_unet_eval": ms}
        out["what"] = ("oracle/ (the reference's ATen call sequence) run eagerly on this GPU: cuDNN / cuBLAS library kernels, "
                       f"batch {B}; sampling entries time the UNet evaluation only")
    except Exception as e:   # informational only
        out["error"] = repr(e)[:200]
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- our arm: workloads
class Train:
    def __init__(self, args, dev, rank, world, precision="bf16"):
        import diffusion_model_universal_b200 as D
        from diffusion_model_universal_b200.trainer import TrainStep
        self.B, self.R = args.batch or WORKLOADS["train"]["batch"], 32
        torch.manual_seed(1234)                 # identical initial weights on every rank (TrainStep also broadcasts rank 0's)
        self.model = D.DDPM(model_config(self.R, precision))
        reseed_zero_init(self.model, 7)
        self.model.to(dev)
        self.ts = TrainStep(self.model, lr=2e-4, ema_decay=0.9999)
        g = torch.Generator().manual_seed(1234 + rank)       # per-rank data
        self.host = [torch.randn(self.B, 3, self.R, self.R, generator=g).pin_memory() for _ in range(4)]
        self.devb = [h.to(dev) for h in self.host]
        torch.manual_seed(4321 + rank)          # per-rank noise / timestep stream
        self.losses = []
        # one-time setup, outside the warm-up / timed steps: launch-plan construction, tensor-map encodes, CUDA-graph capture
        for i in range(6):
            self.ts.step(self.devb[i % 4])
        self.units = self.B
        self.h2d, self.d2h = self.B * 3 * self.R * self.R * 4, 4
        self.api = "TrainStep.step(pinned host batch) -> loss.item()"
        self.flops_per_unit = FLOP_FWD_BWD[self.R]

    def step_dev(self, i):
        self.ts.step(self.devb[i % 4])

    def step_e2e(self, i):
        self.losses.append(float(self.ts.step(self.host[i % 4]).item()))   # D2H read of the loss every step

    def roofline(self, step_ms=None):
        eng = self.model.model.engine
        plan = eng.get_plan(self.devb[0].shape, True)
        g = {}
        if "fwd" in plan.graphs and "bwd" in plan.graphs:
            g = {"fwd": timeit_cuda(lambda: plan.graphs["fwd"].replay()), "bwd": timeit_cuda(lambda: plan.graphs["bwd"].replay())}
        return conv_roofline(plan, ("fwd", "bwd"), g, "one training step", step_ms, 1)

    def extra(self):
        ex = {"last_loss": self.losses[-1] if self.losses else None}
        if hasattr(self, "model"):
            ex["arena_mib"] = round(sum(p.nbytes for lst in self.model.model.engine.plans.values() for p in lst) / 2 ** 20, 1)
        return ex


class Sampler:
    """ddim / ddpm_sample / score: a step is one full sample batch through the public generate_samples()."""

    def __init__(self, workload, args, dev, rank, world):
        import diffusion_model_universal_b200 as D
        w = WORKLOADS[workload]
        self.w, self.workload, self.dev = w, workload, dev
        self.B, self.R = args.batch or w["batch"], w["size"]
        torch.manual_seed(1234)
        if workload == "score":
            self.model = D.ScoreBasedDiffusion(score_config("bf16"))
        else:
            self.model = (D.DDIM if workload == "ddim" else D.DDPM)(model_config(self.R, "bf16"))
        reseed_zero_init(self.model, 7)
        self.model.to(dev)
        torch.manual_seed(1234 + rank)          # config 3: "initial noise torch.randn seed 1234 + rank"
        with torch.no_grad():
            self.last = self.model.generate_samples(self.B, dev)        # builds the plan, captures the graph
        self.hostbuf = torch.empty(self.B, 3, self.R, self.R).pin_memory()
        self.units = self.B * (w["evals"] if workload == "score" else 1)
        self.h2d, self.d2h = 0, self.B * 3 * self.R * self.R * 4
        self.api = ("generate_samples(batch, device) -> samples copied to pinned host memory (the API takes no input tensor: the initial "
                    "noise is drawn on the device, as the reference does with torch.randn(device=...))")
        self.flops_per_unit = FLOP_FWD[self.R] * (1 if workload == "score" else w["evals"])

    def step_dev(self, i):
        with torch.no_grad():
            self.last = self.model.generate_samples(self.B, self.dev)

    def step_e2e(self, i):
        with torch.no_grad():
            self.last = self.model.generate_samples(self.B, self.dev)
        self.hostbuf.copy_(self.last, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def roofline(self, step_ms=None):
        eng = self.model.model.engine
        plan = eng.get_plan((self.B, 3, self.R, self.R), False)
        g = {"fwd": timeit_cuda(lambda: plan.graphs["fwd"].replay())} if "fwd" in plan.graphs else {}
        return conv_roofline(plan, ("fwd",), g, "one UNet evaluation", step_ms, self.w["evals"])

    def extra(self):
        return {"finite": bool(torch.isfinite(self.last).all()), "unet_evals_per_step": self.w["evals"],
                "arena_mib": round(sum(p.nbytes for lst in self.model.model.engine.plans.values() for p in lst) / 2 ** 20, 1)}


class EnergyTrain:
    def __init__(self, args, dev, rank, world):
        import diffusion_model_universal_b200 as D
        from diffusion_model_universal_b200.trainer import TrainStep
        self.B, self.R = args.batch or WORKLOADS["energy"]["batch"], 32
        torch.manual_seed(1234)
        self.model = D.EnergyBasedDiffusion(energy_config("bf16")).to(dev)
        self.ts = TrainStep(self.model, lr=2e-4, ema_decay=None)
        g = torch.Generator().manual_seed(1234 + rank)
        self.host = [torch.randn(self.B, 3, self.R, self.R, generator=g).pin_memory() for _ in range(4)]
        self.devb = [h.to(dev) for h in self.host]
        torch.manual_seed(4321 + rank)
        self.losses = []
        for i in range(3):
            self.ts.step(self.devb[i % 4])
        self.units = self.B
        self.h2d, self.d2h = self.B * 3 * self.R * self.R * 4, 4
        self.api = "TrainStep.step(pinned host batch) -> loss.item()"
        self.flops_per_unit = FLOP_ENERGY_STEP

    def step_dev(self, i):
        self.ts.step(self.devb[i % 4])

    def step_e2e(self, i):
        self.losses.append(float(self.ts.step(self.host[i % 4]).item()))

    def roofline(self, step_ms=None):
        return None     # filled from the step time by the caller (eager launches, no recorded plan)

    def extra(self):
        return {"last_loss": self.losses[-1] if self.losses else None}


def quick_rate(make, steps=3, warm=1):
    """(units/s, ms per step) of a workload object outside the main timed region (the `extra` entries of the default line)."""
    wl = make()
    for i in range(warm):
        wl.step_dev(i)
    ms = timeit_cuda(lambda: wl.step_dev(0), n=steps, warm=0)
    ex = wl.extra()
    r = {"value": wl.units / (ms * 1e-3), "ms_per_step": ms, "per_gpu_batch": wl.B, "model_tflops": wl.units * wl.flops_per_unit / (ms * 1e-3) / 1e12}
    r.update({k: v for k, v in ex.items() if k in ("finite", "unet_evals_per_step")})
    del wl
    torch.cuda.empty_cache()
    return r


def dropin_api_rate(dev, B=128, R=32, steps=10):
    """The literal drop-in the reference trainer runs (trainers/ddpm_trainer.py:542-555): zero_grad -> loss_function().backward()
    -> torch.optim.Adam.step() -> the per-parameter Python EMA loop, on this package's DDPM class (bf16 mode)."""
    import copy
    import diffusion_model_universal_b200 as D
    torch.manual_seed(1234)
    m = D.DDPM(model_config(R, "bf16"))
    reseed_zero_init(m, 7)
    m.to(dev)
    x = torch.randn(B, 3, R, R, device=dev)
    m.loss_function(x).backward()          # arenas exist from here on
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    ema = {k: p.detach().clone() for k, p in m.named_parameters()}

    def step():
        opt.zero_grad()
        loss = m.loss_function(x)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for k, p in m.named_parameters():
                ema[k].mul_(0.9999).add_(p.detach(), alpha=1 - 0.9999)
        return loss
    ms = timeit_cuda(step, n=steps, warm=3)
    del m, opt, ema
    torch.cuda.empty_cache()
    return {"img_per_s": B / (ms * 1e-3), "ms_per_step": ms,
            "api": "loss_function(x).backward() + torch.optim.Adam.step() + Python EMA loop (no TrainStep, no fused optimizer)"}


def run_ours(args):
    import torch.distributed as dist
    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the package's NCCL defaults (diffusion_model_universal_b200/parallel.py: NCCL_MAX_CTAS) must be in the environment before the
        # communicator exists
        import diffusion_model_universal_b200.parallel  # noqa: F401
        dist.init_process_group("nccl", device_id=dev)
    from diffusion_model_universal_b200 import ops, _abi
    _abi.lib()   # fail loudly if the CUDA extension is missing
    w = WORKLOADS[args.workload]

    if args.workload == "train":
        wl = Train(args, dev, rank, world)
    elif args.workload == "energy":
        wl = EnergyTrain(args, dev, rank, world)
    else:
        wl = Sampler(args.workload, args, dev, rank, world)

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

    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(wl.step_dev, args.steps, args.warmup)
    sampler.stop_flag = True
    ms_e2e, _ = timed(wl.step_e2e, args.steps, max(3, args.warmup // 2))

    roof = cpu = eager = None
    extra = {}
    if rank == 0:
        roof = wl.roofline(ms_dev / args.steps)
        if roof is None:       # eager workload without a recorded plan: whole-step model FLOPs against the tensor peak
            pk = peaks()
            ach = wl.units * wl.flops_per_unit / (ms_dev / args.steps * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "whole step (algorithmic FLOPs of the step / step time; the conv kernels are the UNet's)",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "peak_source": pk["src"] + " (sustained cuBLAS bf16)", "traffic": None}
        extra.update(wl.extra())
        if world == 1 and not args.no_cpu:
            v, cores, sample, _ = cpu_sample(args.workload, budget_s=args.cpu_budget)
            cpu = {"value": v, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample, "host_cpus": os.cpu_count()}
        if world == 1 and not args.no_extras:
            eager = gpu_eager_reference(args.workload, dev)
            if args.workload == "train":
                del wl.ts, wl.model
                torch.cuda.empty_cache()

                class _A:
                    batch = None
                mk = lambda name: (lambda: Sampler(name, _A, dev, rank, world))
                extra["ddim50_64x64"] = quick_rate(mk("ddim"), steps=2, warm=1)
                extra["ddpm1000_32x32"] = quick_rate(mk("ddpm_sample"), steps=1, warm=1)
                extra["score_langevin_10x10_32x32"] = quick_rate(mk("score"), steps=3, warm=1)
                extra["energy_train_32x32"] = quick_rate(lambda: EnergyTrain(_A, dev, rank, world), steps=5, warm=2)
                extra["fp32_mode_train"] = quick_rate(lambda: Train(_A, dev, rank, world, precision="fp32"), steps=5, warm=2)
                extra["dropin_api_train"] = dropin_api_rate(dev)
                extra["units"] = {"ddim50_64x64": "img/s", "ddpm1000_32x32": "img/s", "score_langevin_10x10_32x32": "eval*img/s",
                                  "energy_train_32x32": "img/s", "fp32_mode_train": "img/s (precision: fp32, SIMT conv path)"}

    if rank == 0:
        units = wl.units * world
        value = units * args.steps / (ms_dev * 1e-3)
        e2e = units * args.steps / (ms_e2e * 1e-3)
        cfg = config_of(args.workload, world)
        line = {
            "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e, "unit": w["unit"], "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h, "ms_per_step": ms_e2e / args.steps,
                    "api": wl.api},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "flops_per_unit": wl.flops_per_unit, "model_tflops": value * wl.flops_per_unit / 1e12,
            "roofline": roof, "cpu_baseline": cpu, "gpu_eager_reference": eager, "clocks": sampler.summary(), "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", "--no-ddim", dest="no_extras", action="store_true",
                    help="skip the informational entries (other workloads' rates under extra, eager-GPU reference)")
    args = ap.parse_args()
    # defaults sized so that every workload finishes within minutes (a DDPM-1000 step is ~1 s)
    if args.steps is None:
        args.steps = {"train": 20, "energy": 20, "ddim": 5, "ddpm_sample": 3, "score": 10}[args.workload]
    if args.warmup is None:
        args.warmup = {"train": 5}.get(args.workload, 3)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
