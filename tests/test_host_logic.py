"""CPU tier: host-side logic of the package (plan construction, pointer/pitch arithmetic, gradient routing,
state-dict contract, default initialisation) run against the host-memory test double of the C ABI
(tests/fake_device.py).  Kernel numerics are NOT claimed here — see the -m gpu tests."""

import json
import hashlib
import os

import pytest
import torch

import fake_device
from conftest import GOLDEN, load_golden, rel_l2
from oracle import weights as W

SHIPPED_LOSS = {"use_time_weighting": True, "time_weight_type": "snr", "time_weight_params": {"min_weight": 0.1, "max_weight": 1.0}}


def _cfg(C, **kw):
    c = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": C,
         "loss_type": "mse", "loss_config": dict(SHIPPED_LOSS)}
    c.update(kw)
    return c


def _load(model, spec, seed):
    sd = model.state_dict()
    sd.update(W.make_state_dict(spec, seed))
    model.load_state_dict(sd)
    return model


def test_contract_and_default_init_match_reference():
    """Same keys/shapes/order as the reference's state_dict and, for the same torch seed, bit-identical default
    initialisation (the parameter containers are created in the reference's order)."""
    import diffusion_model_universal_b200 as D
    c = json.load(open(os.path.join(GOLDEN, "contract.json")))
    torch.manual_seed(c["default_init_seed"])
    m = D.DDPM({"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "loss_type": "mse"})
    sd = m.state_dict()
    assert [(k, list(v.shape)) for k, v in sd.items()] == [(k, s) for k, s in c["ddpm_keys"]]
    for k, v in sd.items():
        assert hashlib.sha256(v.contiguous().numpy().tobytes()).hexdigest() == c["default_init_sha"][k], k
    d = D.DDIM({"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "ddim_sampling_steps": 50, "eta": 0.0})
    assert [(k, list(v.shape)) for k, v in d.state_dict().items()] == [(k, s) for k, s in c["ddim_keys"]]
    assert d.ddim_timesteps.tolist() == c["ddim_timesteps_uniform"]
    t = load_golden("tables.pt")
    assert torch.equal(m.betas, t["betas"]) and torch.equal(m.alphas_cumprod, t["alphas_cumprod"])
    for method in ("uniform", "quad"):
        for eta in (0.0, 0.5):
            dd = D.DDIM({"beta_start": 1e-4, "beta_end": 0.02, "ddim_sampling_steps": 50, "eta": eta, "ddim_discretize_method": method})
            g = t[f"ddim_{method}_{eta}"]
            assert torch.equal(dd.ddim_timesteps, g["timesteps"]) and torch.equal(dd.ddim_alphas, g["alphas"])
            assert torch.equal(dd.ddim_alphas_prev, g["alphas_prev"])
            assert torch.equal(torch.nan_to_num(dd.ddim_sigmas, nan=-7.0), torch.nan_to_num(g["sigmas"], nan=-7.0))


def test_cpu_input_fails_loudly():
    import diffusion_model_universal_b200 as D
    m = D.DDPM(_cfg(32))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 32, 32), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.loss_function(torch.randn(1, 3, 32, 32))


@pytest.mark.parametrize("tag", ["c32_r32", "c32_r64"])
def test_forward_plan(monkeypatch, tag):
    import diffusion_model_universal_b200 as D
    fake_device.install(monkeypatch)
    f = load_golden("unet_forward.pt")[tag]
    net = D.UNet(3, f["C"], 3)
    net.load_state_dict(W.make_state_dict(W.unet_param_spec(f["C"], 3, ""), f["seed"]))
    with torch.no_grad():
        y = net(f["x"], f["t"])
        y2 = net(f["x"], f["t"])     # cached plan, repacked weights
    assert rel_l2(y, f["eps"]) < 1e-4
    assert torch.equal(y, y2)
    with pytest.raises(ValueError):
        net(torch.randn(1, 3, 24, 24), torch.zeros(1, dtype=torch.long))   # five stride-2 stages need multiples of 32
    with pytest.raises(ValueError):
        net(torch.randn(1, 4, 32, 32), torch.zeros(1, dtype=torch.long))


def test_backward_plan(monkeypatch):
    import diffusion_model_universal_b200 as D
    fake_device.install(monkeypatch)
    f = load_golden("ddpm_train.pt")
    m = _load(D.DDPM(_cfg(f["C"])), W.unet_param_spec(f["C"], 3, "model."), f["wseed"])
    torch.manual_seed(f["rng_seed"])
    loss = m.loss_function(f["x0"])       # same RNG call order as ddpm.py:223-226
    assert abs(loss.item() - f["loss"].item()) < 1e-4 * abs(f["loss"].item())
    loss.backward()
    scale = max(f["grad_norm"].values())
    for k, p in m.named_parameters():
        gn = f["grad_norm"][k]
        assert p.grad is not None and p.grad.shape == p.shape, k
        if gn < 1e-6 * scale:
            assert p.grad.norm().item() < 1e-5 * scale, k
        else:
            assert abs(p.grad.norm().item() - gn) < 1e-3 * gn, k
    for k, g in f["grad_full"].items():
        if g.norm() > 1e-6 * scale:
            assert rel_l2(dict(m.named_parameters())[k].grad, g) < 1e-3, k
    # accumulation without zero_grad must double, zero_grad(set_to_none=False) must not alias the arena
    g1 = m.model.initial_conv.weight.grad.clone()
    torch.manual_seed(f["rng_seed"])
    m.loss_function(f["x0"]).backward()
    assert rel_l2(m.model.initial_conv.weight.grad, 2 * g1) < 1e-5
    m.zero_grad(set_to_none=False)
    torch.manual_seed(f["rng_seed"])
    m.loss_function(f["x0"]).backward()
    assert rel_l2(m.model.initial_conv.weight.grad, g1) < 1e-5


def test_score_model_plan(monkeypatch):
    import diffusion_model_universal_b200 as D
    fake_device.install(monkeypatch)
    f = load_golden("score.pt")
    m = _load(D.ScoreBasedDiffusion(dict(f["cfg"])), W.scorenet_param_spec(f["C"], 3, "model."), f["wseed"])
    with torch.no_grad():
        s = m.forward(f["x0"], f["sigma"])
    assert rel_l2(s, f["score"]) < 1e-4
    torch.manual_seed(f["loss_seed"])
    loss = m.loss_function(f["x0"])      # rand, randn_like, fresh randn_like: same order as the reference
    assert abs(loss.item() - f["loss"].item()) < 1e-4 * abs(f["loss"].item())
    loss.backward()
    assert m.model.time_embed[0].weight.grad is not None


def test_ddim_and_ddpm_loops(monkeypatch):
    import diffusion_model_universal_b200 as D
    fake_device.install(monkeypatch)
    f = load_golden("ddpm_chain.pt")
    # FakeLib has no ddpm/ddim step: the loops below must only need the entry points it implements + these two
    import oracle.process as P
    fake = D._abi.lib()

    def ddpm_step(x, eps, noise, t, betas, alphas, acp, T, out, batch, inner, stream):
        n = batch * inner
        fl = fake_device._flat
        tv = fl(t, batch, dtype=torch.int64)
        z = fl(noise, n).view(batch, inner, 1, 1) if noise else None
        r = P.ddpm_reverse_step(fl(x, n).view(batch, inner, 1, 1), fl(eps, n).view(batch, inner, 1, 1), tv, z,
                                fl(betas, T), fl(alphas, T), fl(acp, T))
        fl(out, n).copy_(r.reshape(-1))
        return 0
    fake.__dict__["dmu_ddpm_step"] = ddpm_step
    m = _load(D.DDPM(_cfg(f["C"], num_timesteps=10)), W.unet_param_spec(f["C"], 3, "model."), f["wseed"])
    torch.manual_seed(f["rng_seed"])
    inter = m.generate_samples_with_intermediates(2, torch.device("cpu"), save_interval=f["save_interval"])
    assert len(inter) == len(f["intermediates"])
    for u, v in zip(inter, f["intermediates"]):     # same RNG order as ddpm.py:249,324
        assert rel_l2(u, v) < 1e-3
    torch.manual_seed(f["rng_seed"])
    assert rel_l2(m.generate_samples(2, torch.device("cpu")), f["final"]) < 1e-3


def test_snr_time_weights_without_host_sync_match_the_reference_formula():
    """losses.py:150-157 sizes its beta table with timesteps.max().item(); the device-resident variant must give the same
    weights (same linspace formula, t_max kept as a tensor) for every t_max, including the degenerate ones."""
    from diffusion_model_universal_b200.losses import DiffusionLoss
    from oracle import losses as OL
    g = torch.Generator().manual_seed(3)
    cases = [torch.randint(0, 1000, (128,), generator=g), torch.randint(0, 10, (16,), generator=g), torch.zeros(4, dtype=torch.long),
             torch.tensor([1, 0, 1]), torch.tensor([999]), torch.tensor([5, 5, 5, 5]), torch.arange(0, 1000, 37)]
    for t in cases:
        a = DiffusionLoss("mse", dict(SHIPPED_LOSS))
        b = DiffusionLoss("mse", dict(SHIPPED_LOSS))
        b.max_t = 1000
        wa, wb = a.time_weights(t), b.time_weights(t)
        assert torch.allclose(wa, wb, rtol=2e-6, atol=1e-7), (t, wa, wb)
        ref = OL.time_weights(t, "snr", 0.1, 1.0).reshape(-1)
        assert torch.allclose(wb, ref, rtol=2e-6, atol=1e-7)


def test_groupnorm_epilogue_fusion_plan_equals_unfused(monkeypatch):
    """Host logic of dmu_conv_params.gn_fuse: with the library answering "yes" for the <= 8x8 layers, the plan drops those
    GroupNorm launches (forward: armed on the producing conv; backward: armed on the dgrad, weight gradient kept behind it,
    shortcut gradient computed first) and computes exactly what the unfused plan computes."""
    import diffusion_model_universal_b200 as D
    fake_device.install(monkeypatch)
    outs = []
    for fuse in (False, True):
        net = D.UNet(3, 64, 3, precision="fp32")
        net.load_state_dict(W.make_state_dict(W.unet_param_spec(64, 3, ""), 3))
        net.engine.fuse_gn_epi = fuse
        g = torch.Generator().manual_seed(5)
        x, t, dout = torch.randn(2, 3, 32, 32, generator=g), torch.randint(0, 1000, (2,), generator=g), torch.randn(2, 3, 32, 32, generator=g)
        y = net(x, t)
        y.backward(dout)
        plan = net.engine.get_plan(x.shape, True)
        assert (plan.gn_fused[0] > 20 and plan.gn_fused[1] > 20) == fuse
        n_gn = sum(1 for op in plan.fwd if op[0] is not None and getattr(op[0], "__name__", "") == "dmu_gn_forward")
        assert n_gn == 50 - plan.gn_fused[0]
        outs.append((y.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    assert rel_l2(outs[1][0], outs[0][0]) < 1e-6
    for k, g0 in outs[0][1].items():
        assert rel_l2(outs[1][1][k], g0) < 1e-5 or g0.norm() < 1e-9, k


def test_nccl_cta_cap_is_a_default_not_an_override():
    """parallel.py caps NCCL at 16 CTAs per collective (the all-reduces run beside the backward's kernels) unless the user has set
    the variable: run in fresh interpreters, because the environment is read at import."""
    import os, subprocess, sys
    code = "import os, diffusion_model_universal_b200.parallel; print(os.environ['NCCL_MAX_CTAS'])"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k != "NCCL_MAX_CTAS"}
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "16"
    env["NCCL_MAX_CTAS"] = "8"
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "8"


def test_forward_norms_ask_for_order_independent_statistics(monkeypatch):
    """Every GroupNorm of a forward plan carries DMU_GN_FIXED_SUMS and reserves the int64 accumulators behind its float sums (three
    times the bytes), so that the launches which add statistics with atomics stay bit-reproducible; DMU_GN_FIXED_SUMS=0 restores
    float atomics and the old footprint."""
    import diffusion_model_universal_b200 as D
    from diffusion_model_universal_b200 import _abi
    fake_device.install(monkeypatch)
    f = load_golden("unet_forward.pt")["c32_r32"]
    sizes = {}
    for fixed in ("1", "0"):
        monkeypatch.setenv("DMU_GN_FIXED_SUMS", fixed)
        net = D.UNet(3, f["C"], 3)
        net.load_state_dict(W.make_state_dict(W.unet_param_spec(f["C"], 3, ""), f["seed"]))
        with torch.no_grad():
            y = net(f["x"], f["t"])
        assert rel_l2(y, f["eps"]) < 1e-4
        plan = net.engine.get_plan(tuple(f["x"].shape), False)
        flags = [p.flags for sub in plan.keep for p in getattr(sub, "keep", []) if isinstance(p, _abi.GnParams)]
        assert len(flags) >= 50, "GroupNorm parameter blocks of the forward plan"
        assert all(fl == (_abi.GN_FIXED_SUMS if fixed == "1" else 0) for fl in flags)
        sizes[fixed] = plan.nbytes
    assert sizes["1"] > sizes["0"]
