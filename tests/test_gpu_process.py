"""GPU parity: fused process/update/loss kernels vs the oracle and the reference-generated fixtures."""

import math

import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import process as P, losses as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _ops():
    from diffusion_model_universal_b200 import ops
    return ops


def test_q_sample_and_ddpm_step_golden(dev):
    ops = _ops()
    f = load_golden("ddpm_chain.pt")
    b, a, acp = (t.to(dev) for t in P.linear_schedule(1e-4, 0.02, 1000))
    s = f["step"]
    x, eps = s["x"].to(dev), s["eps"].to(dev)
    for tt, c in s["cases"].items():
        t = torch.full((3,), tt, dtype=torch.long, device=dev)
        out = ops.ddpm_step(x, eps, t, c["noise"].to(dev) if tt > 0 else None, b, a, acp)
        # tolerance: a few ulp (pow(alpha,-0.5) is not correctly rounded on either side)
        assert rel_l2(out, c["out"]) < 5e-7, tt
    q = f["q_sample"]
    out = ops.q_sample(x, q["t"].to(dev), q["noise"].to(dev), acp)
    assert torch.equal(out.cpu(), q["out"])  # mul/mul/add/sqrt are correctly rounded: bit-exact


def test_ddpm_step_device_branch_and_sizes(dev):
    ops = _ops()
    b, a, acp = (t.to(dev) for t in P.linear_schedule(1e-4, 0.02, 1000))
    g = torch.Generator().manual_seed(3)
    for shape in ((1, 3, 32, 32), (5, 3, 7, 5), (128, 3, 32, 32), (2, 1, 1, 1)):   # ragged inner sizes hit the scalar path
        x, e, z = (torch.randn(shape, generator=g) for _ in range(3))
        for tt in (0, 1, 999):
            t = torch.full((shape[0],), tt, dtype=torch.long)
            ref = P.ddpm_reverse_step(x, e, t, z, *P.linear_schedule(1e-4, 0.02, 1000))
            out = ops.ddpm_step(x.to(dev), e.to(dev), t.to(dev), z.to(dev), b, a, acp)   # noise passed even at t == 0: must be ignored
            assert rel_l2(out, ref) < 5e-7, (shape, tt)
    # in-place form used by the sampling loop
    x = torch.randn(4, 3, 32, 32, generator=g).to(dev)
    e, z = torch.randn_like(x), torch.randn_like(x)
    t = torch.full((4,), 17, dtype=torch.long, device=dev)
    want = ops.ddpm_step(x, e, t, z, b, a, acp)
    got = ops.ddpm_step(x, e, t, z, b, a, acp, out=x)
    assert torch.equal(want, got)


def test_ddpm_step_mixed_batch_and_index_dtype(dev):
    """A batch whose rows carry different timesteps: the t[0] > 0 branch applies to every row (ddpm.py:311,323) and a t == 0 row
    reads alphas_cumprod[-1] = alphas_cumprod[T-1] like the reference's tensor index.  int32 timesteps are rejected, not
    reinterpreted."""
    ops = _ops()
    sched = P.linear_schedule(1e-4, 0.02, 1000)
    b, a, acp = (t.to(dev) for t in sched)
    g = torch.Generator().manual_seed(12)
    x, e, z = (torch.randn(4, 3, 8, 8, generator=g) for _ in range(3))
    for tv in ([5, 0, 999, 1], [0, 7, 0, 3]):
        t = torch.tensor(tv, dtype=torch.long)
        ref = P.ddpm_reverse_step(x, e, t, z, *sched)
        out = ops.ddpm_step(x.to(dev), e.to(dev), t.to(dev), z.to(dev), b, a, acp)
        assert rel_l2(out, ref) < 5e-7, tv
    t = torch.tensor([5, 0, 999, 1], device=dev)
    with pytest.raises(TypeError):
        ops.q_sample(x.to(dev), t.int(), z.to(dev), acp)
    with pytest.raises(TypeError):
        ops.ddpm_step(x.to(dev), e.to(dev), t.int(), z.to(dev), b, a, acp)
    with pytest.raises(ValueError):
        ops.q_sample(x.to(dev), t[:3], z.to(dev), acp)


def test_empty_batch(dev):
    ops = _ops()
    b, a, acp = (t.to(dev) for t in P.linear_schedule(1e-4, 0.02, 1000))
    x = torch.empty(0, 3, 32, 32, device=dev)
    t = torch.empty(0, dtype=torch.long, device=dev)
    assert ops.q_sample(x, t, x, acp).shape == x.shape
    assert ops.ddpm_step(x, x, t, None, b, a, acp).shape == x.shape


def test_ddim_step_golden(dev):
    ops = _ops()
    f = load_golden("ddim_chain.pt")
    _, _, acp = P.linear_schedule(1e-4, 0.02, 1000)
    ts = P.ddim_timesteps(1000, 50)
    s = f["step"]
    tab = [t.to(dev) for t in P.ddim_tables(acp, ts, s["eta"])]
    for i, c in s["cases"].items():
        out = ops.ddim_step(s["x"].to(dev), s["eps"].to(dev), torch.full((3,), i, device=dev), c["noise"].to(dev), *tab)
        assert torch.equal(out.cpu(), c["out"]), i   # only correctly-rounded ops: bit-exact
    tab0 = P.ddim_tables(acp, ts, 0.0)
    g = torch.Generator().manual_seed(4)
    x, e = torch.randn(6, 3, 16, 16, generator=g) * 3, torch.randn(6, 3, 16, 16, generator=g)
    idx = torch.tensor([0, 49, 13, 7, 30, 1])
    ref = P.ddim_step(x, e, idx, tab0, 0.0)
    out = ops.ddim_step(x.to(dev), e.to(dev), idx.to(dev), None, *[t.to(dev) for t in tab0])
    assert torch.equal(out.cpu(), ref)


def test_langevin_and_renoise(dev):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x, s, z = (torch.randn(3, 3, 8, 8, generator=g) for _ in range(3))
    sig = P.score_sigma_ladder(0.01, 50.0, 10)
    for k in (0, 4, 9):
        ref = P.score_langevin_step(x, s, z, sig[k], 0.7)
        out = ops.langevin_score_step(x.to(dev), s.to(dev), z.to(dev), sig.to(dev), k, 0.7)
        assert rel_l2(out, ref) < 1e-6, k
    ref = P.energy_langevin_step(x, s, z, 0.01)
    out = ops.langevin_energy_step(x.to(dev), s.to(dev), z.to(dev), 0.01)
    assert rel_l2(out, ref) < 1e-6
    _, _, acp = P.linear_schedule(1e-4, 0.02, 1000)
    for t in (1, 500, 999):
        ref = P.energy_renoise(x, z, acp, t)
        out = ops.energy_renoise(x.to(dev), z.to(dev), acp.to(dev), t)
        assert rel_l2(out, ref) < 1e-6, t
    with pytest.raises(RuntimeError):
        ops.energy_renoise(x.to(dev), z.to(dev), acp.to(dev), 0)
    sg = torch.tensor([0.3, 2.0, 40.0])
    out = ops.scale_add(x.to(dev), z.to(dev), None, sg.to(dev))
    assert rel_l2(out, x + sg[:, None, None, None] * z) < 1e-7


def test_loss_variants_golden(dev):
    from diffusion_model_universal_b200.losses import DiffusionLoss
    f = load_golden("losses.pt")
    for c in f["cases"]:
        fn = DiffusionLoss(c["loss_type"], dict(c["cfg"]))
        p = f["pred"].to(dev).requires_grad_(True)
        v = fn(p, f["target"].to(dev), f["t"].to(dev))
        assert abs(v.item() - c["value"].item()) <= 2e-6 * max(1.0, abs(c["value"].item())), c["cfg"]
        v.backward()
        if c["dpred"].norm() > 0:
            assert rel_l2(p.grad, c["dpred"]) < 1e-6
        else:
            assert p.grad.abs().max().item() == 0
    with pytest.raises(ValueError):
        DiffusionLoss("hybrid", {})(f["pred"].to(dev), f["target"].to(dev), f["t"].to(dev))   # losses.py:113-114


def test_loss_full_size_properties(dev):
    """BASELINE config-2 size (128x3x32x32): scale linearity and grad == autograd of the oracle."""
    from diffusion_model_universal_b200 import ops
    g = torch.Generator().manual_seed(6)
    p, t = torch.randn(128, 3, 32, 32, generator=g), torch.randn(128, 3, 32, 32, generator=g)
    w = torch.rand(128, generator=g)
    l1, d1 = ops.diffusion_loss(p.to(dev), t.to(dev), w.to(dev), 1.0, 0.0, 0.0, 1.0, True)
    l2, d2 = ops.diffusion_loss(p.to(dev), t.to(dev), (2 * w).to(dev), 1.0, 0.0, 0.0, 1.0, True)
    assert abs(l2.item() - 2 * l1.item()) < 1e-6 * abs(l2.item())
    pr = p.clone().requires_grad_(True)
    ref = ((pr - t) ** 2 * w.view(-1, 1, 1, 1)).mean()
    ref.backward()
    assert abs(l1.item() - ref.item()) < 1e-6 * ref.item()
    assert rel_l2(d1, pr.grad) < 1e-6
    # deterministic: fixed-order reduction
    l3, _ = ops.diffusion_loss(p.to(dev), t.to(dev), w.to(dev), 1.0, 0.0, 0.0, 1.0, False)
    assert l3.item() == l1.item()


def test_cpu_tensor_is_rejected():
    from diffusion_model_universal_b200 import ops
    x = torch.randn(2, 3, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.q_sample(x, torch.zeros(2, dtype=torch.long), x, torch.ones(10))


@pytest.mark.parametrize("tag,precision,tol", [("c16", "fp32", 2e-4), ("c64", "fp32", 2e-4), ("c64", "bf16", 3e-2)])
def test_energy_model_against_live_reference_draws(tag, precision, tol):
    """EnergyNet on the CUDA library: energies, Langevin chain, input gradient, loss value (with the gradient penalty) and
    the contrastive-divergence parameter gradients, driven with the reference's recorded random draws."""
    import diffusion_model_universal_b200 as D
    from oracle import weights as OW
    r = load_golden("energy_draws.pt")[tag]
    dev = torch.device("cuda:0")
    cfg = dict(r["cfg"], precision=precision)
    m = D.EnergyBasedDiffusion(cfg)
    sd = m.state_dict()
    sd.update(OW.make_state_dict(OW.energynet_param_spec(r["C"], 3, "model."), r["wseed"]))
    m.load_state_dict(sd)
    m.to(dev)
    x, t, noise, alpha = r["x"].to(dev), r["t"].to(dev), r["noise"].to(dev), r["alpha"].to(dev)
    lang = [z.to(dev) for z in r["lang"]]
    with torch.no_grad():
        assert rel_l2(m.forward(x), r["energy"]) < tol
        xf = m._langevin_sampling(m._add_noise(x, t, noise), t, _noises=lang)
        assert rel_l2(xf, r["x_fake"]) < tol
        xh = alpha * x + (1 - alpha) * r["x_fake"].to(dev)
        _, g = m.model.energy_and_input_grad(xh)
        assert rel_l2(g, r["grad_x_hat"]) < 10 * tol
    loss = m._loss_from_draws(x, t, noise, lang, alpha)
    assert abs(loss.item() - float(r["loss_gp"])) < 10 * tol * max(abs(float(r["loss_gp"])), 1e-2)
    loss.backward()                                       # CD terms + 0.01 x gradient penalty (double backward)
    scale = max(r["grads_gp_norm"].values())
    for k, p in m.named_parameters():
        gn = r["grads_gp_norm"][k]
        assert p.grad is not None, k
        if gn > 1e-5 * scale:
            assert abs(p.grad.norm().item() - gn) < 20 * tol * gn + 1e-4 * scale, (k, p.grad.norm().item(), gn)
        if k in r["grads_gp"] and gn > 1e-3 * scale:
            assert rel_l2(p.grad, r["grads_gp"][k]) < 20 * tol, k
    m.zero_grad()
    m.regularization_weight = 0.0
    loss = m._loss_from_draws(x, t, noise, lang, alpha)
    assert abs(loss.item() - float(r["loss_cd"])) < 10 * tol * max(abs(float(r["loss_cd"])), 1e-2)
    loss.backward()
    scale = max(r["grads_cd_norm"].values())
    for k, p in m.named_parameters():
        gn = r["grads_cd_norm"][k]
        assert p.grad is not None, k
        if gn > 1e-5 * scale:
            assert abs(p.grad.norm().item() - gn) < 20 * tol * gn + 1e-4 * scale, (k, p.grad.norm().item(), gn)
        if k in r["grads_cd"] and gn > 1e-3 * scale:
            assert rel_l2(p.grad, r["grads_cd"][k]) < 20 * tol, k
    with torch.no_grad():
        s = m.__class__(dict(cfg, num_timesteps=3, langevin_steps=2)).to(dev)
        s.model.load_state_dict(m.model.state_dict())
        out = s.generate_samples(2, dev)
        assert out.shape == (2, 3, r["R"], r["R"]) and torch.isfinite(out).all()


def test_trainstep_with_energy_model(dev):
    """TrainStep over a network without a launch-plan engine (EnergyNet).  Its flat arenas are built by the trainer:
    the gradients of one step equal loss_function(x).backward() from the same RNG state, and the fused Adam moves the weights."""
    import diffusion_model_universal_b200 as D
    from diffusion_model_universal_b200.trainer import TrainStep
    from oracle import weights as W
    cfg = {"num_timesteps": 1000, "beta_start": 1e-4, "beta_end": 0.02, "use_time_conditioning": False, "in_channels": 3,
           "model_channels": 16, "image_size": 16, "image_channels": 3, "loss_type": "energy_based", "regularization_weight": 0.01,
           "langevin_steps": 3, "langevin_step_size": 0.01}
    x = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(2)).to(dev)

    def make():
        m = D.EnergyBasedDiffusion(dict(cfg))
        sd = m.state_dict()
        sd.update(W.make_state_dict(W.energynet_param_spec(16, 3, "model."), 62))
        m.load_state_dict(sd)
        return m.to(dev)
    a = make()
    torch.manual_seed(3)
    la = a.loss_function(x)
    la.backward()
    ga = {k: p.grad.clone() for k, p in a.named_parameters()}
    b = make()
    ts = TrainStep(b, lr=0.0, ema_decay=None)
    torch.manual_seed(3)
    lb = ts.step(x)
    assert abs(lb.item() - la.item()) < 1e-4 * max(1.0, abs(la.item()))
    eng = ts._net.engine
    for k, (o, n) in eng.offs.items():
        g = eng.gflat[o:o + n].view(eng.named[k].shape)
        ref = ga["model." + k]
        assert rel_l2(g, ref) < 1e-3 or ref.norm() < 1e-7, k
    before = eng.flat.clone()
    ts2 = TrainStep(b, lr=1e-3, ema_decay=0.99)
    for _ in range(3):
        assert torch.isfinite(ts2.step(x))
    assert not torch.equal(ts2._net.engine.flat, before)
    assert ts2.opt.ema is not None and torch.isfinite(ts2.opt.ema).all()
    ck = ts2.checkpoint(epoch=1)
    assert set(ck["model_state_dict"]) == set(b.state_dict())


def test_snr_time_weights_kernel_matches_reference_formula(dev):
    """dmu_snr_time_weights (one launch, t_max on the device) against utils/losses.py:144-181 evaluated with the reference's own
    torch calls on the same device (host-synchronising .item() included): same IEEE operations in the same order."""
    from diffusion_model_universal_b200.losses import DiffusionLoss
    cfg = {"use_time_weighting": True, "time_weight_type": "snr", "time_weight_params": {"min_weight": 0.1, "max_weight": 1.0}}
    fused = DiffusionLoss("mse", dict(cfg))
    fused.max_t = 1000
    plain = DiffusionLoss("mse", dict(cfg))          # max_t None: linspace(1e-4, 2e-2, t.max().item() + 1) like the reference
    g = torch.Generator().manual_seed(3)
    cases = [torch.randint(0, 1000, (128,), generator=g), torch.randint(0, 10, (16,), generator=g), torch.zeros(4, dtype=torch.long),
             torch.tensor([1, 0, 1]), torch.tensor([999]), torch.tensor([5, 5, 5, 5]), torch.arange(0, 1000, 37),
             torch.randint(0, 1000, (1000,), generator=g)]
    worst = 0.0
    for t in cases:
        td = t.to(dev)
        a, b = fused.time_weights(td), plain.time_weights(td)
        assert a.shape == b.shape and a.dtype == torch.float32
        worst = max(worst, float((a - b).abs().max()))
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (t, a, b)
        ref = L.time_weights(t, "snr", 0.1, 1.0).reshape(-1)          # the oracle on the CPU (its cumprod rounds differently from the GPU's scan)
        assert torch.allclose(a.cpu(), ref, rtol=1e-4, atol=2e-6), float((a.cpu() - ref).abs().max())
    print(f"snr time weights: largest |kernel - torch ops| = {worst:.3e}")
