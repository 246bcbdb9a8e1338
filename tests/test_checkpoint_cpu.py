"""CPU tier: checkpoint interop with the reference trainer (trainers/ddpm_trainer.py:869-925).  The fused
optimizer's flat moment arenas are exchanged with torch.optim.Adam's per-parameter state through the reference's checkpoint
dict; host logic only (the C ABI is the host-memory test double)."""

import io

import pytest
import torch

import fake_device
from oracle import weights as W

CFG = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": 32, "loss_type": "mse",
       "loss_config": {"use_time_weighting": False}}


def _model(seed=7):
    import diffusion_model_universal_b200 as D
    m = D.DDPM(CFG)
    sd = m.state_dict()
    sd.update(W.make_state_dict(W.unet_param_spec(32, 3, "model."), seed))
    m.load_state_dict(sd)
    return m


def _fill_grads(m, seed):
    g = torch.Generator().manual_seed(seed)
    eng = m.model.engine
    for k, p in m.model.named_parameters():
        o, n = eng.offs[k]
        eng.gflat[o:o + n].copy_(torch.randn(n, generator=g) * 1e-2)


def test_checkpoint_round_trip_with_torch_adam(monkeypatch):
    from diffusion_model_universal_b200.trainer import TrainStep
    fake_device.install(monkeypatch)
    a = _model()
    ts = TrainStep(a, lr=1e-3, betas=(0.9, 0.99), eps=1e-8, ema_decay=0.99)
    empty = ts.checkpoint(epoch=0)
    assert empty["optimizer_state_dict"]["state"] == {} and empty["ema_model_state_dict"] is None
    a.model.engine.prepare(torch.device("cpu"))
    for s in (1, 2):
        _fill_grads(a, s)
        ts.opt.step()
    ck = ts.checkpoint(epoch=4, config={"model_config": CFG}, best_val_loss=0.25)
    assert set(ck) == {"epoch", "model_state_dict", "ema_model_state_dict", "optimizer_state_dict", "config", "best_val_loss",
                       "scheduler_state_dict"}
    buf = io.BytesIO()
    torch.save(ck, buf)                                            # what the reference trainer does with it
    ck = torch.load(io.BytesIO(buf.getvalue()), weights_only=False)

    # (1) the reference side: a plain module + torch.optim.Adam resume from it and take the same third step
    ref = _model(seed=99)
    ref.load_state_dict(ck["model_state_dict"])
    ropt = torch.optim.Adam(ref.parameters(), lr=5.0)              # lr etc. come from the checkpoint
    ropt.load_state_dict(ck["optimizer_state_dict"])
    assert ropt.param_groups[0]["lr"] == 1e-3 and ropt.param_groups[0]["betas"] == (0.9, 0.99)
    ema_ref = _model(seed=98)
    ema_ref.load_state_dict(ck["ema_model_state_dict"])           # same keys as a model state_dict (buffers included)
    _fill_grads(a, 3)
    names = [k for k, _ in a.model.named_parameters()]
    for (k, p) in ref.model.named_parameters():
        o, n = a.model.engine.offs[k]
        p.grad = a.model.engine.gflat[o:o + n].view(p.shape).clone()
    ts.opt.step()
    ropt.step()
    for (k, p), (_, q) in zip(a.model.named_parameters(), ref.model.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), k
    st = ropt.state_dict()["state"]
    assert len(st) == len(names) and all(float(s["step"]) == 3.0 for s in st.values())

    # (2) the other direction: the reference's optimizer / EMA state loads into a fresh fused optimizer
    back = {"epoch": 9, "model_state_dict": ref.state_dict(), "ema_model_state_dict": ema_ref.state_dict(),
            "optimizer_state_dict": ropt.state_dict(), "config": None, "best_val_loss": 0.1, "scheduler_state_dict": None}
    b = _model(seed=5)
    tb = TrainStep(b, lr=7.0, ema_decay=0.99)
    assert tb.load_checkpoint(back) == 9
    assert tb.opt.step_count == 3 and tb.opt.lr == 1e-3 and tb.opt.betas == (0.9, 0.99)
    eng_a, eng_b = a.model.engine, b.model.engine
    for k in names:
        (oa, n), (ob, _) = eng_a.offs[k], eng_b.offs[k]
        assert torch.allclose(tb.opt.m[ob:ob + n], ts.opt.m[oa:oa + n], rtol=1e-5, atol=1e-9), k
        assert torch.allclose(tb.opt.v[ob:ob + n], ts.opt.v[oa:oa + n], rtol=1e-5, atol=1e-12), k
        assert torch.equal(tb.opt.ema[ob:ob + n], ema_ref.state_dict()["model." + k].reshape(-1)), k
        assert torch.equal(eng_b.flat[ob:ob + n], ref.state_dict()["model." + k].reshape(-1)), k

    # malformed optimizer states fail loudly
    bad = ropt.state_dict()
    bad["state"][0]["step"] = torch.tensor(7.0)
    with pytest.raises(ValueError):
        tb.opt.load_state_dict(bad)
    bad = ropt.state_dict()
    bad["param_groups"][0]["amsgrad"] = True
    with pytest.raises(ValueError):
        tb.opt.load_state_dict(bad)
    bad = ropt.state_dict()
    bad["param_groups"][0]["params"] = bad["param_groups"][0]["params"][:-1]
    with pytest.raises(ValueError):
        tb.opt.load_state_dict(bad)
