"""CPU tier, world_size 2 over gloo: the data-parallel step of the package (flat gradient arena -> bucketed all-reduce ->
fused Adam) equals the single-process step on the concatenated batch, and the ranks stay bit-identical.  Runs against the
host-memory test double of the C ABI (tests/fake_device.py): what is under test is the host logic, not the kernels."""

import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup_fake():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_device
    from diffusion_model_universal_b200 import _abi, ops
    fake = fake_device.FakeLib()
    _abi.lib = lambda: fake
    ops._need_cuda = lambda *a: None
    ops._stream = lambda: None


def _model(C=32):
    import diffusion_model_universal_b200 as D
    from oracle import weights as W
    cfg = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": C, "loss_type": "mse",
           "loss_config": {"use_time_weighting": False}}
    m = D.DDPM(cfg)
    sd = m.state_dict()
    sd.update(W.make_state_dict(W.unet_param_spec(C, 3, "model."), 5))
    m.load_state_dict(sd)
    return m


def _shard_grads(m, x, t, noise, between=None):
    """gradient arena of mean-MSE on one shard, through the engine (no RNG: t and noise injected)"""
    from diffusion_model_universal_b200 import ops
    eng = m.model.engine
    eng.prepare(x.device)
    plan = eng.get_plan(x.shape, True)
    eps = eng.run_forward(m._add_noise(x, t, noise), t, plan)
    loss, dpred = ops.diffusion_loss(eps, noise, None, 1.0, 0.0, 0.0, 1.0, True)
    eng.run_backward(plan, dpred, between=between)
    return loss, eng.gflat


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    _setup_fake()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffusion_model_universal_b200.parallel import GradAllReducer
    from diffusion_model_universal_b200.optim import FusedAdamEMA
    g = torch.Generator().manual_seed(100)
    x = torch.randn(4, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (4,), generator=g)
    noise = torch.randn(4, 3, 32, 32, generator=g)
    lo, hi = rank * 2, rank * 2 + 2
    m = _model()
    red = GradAllReducer(m.model, bucket_mb=1.0)      # several buckets
    eng = m.model.engine
    works, snap = [], {}

    def between(a, b):
        # gflat[a:b] must be final here: start reducing it while the next part of the backward still runs
        snap[(a, b)] = float(eng.gflat[a:b].abs().sum())
        works.extend(red.launch(a, b))

    loss, gflat = _shard_grads(m, x[lo:hi], t[lo:hi], noise[lo:hi], between=between)
    scale = red.finish(works)
    assert len(snap) == 3 and all(v > 0 for v in snap.values())
    assert sorted(snap)[0][0] == 0 and sum(b - a for a, b in snap) == gflat.numel()
    opt = FusedAdamEMA(m.model, lr=1e-3, ema_decay=0.99)
    opt.step(grad_scale=scale)
    torch.save({"g": gflat.clone() * scale, "p": m.model.engine.flat.clone(), "ema": opt.ema.clone(), "loss": loss.clone()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_ranges_cover_the_arena_tail_first():
    sys.path.insert(0, ROOT)
    from diffusion_model_universal_b200.parallel import bucket_ranges
    for total, b in [(10, 3), (9, 3), (1, 5), (0, 4), (15_909_955, 4 << 20)]:
        r = bucket_ranges(total, b)
        assert sum(hi - lo for lo, hi in r) == total
        assert all(r[i][0] == r[i + 1][1] for i in range(len(r) - 1))          # contiguous, issued from the tail
        assert (not r) or (r[0][1] == total and r[-1][0] == 0)
        assert all(hi - lo <= b for lo, hi in r)


@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_process_mean(tmp_path, monkeypatch):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["g"], r1["g"]) and torch.equal(r0["p"], r1["p"]) and torch.equal(r0["ema"], r1["ema"])
    # single-process emulation: mean of the two shard gradients == gradient of the mean loss over the full batch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_device
    fake_device.install(monkeypatch)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(4, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (4,), generator=g)
    noise = torch.randn(4, 3, 32, 32, generator=g)
    m = _model()
    loss, gflat = _shard_grads(m, x, t, noise)
    ref = gflat.clone()
    rel = float((r0["g"] - ref).norm() / ref.norm())
    assert rel < 1e-5, rel
    assert abs(float(loss) - 0.5 * (float(r0["loss"]) + float(r1["loss"]))) < 1e-5 * abs(float(loss))
