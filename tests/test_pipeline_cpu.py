"""CPU tier for the ingest / sample-formatting row: the oracle restatement is pinned against the installed
torchvision (the third-party code the reference calls, datasets/dataset_utils.py:58-61, trainers/ddpm_trainer.py:821-834),
and the host logic (wrapper addressing, TrainStep's uint8 route) runs on the host-memory test double."""

import pytest
import torch

import fake_device
from oracle import pipeline as OP
from oracle import weights as W

MEAN, STD = (0.5, 0.4, 0.3), (0.5, 0.25, 0.2)


def _bytes(shape, seed):
    return torch.randint(0, 256, shape, generator=torch.Generator().manual_seed(seed), dtype=torch.uint8)


def test_oracle_ingest_matches_torchvision_bit_for_bit():
    tvf = pytest.importorskip("torchvision.transforms.functional")
    T = pytest.importorskip("torchvision.transforms")
    np = pytest.importorskip("numpy")
    batch = _bytes((3, 8, 6, 3), 0)
    tf = T.Compose([T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])        # what dataset_utils.py builds
    ref = torch.stack([tf(np.asarray(img)) for img in batch.numpy()])
    assert torch.equal(OP.ingest(batch, MEAN, STD, "NHWC"), ref)
    assert torch.equal(OP.ingest(batch.permute(0, 3, 1, 2).contiguous(), MEAN, STD, "NCHW"), ref)
    assert torch.equal(OP.ingest(batch, None, None, "NHWC"), torch.stack([tvf.to_tensor(np.asarray(i)) for i in batch.numpy()]))
    # every byte value, the shipped (0.5, 0.5) normalisation: [-1, 1] end points exact
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 16, 16, 1)
    x = OP.ingest(ramp, (0.5,), (0.5,), "NHWC")
    assert x.min() == -1.0 and x.max() == 1.0


@pytest.mark.parametrize("n,c,nrow,pad,pv", [(8, 3, 4, 2, 0.0), (7, 3, 3, 1, 0.5), (1, 3, 8, 2, 0.0), (5, 1, 8, 2, 1.0), (6, 3, 11, 0, 0.0)])
def test_oracle_grid_matches_torchvision_bit_for_bit(n, c, nrow, pad, pv, tmp_path):
    tvu = pytest.importorskip("torchvision.utils")
    x = torch.randn(n, c, 5, 7, generator=torch.Generator().manual_seed(n)) * 0.7 + 0.5     # values on both sides of [0, 1]
    g = OP.make_grid(x, nrow, pad, pv)
    assert torch.equal(g, tvu.make_grid(x, nrow=nrow, padding=pad, pad_value=pv))
    Image = pytest.importorskip("PIL.Image")
    np = pytest.importorskip("numpy")
    path = tmp_path / "g.png"
    tvu.save_image(x, path, nrow=nrow, padding=pad, pad_value=pv)                           # lossless PNG round trip
    assert torch.equal(OP.to_u8_hwc(g), torch.from_numpy(np.array(Image.open(path))))


@pytest.mark.parametrize("rng", [(-1.0, 1.0), (0.0, 1.0), (-0.5, 2.0)])
def test_oracle_normalised_grid_matches_torchvision(rng, tmp_path):
    """scripts/generate.py:119-133: save_image(samples, nrow=, normalize=True, value_range=(-1, 1))."""
    tvu = pytest.importorskip("torchvision.utils")
    Image = pytest.importorskip("PIL.Image")
    np = pytest.importorskip("numpy")
    x = torch.randn(9, 3, 6, 5, generator=torch.Generator().manual_seed(4)) * 1.5
    g = OP.make_grid(x, 3, 2, 0.0, value_range=rng)
    assert torch.equal(g, tvu.make_grid(x, nrow=3, padding=2, normalize=True, value_range=rng))
    path = tmp_path / "g.png"
    tvu.save_image(x, path, nrow=3, normalize=True, value_range=rng)
    assert torch.equal(OP.to_u8_hwc(g), torch.from_numpy(np.array(Image.open(path))))
    one = tvu.make_grid(x[0], normalize=True, value_range=rng)          # single images, as generate.py's per-sample loop
    assert torch.equal(OP.make_grid(x[:1], value_range=rng), one)


def test_wrappers_address_images_like_the_trainer(monkeypatch):
    from diffusion_model_universal_b200 import ops
    fake_device.install(monkeypatch)
    g = torch.Generator().manual_seed(3)
    inter = [torch.rand(4, 3, 6, 5, generator=g) for _ in range(11)]            # 11 saved steps of 4 samples
    ref = OP.to_u8_hwc(OP.make_grid(OP.denoising_rows(inter), nrow=11, padding=2))
    got = ops.image_grid_u8(torch.stack(inter), nrow=11, padding=2, transpose=True)
    assert got.shape == ref.shape and torch.equal(got, ref)
    assert ops.image_grid_shape(44, 3, 6, 5, 11, 2) == tuple(ref.shape)
    x = torch.rand(7, 1, 4, 4, generator=g)
    assert torch.equal(ops.image_grid_u8(x, nrow=3, padding=1, pad_value=0.25), OP.to_u8_hwc(OP.make_grid(x, 3, 1, 0.25)))
    assert torch.equal(ops.image_grid_u8(x[:1]), OP.to_u8_hwc(OP.make_grid(x[:1])))
    y = torch.randn(5, 3, 4, 4, generator=g)
    assert torch.equal(ops.image_grid_u8(y, nrow=2, value_range=(-1, 1)), OP.to_u8_hwc(OP.make_grid(y, 2, 2, 0.0, (-1, 1))))
    b = _bytes((2, 6, 4, 3), 1)
    mean, std = torch.tensor(MEAN), torch.tensor(STD)
    x0, xt = ops.ingest_u8(b, mean, std, "NHWC")
    assert xt is None and torch.equal(x0, OP.ingest(b, MEAN, STD, "NHWC"))
    with pytest.raises(TypeError):
        ops.ingest_u8(b.float(), mean, std, "NHWC")
    with pytest.raises(ValueError):
        ops.ingest_u8(b, mean, std, "HWCN")
    with pytest.raises(ValueError):
        ops.ingest_u8(b, mean, std, "NHWC", want_x0=False)
    with pytest.raises(TypeError):
        ops.ingest_u8(b, mean[:2], std, "NHWC")


def test_trainstep_takes_image_bytes(monkeypatch):
    """uint8 batch + input_norm == the fp32 batch the reference's DataLoader would have produced: same RNG draws, same loss,
    same parameters after the step."""
    import diffusion_model_universal_b200 as D
    from diffusion_model_universal_b200.trainer import TrainStep
    fake_device.install(monkeypatch)
    cfg = {"beta_start": 1e-4, "beta_end": 0.02, "image_size": 32, "image_channels": 3, "model_channels": 32, "loss_type": "mse",
           "loss_config": {"use_time_weighting": True, "time_weight_type": "snr", "time_weight_params": {"min_weight": 0.1, "max_weight": 1.0}}}
    b = _bytes((2, 32, 32, 3), 5)
    out = []
    for kind in ("u8", "f32"):
        m = D.DDPM(cfg)
        sd = m.state_dict()
        sd.update(W.make_state_dict(W.unet_param_spec(32, 3, "model."), 7))
        m.load_state_dict(sd)
        ts = TrainStep(m, lr=1e-3, input_norm=(MEAN, STD), input_layout="NHWC")
        torch.manual_seed(11)
        loss = ts.step(b if kind == "u8" else OP.ingest(b, MEAN, STD, "NHWC"))
        out.append((loss.clone(), m.model.engine.flat.clone()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    with pytest.raises(TypeError):
        ts.step(b.to(torch.int32))
    with pytest.raises(ValueError):
        TrainStep(m, input_layout="HWC")


def test_trainstep_image_bytes_on_the_generic_route(monkeypatch):
    """Score-based model: TrainStep has no fused front end for it, so the bytes are normalised first (one launch) and the
    model's own loss_function runs on the result — same step as with the float batch."""
    import diffusion_model_universal_b200 as D
    from diffusion_model_universal_b200.trainer import TrainStep
    from conftest import load_golden
    fake_device.install(monkeypatch)
    f = load_golden("score.pt")
    b = _bytes((2, 3, 32, 32), 8)
    out = []
    for kind in ("u8", "f32"):
        m = D.ScoreBasedDiffusion(dict(f["cfg"]))
        sd = m.state_dict()
        sd.update(W.make_state_dict(W.scorenet_param_spec(f["C"], 3, "model."), f["wseed"]))
        m.load_state_dict(sd)
        ts = TrainStep(m, lr=1e-3, input_norm=(MEAN, STD), input_layout="NCHW", ema_decay=None)
        torch.manual_seed(3)
        loss = ts.step(b if kind == "u8" else OP.ingest(b, MEAN, STD, "NCHW"))
        out.append((loss.clone(), m.model.engine.flat.clone()))
    assert torch.isfinite(out[0][0]) and torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_grid_shape_entry_point_matches_torchvision_for_random_geometries():
    """dmu_image_grid_shape is host-only arithmetic in the real library: check it against make_grid's output shape."""
    import ctypes
    import random
    tvu = pytest.importorskip("torchvision.utils")
    from diffusion_model_universal_b200 import _abi
    h = _abi.lib()
    rnd = random.Random(0)
    for _ in range(200):
        n, c = rnd.randint(1, 40), rnd.choice([1, 3, 4])
        hh, ww, nrow, pad = rnd.randint(1, 9), rnd.randint(1, 9), rnd.randint(1, 12), rnd.randint(0, 3)
        ref = tvu.make_grid(torch.zeros(n, c, hh, ww), nrow=nrow, padding=pad)
        gh, gw, gc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        assert h.dmu_image_grid_shape(n, c, hh, ww, nrow, pad, ctypes.byref(gh), ctypes.byref(gw), ctypes.byref(gc)) == 0
        assert (gc.value, gh.value, gw.value) == tuple(ref.shape), (n, c, hh, ww, nrow, pad)
