"""GPU parity of the normalised sample grid (scripts/generate.py:119-133: save_image(..., normalize=True, value_range=(-1, 1)))
against the torchvision-pinned oracle (tests/test_pipeline_cpu.py)."""

import pytest
import torch

from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,c,nrow,rng", [(16, 3, 4, (-1.0, 1.0)), (1, 3, 8, (-1.0, 1.0)), (7, 1, 3, (0.0, 1.0)), (9, 3, 3, (-0.5, 2.0))])
def test_image_grid_value_range_bit_exact(n, c, nrow, rng):
    from diffusion_model_universal_b200 import ops, pipeline
    dev = torch.device("cuda:0")
    x = torch.randn(n, c, 32, 32, generator=torch.Generator().manual_seed(n)) * 1.5
    x[0, 0, 0, :4] = torch.tensor([rng[0], rng[1], rng[0] - 1, rng[1] + 1])
    ref = OP.to_u8_hwc(OP.make_grid(x, nrow, 2, 0.0, value_range=rng))
    got = ops.image_grid_u8(x.to(dev), nrow, 2, 0.0, value_range=rng)
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    assert torch.equal(pipeline.image_grid(x.to(dev), nrow=nrow, value_range=rng).cpu(), ref)
