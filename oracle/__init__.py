"""CPU oracle for the UNet-denoiser hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (ATen, fp32) functional restatement of the
reference's algorithm for the path named in BASELINE.json: the UNet denoiser
and the DDPM / DDIM / score / energy step arithmetic around it.  Every function
cites the reference file:line it follows (paths relative to the upstream
repository root).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  Nothing under ``diffusion_model_universal_b200/`` imports
it: the product path is CUDA-only and raises if its extension is missing.

Parity pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the live reference executed in the
build container.  ``tests/golden/make_golden.py`` (committed) loads weights
made by :func:`oracle.weights.make_state_dict` into the unmodified reference
classes, runs them and stores the small input/output fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this package against
them on CPU.
"""

from . import weights, unet, process, losses  # noqa: F401
