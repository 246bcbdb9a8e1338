"""CPU tier: the C-ABI library builds, loads and exports every symbol include/dmu_b200.h declares."""

import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "dmu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from diffusion_model_universal_b200 import _abi, build_ext
    build_ext.build()
    assert os.path.exists(_abi.LIB_PATH)
    h = ctypes.CDLL(_abi.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/dmu_b200.h but not exported"
    # and the binding covers the whole header
    assert set(names) == set(_abi.EXPORTS), set(names) ^ set(_abi.EXPORTS)


def test_struct_mirrors_match():
    from diffusion_model_universal_b200 import _abi
    h = _abi.lib()   # runs the size checks, raises on drift
    for name, st in _abi.STRUCTS.items():
        assert h.dmu_sizeof(name.encode()) == ctypes.sizeof(st)
    assert h.dmu_sizeof(b"nope") == -1


def test_error_reporting_without_a_device():
    """Argument validation happens before any launch, so the error path is testable on CPU."""
    from diffusion_model_universal_b200 import _abi
    h = _abi.lib()
    assert h.dmu_q_sample(None, None, None, None, None, 1, 1, None) != 0
    assert b"null pointer" in h.dmu_last_error()
    p = _abi.ConvParams()
    assert h.dmu_conv2d(ctypes.byref(p), None) != 0
    assert h.dmu_sinusoidal_embedding(1, 0, 1, 1, 2, None) != 0   # dim 2 divides by zero in the reference formula
    assert h.dmu_ingest_u8(None, 1, None, None, None, None, None, None, None, 2, 3, 16, None) != 0
    assert b"null image pointer" in h.dmu_last_error()
    assert h.dmu_ingest_u8(None, 1, None, None, None, None, None, None, None, 0, 3, 16, None) == 0    # empty batch
    assert h.dmu_image_grid_u8(None, 4, 4, 48, 0, 3, 4, 4, 2, 2, 0.0, None, None) != 0
    assert h.dmu_image_grid_range_u8(1, 4, 4, 48, 0, 3, 4, 4, 2, 2, 0.0, 1.0, -1.0, 1, None) != 0
    assert b"empty value range" in h.dmu_last_error()
    gh, gw, gc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
    assert h.dmu_image_grid_shape(88, 3, 32, 32, 11, 2, ctypes.byref(gh), ctypes.byref(gw), ctypes.byref(gc)) == 0
    assert (gh.value, gw.value, gc.value) == (8 * 34 + 2, 11 * 34 + 2, 3)          # trainers/ddpm_trainer.py:832 grid
    assert h.dmu_image_grid_shape(0, 3, 32, 32, 11, 2, None, None, None) != 0
