"""The C-ABI shared library: loads, exports every symbol include/snb.h declares, and rejects bad
arguments with the documented error codes.  No compute is launched (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from semnerf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "snb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"libsnb.so does not export {s}"
    # and the ctypes table covers the header exactly
    assert set(syms) == set(_lib.SIGNATURES.keys())


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.snb_version() == 200
    assert isinstance(lib.snb_last_error(), bytes)


def test_model_layout_is_host_side():
    lib = _lib.load()
    for kind, C_, n in ((1, 6, 2827023), (1, 5, 2827023 - 257), (0, 0, 2635785)):
        h = C.c_void_p()
        assert lib.snb_model_create(C.byref(h), kind, C_, 1, 0, 4, 10) == 0
        assert lib.snb_model_param_count(h) == n          # SURVEY 8a row a6 parameter counts
        assert lib.snb_model_packed_bytes(h) > 2 * n       # forward + transposed bf16 copies
        w_inf = lib.snb_mlp_workspace_bytes(h, 65536, 0)
        w_tr = lib.snb_mlp_workspace_bytes(h, 65536, 1)
        assert 0 < w_inf < w_tr
        lib.snb_model_destroy(h)


def test_argument_validation_error_codes():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.snb_model_create(C.byref(h), 7, 6, 1, 0, 4, 10) == -1         # SNB_ERR_INVALID
    assert lib.snb_model_create(C.byref(h), 1, 11, 1, 0, 4, 10) == -2        # SNB_ERR_UNSUPPORTED (C > 10)
    assert b"n_classes" in lib.snb_last_error()
    assert lib.snb_model_create(C.byref(h), 1, 6, 1, 0, 13, 10) == -2        # the embedding must fit the 16 per-ray columns
    assert lib.snb_model_create(C.byref(h), 1, 6, 1, 8, 7, 10) == -2         # ... twice with a second embedding
    assert b"t_embedding_tau" in lib.snb_last_error()
    assert lib.snb_model_create(C.byref(h), 2, 0, 1, 16, 4, 10) == -2        # fc_use_full_features: SatNeRF / semantic model only
    assert lib.snb_model_create(C.byref(h), 1, 6, 1, 0, 4, 11) == -2         # K1 writes 10 positional frequencies
    assert lib.snb_model_create(C.byref(h), 0, 0, 1, 0, 4, 6) == -2          # ... and SatNeRF takes raw xyz
    # null pointers / S < 2 are rejected before any launch
    assert lib.snb_composite_forward(None, None, 4, 64, 15, 6, 0, None, None, None, None, None, None, None) == -1
    one = C.c_void_p(16)
    assert lib.snb_composite_forward(one, one, 4, 1, 15, 6, 0, one, one, one, one, one, one, None) == -2
    assert lib.snb_sample_encode(None, None, None, 0, None, 0, None, None, 0, 0, None, None, None, None, 0, 4, 64, 1, 0,
                                 None, None, None, None, None, None) == -1
    assert lib.snb_adam_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-8, 1, None, 1.0, None) == -1


def test_python_wrappers_refuse_cpu_tensors():
    import torch
    with pytest.raises(_lib.SnbError):
        _lib.ptr(torch.zeros(4))
