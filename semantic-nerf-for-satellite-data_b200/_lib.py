"""ctypes binding of libsnb.so (the C ABI declared in include/snb.h).

This is the only way Python reaches the CUDA kernels.  There is no fallback: if the library is
missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libsnb.so")

MODEL_SATNERF, MODEL_SEMANTIC, MODEL_NERF, MODEL_SNERF = 0, 1, 2, 3
# the kind K1 (sample + encode) sees: raw xyz (SatNeRF, S-NeRF) or the 10-frequency positional encoding (semantic, NeRF)
K1_KIND = {0: 0, 1: 1, 2: 1, 3: 0}
HEADS_ALL, HEADS_SOLAR, HEADS_DEPTH = 63, 5, 1
COMPOSITE_NO_CLAMP, COMPOSITE_BETA_S = 1, 2
VARIANT_TJ_FOR_S, VARIANT_TJ_INSTEAD_OF_BETA, VARIANT_SEPARATE_BETA_S, VARIANT_SEPARATE_TJ_S, VARIANT_FULL_FEATURES = 1, 2, 4, 8, 16
VARIANT_RELU = 32
EPI_SIN, EPI_LINEAR, EPI_MUL, EPI_HEADOUT, EPI_F32ROWS, EPI_WGRAD = range(6)

_vp, _i, _i64, _u64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t

# name -> (restype, argtypes): must list every symbol include/snb.h declares (tests check this)
SIGNATURES = {
    "snb_version": (_i, []),
    "snb_last_error": (C.c_char_p, []),
    "snb_device_sms": (_i, []),
    "snb_sample_encode": (_i, [_vp, _vp, _vp, _u64, _vp, _u64, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                               _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_encode_points": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "snb_model_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "snb_model_destroy": (None, [_vp]),
    "snb_model_param_count": (_i64, [_vp]),
    "snb_model_num_tensors": (_i, [_vp]),
    "snb_model_tensor_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i)]),
    "snb_model_packed_bytes": (_sz, [_vp]),
    "snb_model_pack": (_i, [_vp, _vp, _vp, _vp]),
    "snb_mlp_workspace_bytes": (_sz, [_vp, _i64, _i]),
    "snb_mlp_forward": (_i, [_vp, _vp, _vp, _sz, _i64, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "snb_mlp_backward": (_i, [_vp, _vp, _vp, _sz, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "snb_mlp_forward_with_solar": (_i, [_vp, _vp, _vp, _sz, _i64, _i64, _vp, _vp, _vp, _i, _vp, _vp]),
    "snb_mlp_backward_with_solar": (_i, [_vp, _vp, _vp, _sz, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_model_grad_buckets": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "snb_nerf_aux": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "snb_mlp_fp32_workspace_bytes": (_sz, [_vp, _i64]),
    "snb_mlp_forward_fp32": (_i, [_vp, _vp, _vp, _sz, _i64, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "snb_ray_param_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "snb_composite_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_composite_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_composite_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "snb_label_counts": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "snb_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _vp, _f, _vp]),
    "snb_set_chained_mlp": (_i, [_i]),
    "snb_chain_schedule": (_i, [_i, _i, _i, _i, _i, C.POINTER(_i), _i]),
    "snb_profile_begin": (None, [_i]),
    "snb_profile_launch_count": (_i64, []),
    "snb_profile_add_launches": (None, [_i64]),
    "snb_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(C.c_double)]),
    "snb_gemm_bf16": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp, _i64, _vp, _vp, _f, _i, _vp]),
}



class LossParams(C.Structure):
    """snb_loss_params (include/snb.h)"""
    _fields_ = [("mode", _i), ("color", _i), ("beta_min", _f), ("inv_n", _f), ("lambda_s", _f), ("ignore_index", _i),
                ("lambda_c", _f), ("car_label", _i), ("lambda_sc", _f), ("lambda_ds", _f), ("flags", _i), ("sem_unc", _i)]


_lib = None


class SnbError(RuntimeError):
    pass


def load():
    """Load libsnb.so (once).  Raises if it has not been built - there is no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SnbError(f"{LIB_PATH} not found: run `python __graft_entry__.py` (build()) first; "
                           "the CUDA library is the only implementation of this path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().snb_last_error().decode(errors="replace")
        raise SnbError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """device pointer of a tensor (None -> NULL); refuses anything that is not a CUDA tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SnbError("libsnb has no CPU path: expected a CUDA tensor")
    if not t.is_contiguous():
        raise SnbError("libsnb expects contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
