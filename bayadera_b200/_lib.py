"""ctypes binding of libbayadera_b200.so (include/bayadera_b200.h).

The library is the product; there is no Python or CPU fallback.  If the shared object is
missing this module raises immediately, and every compute call fails with ``BayaderaError``
when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

# BAYADERA_B200_LIB: an alternative build of the same library (kernel experiments); default = the in-tree build
LIB_PATH = Path(os.environ.get("BAYADERA_B200_LIB") or Path(__file__).resolve().parent / "libbayadera_b200.so")

OK, EINVAL, EINVAL_WALKERS, EACOR_TOO_SHORT, ECOMPILE, ECUDA, ENCCL, ENOTSUP = 0, -1, -2, -3, -4, -5, -6, -7


class BayaderaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class WalkerCountError(BayaderaError, ValueError):
    """IllegalArgumentException "Number of walkers (%d) must be a multiple of %d." (nvidia_gtx.clj:609-610)"""


class AcorTooShortError(BayaderaError, ValueError):
    """IllegalArgumentException "The autocorrelation time is too long…" (nvidia_gtx.clj:275-278)"""


class ModelCompileError(BayaderaError):
    pass


_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_pp = C.POINTER(C.c_void_p)

# every symbol include/bayadera_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "bay_last_error": (C.c_char_p, []),
    "bay_version": (C.c_char_p, []),
    "bay_engine_create": (C.c_int, [C.c_int, C.c_uint64, C.c_int, _pp]),
    "bay_engine_create_current": (C.c_int, [C.c_uint64, C.c_int, _pp]),
    "bay_engine_release": (C.c_int, [_vp]),
    "bay_engine_processing_elements": (C.c_int, [_vp, C.POINTER(_i64)]),
    "bay_engine_stream": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "bay_engine_synchronize": (C.c_int, [_vp]),
    "bay_nccl_unique_id": (C.c_int, [_u8]),
    "bay_engine_comm_init": (C.c_int, [_vp, _u8, C.c_int, C.c_int]),
    "bay_model_compile": (C.c_int, [_vp, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_uint32, _pp]),
    "bay_model_release": (C.c_int, [_vp]),
    "bay_model_compile_check": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_uint32,
                                          C.POINTER(_i64), C.c_char_p, _i64]),
    "bay_model_uses_quadform": (C.c_int, [_vp]),
    "bay_model_kernel_info": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bay_sampler_create": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _pp]),
    "bay_sampler_create_dev": (C.c_int, [_vp, _i32, _i64, C.c_uint64, _i64, _pp]),
    "bay_sampler_release": (C.c_int, [_vp]),
    "bay_init": (C.c_int, [_vp, _i32]),
    "bay_init_position_uniform": (C.c_int, [_vp, _i32, _f32]),
    "bay_init_position_from": (C.c_int, [_vp, _vp]),
    "bay_burn_in": (C.c_int, [_vp, _i64, C.c_float]),
    "bay_anneal": (C.c_int, [_vp, _f32, _i64, C.c_float]),
    "bay_acc_rate": (C.c_int, [_vp, C.c_float, C.POINTER(C.c_double)]),
    "bay_run_sampler": (C.c_int, [_vp, _i64, C.c_float, C.POINTER(C.c_double), _vp, _vp, _vp, C.POINTER(_i64)]),
    "bay_last_means": (C.c_int, [_vp, _f32, _i64]),
    "bay_init_move": (C.c_int, [_vp, C.c_float]),
    "bay_move": (C.c_int, [_vp]),
    "bay_move_bare": (C.c_int, [_vp]),
    "bay_set_temperature": (C.c_int, [_vp, C.c_float]),
    "bay_move_bare_half": (C.c_int, [_vp, C.c_int]),
    "bay_set_a": (C.c_int, [_vp, C.c_float]),
    "bay_accu_blocks": (C.c_int, [_vp, _vp, _vp]),
    "bay_sample": (C.c_int, [_vp, _i64, _vp, C.c_int]),
    "bay_histogram": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "bay_histogram_counts": (C.c_int, [_vp, _u32]),
    "bay_mean": (C.c_int, [_vp, _f32]),
    "bay_variance": (C.c_int, [_vp, _f32]),
    "bay_sd": (C.c_int, [_vp, _f32]),
    "bay_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "bay_get_state": (C.c_int, [_vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64)]),
    "bay_set_state": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i64, _i64]),
    "bay_get_state64": (C.c_int, [_vp, _vp, _vp]),
    "bay_set_state64": (C.c_int, [_vp, _vp, _vp]),
    "bay_dataset_mean": (C.c_int, [_vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _f32]),
    "bay_dataset_variance": (C.c_int, [_vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _f32]),
    "bay_dataset_histogram": (C.c_int, [_vp, _vp, C.c_int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "bay_acor": (C.c_int, [_vp, _f32, _i64, _i64, _vp, _vp, _vp, C.POINTER(_i64)]),
    "bay_model_logfn": (C.c_int, [_vp, _vp, _i64, _f32, _i64, _f32]),
    "bay_model_density": (C.c_int, [_vp, _vp, _i64, _f32, _i64, C.c_int, _f32]),
    "bay_model_evidence": (C.c_int, [_vp, _vp, _i64, _f32, _i64, C.POINTER(C.c_double)]),
    "bay_model_density_dev": (C.c_int, [_vp, C.c_uint64, _i64, C.c_uint64, _i64, C.c_int, C.c_uint64]),
    "bay_model_evidence_dev": (C.c_int, [_vp, C.c_uint64, _i64, C.c_uint64, _i64, C.POINTER(C.c_double)]),
    "bay_direct_sample": (C.c_int, [_vp, C.c_int, _i32, _f32, C.c_int, _i64, _vp, C.c_int]),
    "bay_hdi": (C.c_int, [_vp, C.c_double, _vp, _vp, _vp, C.c_int]),
    "bay_hdi_histogram": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp, C.c_int]),
    "bay_mix": (C.c_int, [_vp, _i64, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "bay_glm_loglik_probe": (C.c_int, [_vp, _f32, _i64, C.c_int, _vp]),
    "bay_launch_count": (_i64, []),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA engine.  Fails loudly if it has not been built (``make`` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} is missing: build the CUDA engine first (run `make` at the repo "
                              "root or __graft_entry__.build()); bayadera_b200 has no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = load().bay_last_error().decode("utf-8", "replace")
    cls = {EINVAL_WALKERS: WalkerCountError, EACOR_TOO_SHORT: AcorTooShortError, ECOMPILE: ModelCompileError}.get(rc, BayaderaError)
    raise cls(rc, msg)


def ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
