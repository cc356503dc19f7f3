"""ctypes front-end of the direct-sampler part of the CPU oracle (oracle/bayadera_oracle_rng.c) — TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

from . import oracle as _o

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def _lib():
    L = _o.lib()
    if not getattr(L, "_rng_ready", False):
        L.orc_direct_gaussian.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32p]
        L.orc_direct_exponential.argtypes = [C.c_uint32, C.c_uint32, C.c_float, _f32p]
        L.orc_direct_erlang.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32p]
        L._rng_ready = True
    return L


def _out(n):
    return np.zeros((n + 3) // 4 * 4, dtype=np.float32)


def direct_sample(family: str, n: int, seed: int, params) -> np.ndarray:
    """family in uniform | gaussian | exponential | erlang; params as the reference packs them
    (uniform [a b], gaussian [mu sigma], exponential [lambda], erlang [lambda k])."""
    p = [float(np.float32(v)) for v in params]
    seed &= 0xFFFFFFFF
    if family == "uniform":
        return _o.direct_uniform(n, seed, p[0], p[1])
    x = _out(n)
    if family == "gaussian":
        _lib().orc_direct_gaussian(n, seed, p[0], p[1], x)
    elif family == "exponential":
        _lib().orc_direct_exponential(n, seed, p[0], x)
    elif family == "erlang":
        _lib().orc_direct_erlang(n, seed, p[0], p[1], x)
    else:
        raise ValueError(family)
    return x[:n]
