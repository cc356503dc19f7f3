"""ctypes front-end of the CPU oracle — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  It wraps oracle/bayadera_oracle.c (the C restatement of the
reference's kernels, pinned on the reference's golden vectors by
tests/test_oracle_golden.py) and re-states the host sequencing of ``GTXStretch``
(/root/reference/src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj:282-541)
in ``OracleStretch``, in the reference's own AoS layout.

Model callbacks are built from the SAME C source strings the CUDA engine hands to NVRTC
(bayadera_b200.models), compiled here with g++ behind a 4-line prelude.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libbayadera_oracle.so"
_BUILD = _HERE / "_build"
_CXX = "/usr/bin/g++"

LOGFN_T = C.CFUNCTYPE(C.c_float, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_float))

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    newest = max((_HERE / f).stat().st_mtime for f in ("bayadera_oracle.c", "bayadera_oracle_rng.c"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        vp = C.c_void_p
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_u01.restype = C.c_float
        L.orc_u01.argtypes = [C.c_uint32]
        L.orc_direct_uniform.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, _f32p]
        L.orc_init_walkers.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.orc_logfn.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p]
        L.orc_stretch_coeffs.argtypes = [C.c_float, _f32p]
        L.orc_stretch_half.restype = C.c_uint32
        L.orc_stretch_half.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_int,
                                       vp, vp, vp, vp]
        L.orc_block_tree_sum.restype = C.c_float
        L.orc_block_tree_sum.argtypes = [_f32p, C.c_uint32]
        L.orc_accu_epilogue.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _u8p, _u32p, _f32p]
        L.orc_step_means.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.orc_min_max.argtypes = [C.c_uint32, C.c_uint64, _f32p, C.c_uint64, C.c_uint64, _f32p]
        L.orc_histogram.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, _f32p, _f32p, C.c_uint64, C.c_uint64, _u32p]
        L.orc_uint_to_real.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, _f32p, _u32p, _f32p]
        L.orc_bin_ranks.argtypes = [C.c_uint32, C.c_uint32, _f32p, _f32p]
        L.orc_mean_variance.argtypes = [C.c_uint32, C.c_uint64, _f32p, C.c_uint64, C.c_uint64, _f32p, vp]
        L.orc_acor.restype = C.c_int
        L.orc_acor.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_uint32)]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------------------ models --
_PRELUDE = """#include <math.h>
#include <stdint.h>
#define REAL float
#define REAL2 float2
#define ACCUMULATOR float
struct alignas(16) float4 { float x, y, z, w; };   /* CUDA built-in vector type used by some models */
"""
_model_cache: dict = {}


def compile_model(model, opt: str = "-O2"):
    """g++-compile a DeviceModel's source strings; returns (ctypes fn pointer as void*, keepalive)."""
    text = _PRELUDE + f"#define DIM {model.dimension}\n" + "\n".join(model.source) + (
        '\nextern "C" float orc_model_entry(uint32_t data_len, uint32_t params_len, const float* params,'
        " uint32_t dim, const float* x) {\n"
        f"    return {model.mcmc_logpdf}(data_len, params_len, (const float*)params, dim, (const float*)x);\n}}\n")
    key = hashlib.sha1((text + opt).encode()).hexdigest()[:16]
    if key in _model_cache:
        return _model_cache[key]
    _BUILD.mkdir(exist_ok=True)
    so = _BUILD / f"model_{model.name}_{key}.so"
    if not so.exists():
        cpp = _BUILD / f"model_{model.name}_{key}.cpp"
        cpp.write_text(text)
        tmp = so.with_suffix(f".{os.getpid()}.tmp")
        subprocess.run([_CXX, opt, "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", str(tmp), str(cpp)],
                       check=True, capture_output=True)
        os.replace(tmp, so)
    dll = C.CDLL(str(so))
    fn = C.cast(dll.orc_model_entry, C.c_void_p)
    _model_cache[key] = (fn, dll)
    return _model_cache[key]


# --------------------------------------------------------------------------- primitives --
def philox(ctr, key) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(np.asarray(ctr, dtype=np.uint32), np.asarray(key, dtype=np.uint32), out)
    return out


def direct_uniform(n: int, seed: int, lower: float, upper: float) -> np.ndarray:
    out = np.zeros((n + 3) // 4 * 4, dtype=np.float32)
    lib().orc_direct_uniform(n, seed & 0xFFFFFFFF, lower, upper, out)
    return out[:n]


def stretch_coeffs(a: float) -> np.ndarray:
    out = np.zeros(3, dtype=np.float32)
    lib().orc_stretch_coeffs(a, out)
    return out


def block_tree_sum(v: np.ndarray) -> float:
    w = np.ascontiguousarray(v, dtype=np.float32).copy()
    return float(lib().orc_block_tree_sum(w, w.size))


def min_max(data: np.ndarray, dim: int, n: int, offset: int = 0, ld: Optional[int] = None) -> np.ndarray:
    limits = np.zeros(2 * dim, dtype=np.float32)
    lib().orc_min_max(dim, n, np.ascontiguousarray(data, dtype=np.float32).reshape(-1), offset, ld or dim, limits)
    return limits


def histogram_counts(data: np.ndarray, dim: int, n: int, wgs: int, limits: np.ndarray,
                     counts: Optional[np.ndarray] = None, offset: int = 0, ld: Optional[int] = None) -> np.ndarray:
    if counts is None:
        counts = np.zeros(wgs * dim, dtype=np.uint32)
    lib().orc_histogram(dim, n, wgs, np.ascontiguousarray(limits, dtype=np.float32),
                        np.ascontiguousarray(data, dtype=np.float32).reshape(-1), offset, ld or dim, counts)
    return counts


def uint_to_real(counts: np.ndarray, dim: int, wgs: int, n: int, limits: np.ndarray) -> np.ndarray:
    pdf = np.zeros(wgs * dim, dtype=np.float32)
    lib().orc_uint_to_real(wgs, dim, n, np.ascontiguousarray(limits, dtype=np.float32), counts, pdf)
    return pdf


def bin_ranks(pdf: np.ndarray, dim: int, wgs: int) -> np.ndarray:
    out = np.zeros(wgs * dim, dtype=np.float32)
    lib().orc_bin_ranks(wgs, dim, np.ascontiguousarray(pdf, dtype=np.float32), out)
    return out


def mean_variance(data: np.ndarray, dim: int, n: int, offset: int = 0, ld: Optional[int] = None):
    mean = np.zeros(dim, dtype=np.float32)
    var = np.zeros(dim, dtype=np.float32)
    lib().orc_mean_variance(dim, n, np.ascontiguousarray(data, dtype=np.float32).reshape(-1), offset, ld or dim,
                            mean, var.ctypes.data_as(C.c_void_p))
    return mean, var


def acor(series: np.ndarray, dim: int, n: int, wgs: int):
    """series: dim x n column-major (flat index dim*t + d).  Returns (tau, mean, sigma, lag) or raises."""
    work = np.ascontiguousarray(series, dtype=np.float32).reshape(-1).copy()
    tau = np.zeros(dim, dtype=np.float32)
    mean = np.zeros(dim, dtype=np.float32)
    sigma = np.zeros(dim, dtype=np.float32)
    lag = C.c_uint32(0)
    rc = lib().orc_acor(dim, n, wgs, work, tau, mean, sigma, C.byref(lag))
    if rc != 0:
        raise ValueError("The autocorrelation time is too long relative to the variance. "
                         f"Number of steps ({n}) must not be less than {lag.value * 5}.")
    return tau, mean, sigma, lag.value


# ------------------------------------------------------------------------ OracleStretch --
class OracleStretch:
    """Host sequencing of GTXStretch (nvidia_gtx.clj:282-541) over the C oracle kernels.

    xs is the reference's AoS buffer: walker w at xs[w*D:(w+1)*D]; s0 = walkers [0,H), s1 = [H,W).
    Deviations from the CUDA twin of the reference follow SURVEY Appendix B (zero-filled accept,
    log-densities initialised for both halves, walker-aligned partner unless literal_partner)."""

    def __init__(self, model, seed: int, walkers: int, params: np.ndarray, wgs: int = 256,
                 literal_partner: bool = False):
        if walkers < 2 * wgs or walkers % (2 * wgs) != 0:
            raise ValueError(f"Number of walkers ({walkers}) must be a multiple of {2 * wgs}.")
        self.model, self.W, self.H, self.D, self.wgs = model, walkers, walkers // 2, model.dimension, wgs
        self.G = (self.H + wgs - 1) // wgs
        self.fn, self._keep = compile_model(model)
        self.params = np.ascontiguousarray(params, dtype=np.float32).reshape(-1)
        if self.params.size == 0:
            self.params = np.zeros(1, dtype=np.float32)
            self._params_count = 0
        else:
            self._params_count = self.params.size
        self.params_len = model.params_size
        self.data_len = max(0, self._params_count - model.params_size)
        self.xs = np.zeros(self.W * self.D, dtype=np.float32)
        self.lp = np.zeros(self.W, dtype=np.float32)
        self.accept = np.zeros(self.G, dtype=np.uint32)
        self.blk_sums = np.zeros(self.D * self.G, dtype=np.float32)
        self.means: list = []
        self.literal = int(literal_partner)
        self.iterations = 0
        self.diag = None  # per-walker diagnostics of the last half-step when requested
        self.init(seed)

    # -- MCMC protocol ------------------------------------------------------------------
    def init(self, seed: int):
        self.bare_seed, self.a_bare, self.beta, self.bare_counter, self.move_seed = seed, 2.0, 1.0, 0, seed
        return self

    def _logfn_all(self):
        lib().orc_logfn(self.fn, self.W, self.D, self.data_len, self.params_len, self.params, self.xs, self.lp)

    def init_position(self, seed: int, limits: np.ndarray):
        lim = np.ascontiguousarray(limits, dtype=np.float32).reshape(-1)
        lib().orc_init_walkers(self.W * self.D // 4, self.D, seed & 0xFFFFFFFF, lim, self.xs)
        self._logfn_all()
        self.iterations = 0
        return self

    def set_positions(self, xs: np.ndarray, lp: Optional[np.ndarray] = None):
        self.xs[:] = np.asarray(xs, dtype=np.float32).reshape(-1)
        if lp is None:
            self._logfn_all()
        else:
            self.lp[:] = lp
        return self

    def _half(self, half: int, seed: int, tag: int, step: int, a: float, beta: float, want_diag: bool = False):
        H, D = self.H, self.D
        act = self.xs[half * H * D:(half + 1) * H * D]
        cmp_ = self.xs[(1 - half) * H * D:(2 - half) * H * D]
        lp = self.lp[half * H:(half + 1) * H]
        acc = np.zeros(H, dtype=np.uint8)
        ptr = lambda arr: arr.ctypes.data_as(C.c_void_p)
        ly = q = uz = None
        if want_diag:
            ly, q, uz = (np.zeros(H, dtype=np.float32) for _ in range(3))
        n = lib().orc_stretch_half(self.fn, H, D, seed & 0xFFFFFFFF, tag, step & 0xFFFFFFFF, self.data_len,
                                   self.params_len, self.params, cmp_, act, lp, a, beta, self.literal,
                                   ptr(acc), ptr(ly) if want_diag else None, ptr(q) if want_diag else None,
                                   ptr(uz) if want_diag else None)
        if want_diag:
            self.diag = dict(acc=acc, ly=ly, q=q, uz=uz)
        return n, acc

    def half_bare(self, half: int, want_diag: bool = False):
        """One half of move-bare! (odd: half 0, tag 3333, seed; even: half 1, tag 4444, seed+1)."""
        seed = self.bare_seed + half
        return self._half(half, seed, 4444 if half else 3333, self.bare_counter, self.a_bare, self.beta, want_diag)

    def move_bare(self):
        self.half_bare(0)
        self.half_bare(1)
        self.bare_counter += 1
        return self

    def set_temperature(self, t: float):
        self.beta = float(np.float32(1.0 / t))
        return self

    def burn_in(self, n: int, a: float = 2.0):
        self.a_bare, self.beta = a, 1.0
        for _ in range(n):
            self.move_bare()
        self.iterations += n
        return self

    def anneal(self, schedule, n: int, a: float = 2.0):
        self.a_bare = a
        for i in range(n):
            self.set_temperature(schedule(i))
            self.move_bare()
        self.iterations += n
        return self

    def init_move(self, a: float):
        self.move_seed += 2
        self.move_counter = 0
        self.a_move = a
        self.accept[:] = 0
        self.blk_sums[:] = 0
        self.means = []
        return self

    def move(self):
        H, D = self.H, self.D
        self.blk_sums[:] = 0
        for half, tag in ((0, 1111), (1, 2222)):
            _, acc = self._half(half, self.move_seed + half, tag, self.move_counter, self.a_move, 1.0)
            act = self.xs[half * H * D:(half + 1) * H * D]
            lib().orc_accu_epilogue(H, D, self.wgs, act, acc, self.accept, self.blk_sums)
        m = np.zeros(D, dtype=np.float32)
        lib().orc_step_means(self.G, D, self.wgs, self.blk_sums, m)
        self.means.append(m)
        self.move_counter += 1
        return self

    def acc_rate(self, a: float = 2.0) -> float:
        self.init_move(a)
        self.move()
        self.iterations += 1
        return float(self.accept.astype(np.uint64).sum()) / self.W

    def run_sampler(self, n: int, a: float = 2.0) -> dict:
        self.init_move(a)
        for _ in range(n):
            self.move()
        self.iterations += n
        means = np.stack(self.means, axis=0)  # n x D  == D x n column-major
        tau, mean, sigma, lag = acor(means, self.D, n, self.wgs)
        return {"acceptance-rate": float(self.accept.astype(np.uint64).sum()) / (self.W * n), "a": a,
                "autocorrelation": {"tau": tau, "mean": mean, "sigma": sigma, "steps": n, "lag": lag},
                "means": means}

    def sample(self, n: Optional[int] = None) -> np.ndarray:
        n = self.W if n is None else n
        out = np.zeros(n * self.D, dtype=np.float32)
        self.set_temperature(1.0)
        done = 0
        while done < n:
            self.move_bare()
            self.iterations += 1
            take = min(n - done, self.W)
            out[done * self.D:(done + take) * self.D] = self.xs[:take * self.D]
            done += take
        return out.reshape(n, self.D)  # row = walker (== DIM x n column-major)

    # -- estimate engine ----------------------------------------------------------------
    def histogram(self, cycles: int = 1) -> dict:
        D, W, wgs = self.D, self.W, self.wgs
        limits = min_max(self.xs, D, W)
        counts = histogram_counts(self.xs, D, W, wgs, limits)
        self.set_temperature(1.0)
        for _ in range(cycles - 1):
            self.move_bare()
            histogram_counts(self.xs, D, W, wgs, limits, counts)
        self.iterations += cycles - 1
        pdf = uint_to_real(counts, D, wgs, cycles * W, limits)
        return {"limits": limits.reshape(D, 2), "pdf": pdf.reshape(D, wgs), "bin-ranks": bin_ranks(pdf, D, wgs).reshape(D, wgs),
                "counts": counts.reshape(D, wgs)}

    def mean(self) -> np.ndarray:
        return mean_variance(self.xs, self.D, self.W)[0]

    def variance(self) -> np.ndarray:
        return mean_variance(self.xs, self.D, self.W)[1]

    def sd(self) -> np.ndarray:
        return np.sqrt(self.variance())
