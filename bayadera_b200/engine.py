"""Host-side mirror of the reference's device-engine interface, over the C ABI.

Names and argument meaning follow uncomplicate.bayadera.internal.protocols
(/root/reference/src/clojure/uncomplicate/bayadera/internal/protocols.clj) and the host types of
/root/reference/src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj:

    GTXBayaderaFactory  -> B200BayaderaFactory   (EngineFactory: mcmc-factory, dataset-engine, processing-elements)
    GTXStretchFactory   -> B200StretchFactory    (SamplerFactory: create-sampler [seed walkers params])
    GTXStretch          -> B200Stretch           (MCMC, MCMCStretch, RandomSampler, EstimateEngine, Location, Spread)
    GTXDatasetEngine    -> B200DatasetEngine     (DatasetEngine, EstimateEngine on a matrix)
    GTXAcorEngine       -> B200AcorEngine        (AcorEngine)

Clojure's ``foo-bar!`` becomes ``foo_bar``.  Every method is a thin call into
libbayadera_b200.so; nothing is computed in Python and there is no CPU path.
Matrices are numpy float32 in the reference's layout: a ``DIM x n`` column-major matrix is a
C-contiguous ``(n, DIM)`` array (row = one walker / sample).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check, ptr
from .models import DeviceModel


@dataclass
class Histogram:
    """protocols.clj:13-18 — limits 2 x DIM, pdf WGS x DIM, bin-ranks WGS x DIM (stored row = dimension)."""
    limits: np.ndarray
    pdf: np.ndarray
    bin_ranks: np.ndarray


@dataclass
class Autocorrelation:
    """protocols.clj:111-116"""
    tau: np.ndarray
    mean: np.ndarray
    sigma: np.ndarray
    steps: int
    lag: int


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


@dataclass
class DeviceParams:
    """A params vector that already lives in device memory: raw pointer, element count and an object that
    keeps the allocation alive (e.g. the torch tensor it came from)."""
    ptr: int
    count: int
    owner: object = None

    @staticmethod
    def from_torch(t) -> "DeviceParams":
        import torch
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        return DeviceParams(int(t.data_ptr()), int(t.numel()), t)


class B200BayaderaFactory:
    """``gtx-bayadera-factory [ctx hstream compute-units WGS]`` (nvidia_gtx.clj:791-807).

    ``stream`` is a raw CUDA stream handle (e.g. ``torch.cuda.current_stream().cuda_stream``);
    0 lets the engine create its own.  The reference's default WGS is max-block-dim-x = 1024."""

    def __init__(self, device: int = 0, stream: int = 0, wgs: int = 1024, current_context: bool = False):
        self._L = _lib.load()
        h = C.c_void_p()
        if current_context:
            # adopt the CUDA context current on this thread (ClojureCUDA's with-default context, cuda.clj:27-31)
            check(self._L.bay_engine_create_current(stream, wgs, C.byref(h)))
        else:
            check(self._L.bay_engine_create(device, stream, wgs, C.byref(h)))
        self._h, self.device, self.wgs = h, device, wgs
        self._dataset_engine = B200DatasetEngine(self)
        self._acor_engine = B200AcorEngine(self)

    # EngineFactory, protocols.clj:132-138
    def mcmc_factory(self, model: DeviceModel) -> "B200StretchFactory":
        return B200StretchFactory(self, model)

    def dataset_engine(self) -> "B200DatasetEngine":
        return self._dataset_engine

    def acor_engine(self) -> "B200AcorEngine":
        return self._acor_engine

    def processing_elements(self) -> int:
        out = C.c_int64()
        check(self._L.bay_engine_processing_elements(self._h, C.byref(out)))
        return out.value

    def stream(self) -> int:
        out = C.c_uint64()
        check(self._L.bay_engine_stream(self._h, C.byref(out)))
        return out.value

    def synchronize(self) -> None:
        check(self._L.bay_engine_synchronize(self._h))

    def comm_init(self, unique_id: np.ndarray, nranks: int, rank: int) -> None:
        check(self._L.bay_engine_comm_init(self._h, np.ascontiguousarray(unique_id, dtype=np.uint8), nranks, rank))

    def release(self) -> None:
        if self._h:
            self._L.bay_engine_release(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()


def nccl_unique_id() -> np.ndarray:
    out = np.zeros(128, dtype=np.uint8)
    check(_lib.load().bay_nccl_unique_id(out))
    return out


class B200StretchFactory:
    """``gtx-stretch-factory`` + ``GTXStretchFactory`` (nvidia_gtx.clj:543-610, 747-757): NVRTC-compiles the
    model's C sources together with the stretch kernels for sm_100a."""

    def __init__(self, factory: B200BayaderaFactory, model: DeviceModel):
        self._L, self.factory, self.model = factory._L, factory, model
        srcs = (C.c_char_p * len(model.source))(*[s.encode() for s in model.source])
        h = C.c_void_p()
        check(self._L.bay_model_compile(factory._h, srcs, len(model.source), model.mcmc_logpdf.encode(),
                                        model.dimension, model.params_size, model.flags, C.byref(h)))
        self._h = h

    # SamplerFactory, protocols.clj:120-121
    def create_sampler(self, seed: int, walkers: int, params) -> "B200Stretch":
        return B200Stretch(self, seed, walkers, params)

    def uses_quadform(self) -> bool:
        """True when moves of this model run on the tensor-core quadratic-form kernel (BAY_MODEL_QUADFORM)."""
        return bool(self._L.bay_model_uses_quadform(self._h))

    def kernel_info(self, kernel: str = "bay_stretch_bare") -> dict:
        r, l, s = C.c_int(), C.c_int(), C.c_int()
        check(self._L.bay_model_kernel_info(self._h, kernel.encode(), C.byref(r), C.byref(l), C.byref(s)))
        return {"registers": r.value, "local_bytes": l.value, "shared_bytes": s.value}

    def log_density(self, params, x) -> np.ndarray:
        """DensityEngine.log-density over a DIM x n point matrix (nvidia-gtx-distribution.cu:5-13),
        evaluated with the model's mcmc-logpdf."""
        x = _f32(x).reshape(-1, self.model.dimension)
        p = _f32(params).reshape(-1)
        out = np.zeros(x.shape[0], dtype=np.float32)
        check(self._L.bay_model_logfn(self._h, ptr(p), p.size, x.reshape(-1), x.shape[0], out))
        return out

    def release(self) -> None:
        if self._h:
            self._L.bay_model_release(self._h)
            self._h = None


class B200Stretch:
    """``GTXStretch`` (nvidia_gtx.clj:282-541)."""

    def __init__(self, sfactory: B200StretchFactory, seed: int, walkers: int, params):
        self._L, self.sfactory = sfactory._L, sfactory
        self.model = sfactory.model
        self.DIM, self.WGS = self.model.dimension, sfactory.factory.wgs
        h = C.c_void_p()
        if isinstance(params, DeviceParams):
            # params already on the engine's device (the reference borrows a cuda-float vector, nvidia_gtx.clj:558)
            self._params_keepalive = params
            check(self._L.bay_sampler_create_dev(sfactory._h, seed, walkers, params.ptr, params.count, C.byref(h)))
        else:
            p = _f32(params).reshape(-1)
            check(self._L.bay_sampler_create(sfactory._h, seed, walkers, ptr(p), p.size, C.byref(h)))
        self._h, self.walker_count = h, walkers

    # Info (nvidia_gtx.clj:332-335)
    def info(self) -> dict:
        w, it = C.c_int64(), C.c_int64()
        check(self._L.bay_info(self._h, C.byref(w), C.byref(it)))
        return {"walker-count": w.value, "iteration-counter": it.value}

    # MCMC protocol (protocols.clj:97-103)
    def init(self, seed: int) -> "B200Stretch":
        check(self._L.bay_init(self._h, seed))
        return self

    def init_position(self, seed_or_position, limits=None) -> "B200Stretch":
        if isinstance(seed_or_position, B200Stretch):
            check(self._L.bay_init_position_from(self._h, seed_or_position._h))
        else:
            lim = _f32(self.model.limits_array() if limits is None else limits).reshape(-1)
            if lim.size != 2 * self.DIM:
                raise ValueError(f"limits must be 2 x {self.DIM}")
            check(self._L.bay_init_position_uniform(self._h, seed_or_position, lim))
        return self

    def burn_in(self, n: int, a: float = 2.0) -> "B200Stretch":
        check(self._L.bay_burn_in(self._h, n, a))
        return self

    def anneal(self, schedule: Callable[[int], float], n: int, a: float = 2.0) -> "B200Stretch":
        temps = _f32([schedule(i) for i in range(n)]) if n > 0 else np.zeros(1, dtype=np.float32)
        check(self._L.bay_anneal(self._h, temps, n, a))
        return self

    def acc_rate(self, a: float = 2.0) -> float:
        out = C.c_double()
        check(self._L.bay_acc_rate(self._h, a, C.byref(out)))
        return out.value

    def run_sampler(self, n: int, a: float = 2.0) -> dict:
        acc, lag = C.c_double(), C.c_int64()
        tau, mean, sigma = (np.zeros(self.DIM, dtype=np.float32) for _ in range(3))
        check(self._L.bay_run_sampler(self._h, n, a, C.byref(acc), ptr(tau), ptr(mean), ptr(sigma), C.byref(lag)))
        return {"acceptance-rate": acc.value, "a": a,
                "autocorrelation": Autocorrelation(tau, mean, sigma, n, lag.value)}

    def last_means(self, n: int) -> np.ndarray:
        out = np.zeros((n, self.DIM), dtype=np.float32)
        check(self._L.bay_last_means(self._h, out.reshape(-1), n))
        return out

    # MCMCStretch protocol (protocols.clj:105-109)
    def init_move(self, a: float = 2.0) -> "B200Stretch":
        check(self._L.bay_init_move(self._h, a))
        return self

    def move(self) -> "B200Stretch":
        check(self._L.bay_move(self._h))
        return self

    def move_bare(self) -> "B200Stretch":
        check(self._L.bay_move_bare(self._h))
        return self

    def set_temperature(self, t: float) -> "B200Stretch":
        check(self._L.bay_set_temperature(self._h, t))
        return self

    def move_bare_half(self, half: int) -> "B200Stretch":
        """One raw stretch_move_bare launch (nvidia_gtx_test.clj:217-235); the step counter is not advanced."""
        check(self._L.bay_move_bare_half(self._h, half))
        return self

    def set_a(self, a: float) -> "B200Stretch":
        check(self._L.bay_set_a(self._h, a))
        return self

    def accu_blocks(self):
        g = (self.walker_count // 2 + self.WGS - 1) // self.WGS
        accept = np.zeros(g, dtype=np.uint32)
        sums = np.zeros((self.DIM, g), dtype=np.float32)
        check(self._L.bay_accu_blocks(self._h, ptr(accept), ptr(sums)))
        return accept, sums

    # RandomSampler (protocols.clj:94-95)
    def sample(self, n: Optional[int] = None) -> np.ndarray:
        n = self.walker_count if n is None else n
        out = np.zeros((n, self.DIM), dtype=np.float32)
        check(self._L.bay_sample(self._h, n, ptr(out), 0))
        return out

    def sample_into_device(self, n: int, device_ptr: int) -> None:
        check(self._L.bay_sample(self._h, n, C.c_void_p(device_ptr), 1))

    # EstimateEngine / Location / Spread (protocols.clj:20-28, 82-84)
    def histogram(self, cycles: int = 1) -> Histogram:
        lim = np.zeros((self.DIM, 2), dtype=np.float32)
        pdf = np.zeros((self.DIM, self.WGS), dtype=np.float32)
        ranks = np.zeros((self.DIM, self.WGS), dtype=np.float32)
        check(self._L.bay_histogram(self._h, cycles, ptr(lim), ptr(pdf), ptr(ranks)))
        return Histogram(lim, pdf, ranks)

    def histogram_counts(self) -> np.ndarray:
        out = np.zeros(self.DIM * self.WGS, dtype=np.uint32)
        check(self._L.bay_histogram_counts(self._h, out))
        return out.reshape(self.DIM, self.WGS)

    def hdi(self, mass: float = 0.95, max_regions: Optional[int] = None) -> list:
        """``hdi`` (util.clj:102-110) of every dimension of the LATEST ``histogram()``, computed on the device
        (SURVEY §8f row 4).  Returns one ``(count, regions)`` per dimension, regions a (k, 2) array of [lo, hi]."""
        max_regions = (self.WGS + 1) // 2 if max_regions is None else max_regions   # runs of bins: at most half
        counts = np.zeros(self.DIM, dtype=np.int32)
        nreg = np.zeros(self.DIM, dtype=np.int32)
        regions = np.zeros((self.DIM, max_regions, 2), dtype=np.float32)
        check(self._L.bay_hdi(self._h, float(mass), ptr(counts), ptr(nreg), ptr(regions), max_regions))
        return [(int(counts[d]), regions[d, :min(int(nreg[d]), max_regions)].copy()) for d in range(self.DIM)]

    def mix(self, options: Optional[dict] = None) -> dict:
        """``mix!`` (mcmc.clj:66-101) in one boundary crossing (``bay_mix``); same options and result map as
        ``bayadera_b200.mcmc.mix``, for the three built-in cooling schedules."""
        from . import mcmc
        o = dict(options or {})
        sched = o.get("cooling-schedule", mcmc.minus_n)
        power = 1.0
        if sched is mcmc.minus_n:
            kind = 0
        elif sched is mcmc.sqrt_n:
            kind = 1
        elif getattr(sched, "pow_n_power", None) is not None:
            kind, power = 2, float(sched.pow_n_power)
        else:
            return mcmc.mix(self, o)          # arbitrary host schedule: the protocol-call loop
        a, r1, r2 = C.c_double(), C.c_double(), C.c_double()
        check(self._L.bay_mix(self._h, int(o.get("step", 64)), float(o.get("dimension-power", 0.8)), kind, power,
                              float(o.get("a", 2.0)), float(o.get("min-acc-rate", 0.2)),
                              float(o.get("max-acc-rate", 0.5)), C.byref(a), C.byref(r1), C.byref(r2)))
        return {"a": a.value, "acc-rate": r1.value, "acc-rate-2.0": r2.value}

    def mean(self) -> np.ndarray:
        out = np.zeros(self.DIM, dtype=np.float32)
        check(self._L.bay_mean(self._h, out))
        return out

    def variance(self) -> np.ndarray:
        out = np.zeros(self.DIM, dtype=np.float32)
        check(self._L.bay_variance(self._h, out))
        return out

    def sd(self) -> np.ndarray:
        out = np.zeros(self.DIM, dtype=np.float32)
        check(self._L.bay_sd(self._h, out))
        return out

    # state hand-off
    def get_state(self) -> dict:
        xs = np.zeros((self.walker_count, self.DIM), dtype=np.float32)
        lp = np.zeros(self.walker_count, dtype=np.float32)
        bs, ms, bc, mc = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        check(self._L.bay_get_state(self._h, ptr(xs), ptr(lp), C.byref(bs), C.byref(ms), C.byref(bc), C.byref(mc)))
        return {"xs": xs, "logfn": lp, "bare_seed": bs.value, "move_seed": ms.value,
                "bare_counter": bc.value, "move_counter": mc.value}

    def set_state(self, xs=None, logfn=None, bare_seed=None, move_seed=None, bare_counter=None,
                  move_counter=None) -> "B200Stretch":
        cur = None
        if None in (bare_seed, move_seed, bare_counter, move_counter):
            bs, ms, bc, mc = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
            check(self._L.bay_get_state(self._h, None, None, C.byref(bs), C.byref(ms), C.byref(bc), C.byref(mc)))
            cur = (bs.value, ms.value, bc.value, mc.value)
        pick = lambda v, i: cur[i] if v is None else v
        x = None if xs is None else _f32(xs).reshape(-1)
        l = None if logfn is None else _f32(logfn).reshape(-1)
        check(self._L.bay_set_state(self._h, ptr(x), ptr(l), pick(bare_seed, 0), pick(move_seed, 1),
                                    pick(bare_counter, 2), pick(move_counter, 3)))
        return self

    def get_state64(self, xs_out: Optional[np.ndarray] = None, logfn_out: Optional[np.ndarray] = None):
        """Positions (W x DIM fp32) and log-densities in double; optional preallocated (e.g. pinned) outputs."""
        xs = np.zeros((self.walker_count, self.DIM), dtype=np.float32) if xs_out is None else xs_out
        lp = np.zeros(self.walker_count, dtype=np.float64) if logfn_out is None else logfn_out
        check(self._L.bay_get_state64(self._h, ptr(xs), ptr(lp)))
        return xs, lp

    def set_state64(self, xs, logfn64) -> "B200Stretch":
        x = _f32(xs).reshape(-1)
        l = np.ascontiguousarray(logfn64, dtype=np.float64).reshape(-1)
        check(self._L.bay_set_state64(self._h, ptr(x), ptr(l)))
        return self

    def uses_quadform(self) -> bool:
        return self.sfactory.uses_quadform()

    def glm_loglik_probe(self, points, method: int = 0) -> np.ndarray:
        """Row-additive GLM samplers: sum over the dataset of A(x_row . point) for each of the given points
        ((n, DIM) array), by method 0 = the sampler's own path (tensor cores when eligible), 1 = fp32 SIMT, 2 = fp64."""
        pts = _f32(points).reshape(-1, self.DIM)
        out = np.zeros(pts.shape[0], dtype=np.float64)
        check(self._L.bay_glm_loglik_probe(self._h, pts.reshape(-1), pts.shape[0], method, ptr(out)))
        return out

    def release(self) -> None:
        if self._h:
            self._L.bay_sampler_release(self._h)
            self._h = None


class B200DatasetEngine:
    """``GTXDatasetEngine`` (nvidia_gtx.clj:145-228): data-mean, data-variance, histogram of an m x n matrix
    given as a C-contiguous (n, m) numpy array (column-major m x n)."""

    def __init__(self, factory: B200BayaderaFactory):
        self._L, self.factory = factory._L, factory

    def _args(self, data):
        d = _f32(data)
        if d.ndim != 2:
            raise ValueError("data must be 2-D (n samples x m dimensions)")
        n, m = d.shape
        return d, m, n

    def data_mean(self, data) -> np.ndarray:
        d, m, n = self._args(data)
        out = np.zeros(m, dtype=np.float32)
        check(self._L.bay_dataset_mean(self.factory._h, ptr(d), 0, m, n, 0, m, out))
        return out

    def data_variance(self, data) -> np.ndarray:
        d, m, n = self._args(data)
        out = np.zeros(m, dtype=np.float32)
        check(self._L.bay_dataset_variance(self.factory._h, ptr(d), 0, m, n, 0, m, out))
        return out

    def histogram(self, data, with_counts: bool = False):
        d, m, n = self._args(data)
        wgs = self.factory.wgs
        lim = np.zeros((m, 2), dtype=np.float32)
        pdf = np.zeros((m, wgs), dtype=np.float32)
        ranks = np.zeros((m, wgs), dtype=np.float32)
        counts = np.zeros((m, wgs), dtype=np.uint32) if with_counts else None
        check(self._L.bay_dataset_histogram(self.factory._h, ptr(d), 0, m, n, 0, m, ptr(lim), ptr(pdf), ptr(ranks),
                                            ptr(counts)))
        h = Histogram(lim, pdf, ranks)
        return (h, counts) if with_counts else h


class B200AcorEngine:
    """``GTXAcorEngine`` (nvidia_gtx.clj:230-278); series is (n steps, dim) C-contiguous."""

    def __init__(self, factory: B200BayaderaFactory):
        self._L, self.factory = factory._L, factory

    def acor(self, series) -> Autocorrelation:
        s = _f32(series)
        n, dim = s.shape
        tau, mean, sigma = (np.zeros(dim, dtype=np.float32) for _ in range(3))
        lag = C.c_int64()
        check(self._L.bay_acor(self.factory._h, s.reshape(-1), dim, n, ptr(tau), ptr(mean), ptr(sigma), C.byref(lag)))
        return Autocorrelation(tau, mean, sigma, n, lag.value)


def hdi_histogram(factory: B200BayaderaFactory, histogram: "Histogram", mass: float = 0.95,
                  forced_counts=None, max_regions: Optional[int] = None) -> list:
    """hdi-rank-count + hdi-regions (util.clj:52-100) of every column of a caller-held ``Histogram`` on the device."""
    lim, pdf, ranks = _f32(histogram.limits), _f32(histogram.pdf), _f32(histogram.bin_ranks)
    dim, bins = pdf.shape
    max_regions = (bins + 1) // 2 if max_regions is None else max_regions
    counts = np.zeros(dim, dtype=np.int32)
    nreg = np.zeros(dim, dtype=np.int32)
    regions = np.zeros((dim, max_regions, 2), dtype=np.float32)
    forced = None if forced_counts is None else np.ascontiguousarray(forced_counts, dtype=np.int32)
    check(factory._L.bay_hdi_histogram(factory._h, bins, dim, ptr(lim), ptr(pdf), ptr(ranks), float(mass), ptr(forced),
                                       ptr(counts), ptr(nreg), ptr(regions), max_regions))
    return [(int(counts[d]), regions[d, :min(int(nreg[d]), max_regions)].copy()) for d in range(dim)]


def launch_count() -> int:
    return int(_lib.load().bay_launch_count())
