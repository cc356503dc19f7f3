"""Host mirror of the engines on either side of the sampler:

    GTXDistributionEngine  -> B200DistributionEngine   (DensityEngine: log-density, density)
    GTXLikelihoodEngine    -> B200LikelihoodEngine     (DensityEngine + LikelihoodEngine.evidence)
    GTXDirectSamplerEngine -> B200DirectSamplerEngine  (RandomSamplerEngine.sample [seed params res])

(/root/reference/src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj:48-141).  Thin calls into the C ABI.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from ._lib import check, ptr
from .engine import B200BayaderaFactory, B200StretchFactory, _f32
from .models import DeviceModel, distribution_source

FAMILIES = {"uniform": 0, "gaussian": 1, "exponential": 2, "erlang": 3}


def likelihood_engine_model(lik: DeviceModel) -> DeviceModel:
    """Adapts a likelihood model's ``loglik(data_len, data, dim, x)`` to the LOGFN signature the engine compiles:
    the whole params vector is the data (params-size 0), as in GTXLikelihoodEngine (nvidia_gtx.clj:103-141)."""
    if lik.loglik is None:
        raise ValueError(f"model {lik.name} has no loglik")
    name = f"{lik.name}_loglik_engine"
    wrapper = distribution_source(name, f"return {lik.loglik}(data_len + params_len, params, dim, x);")
    return dataclasses.replace(lik, name=name, source=lik.source + (wrapper,), mcmc_logpdf=name, logpdf=name,
                               params_size=0)


class B200DistributionEngine:
    """``gtx-distribution-engine``: log-density / density with the model's normalised ``logpdf``."""

    def __init__(self, factory: B200BayaderaFactory, model: DeviceModel, logfn: str | None = None):
        fn = logfn or model.logpdf or model.mcmc_logpdf
        self.model = model
        self._sf = B200StretchFactory(factory, dataclasses.replace(model, mcmc_logpdf=fn))
        self._L = factory._L

    def _density(self, params, x, exponentiate: int) -> np.ndarray:
        pts = _f32(x).reshape(-1, self.model.dimension)
        p = _f32(params).reshape(-1)
        out = np.zeros(pts.shape[0], dtype=np.float32)
        check(self._L.bay_model_density(self._sf._h, ptr(p), p.size, pts.reshape(-1), pts.shape[0], exponentiate, out))
        return out

    def log_density(self, params, x) -> np.ndarray:
        return self._density(params, x, 0)

    def density(self, params, x) -> np.ndarray:
        return self._density(params, x, 1)

    def density_dev(self, params_ptr: int, params_count: int, x_ptr: int, n: int, out_ptr: int,
                    exponentiate: bool = False) -> None:
        """The reference's calling convention: params, the DIM x n point matrix and the result are blocks of the
        engine's CUDA context (nvidia_gtx.clj:83-104); nothing crosses to the host."""
        check(self._L.bay_model_density_dev(self._sf._h, params_ptr, params_count, x_ptr, n, int(exponentiate), out_ptr))

    def release(self) -> None:
        self._sf.release()


class B200LikelihoodEngine(B200DistributionEngine):
    """``gtx-likelihood-engine``: log-density = loglik, density = lik, plus ``evidence``."""

    def __init__(self, factory: B200BayaderaFactory, lik: DeviceModel):
        super().__init__(factory, likelihood_engine_model(lik))

    def evidence(self, data, x) -> float:
        pts = _f32(x).reshape(-1, self.model.dimension)
        p = _f32(data).reshape(-1)
        out = C.c_double()
        check(self._L.bay_model_evidence(self._sf._h, ptr(p), p.size, pts.reshape(-1), pts.shape[0], C.byref(out)))
        return out.value

    def evidence_dev(self, data_ptr: int, data_count: int, x_ptr: int, n: int) -> float:
        out = C.c_double()
        check(self._L.bay_model_evidence_dev(self._sf._h, data_ptr, data_count, x_ptr, n, C.byref(out)))
        return out.value


class B200DirectSamplerEngine:
    """``gtx-direct-sampler-engine``: i.i.d. draws from uniform / gaussian / exponential / erlang."""

    def __init__(self, factory: B200BayaderaFactory, family: str):
        if family not in FAMILIES:
            raise ValueError(f"no direct sampler for {family}")
        self._L, self.factory, self.family = factory._L, factory, family

    def sample(self, seed: int, params, n: int) -> np.ndarray:
        p = _f32(params).reshape(-1)
        out = np.zeros(n, dtype=np.float32)
        check(self._L.bay_direct_sample(self.factory._h, FAMILIES[self.family], seed, p, p.size, n, ptr(out), 0))
        return out
