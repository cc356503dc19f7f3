"""bayadera_b200 — B200-native (sm_100a) engine for Bayadera's stretch-move ensemble MCMC hot path.

The product is ``libbayadera_b200.so`` (hand-written CUDA + NVRTC-compiled model kernels behind the C ABI of
include/bayadera_b200.h); this package is the thin host-side mirror of the reference's protocol interface.
Importing the package does not need a GPU; any compute call without one raises ``BayaderaError``.
"""
from . import mcmc, models
from ._lib import (AcorTooShortError, BayaderaError, ModelCompileError, WalkerCountError, LIB_PATH)
from .engine import (Autocorrelation, B200AcorEngine, DeviceParams, B200BayaderaFactory, B200DatasetEngine, B200Stretch,
                     B200StretchFactory, Histogram, launch_count, nccl_unique_id)
from .engines import B200DirectSamplerEngine, B200DistributionEngine, B200LikelihoodEngine
from .models import DeviceModel

__all__ = ["mcmc", "models", "DeviceModel", "B200BayaderaFactory", "B200StretchFactory", "B200Stretch",
           "B200DatasetEngine", "B200AcorEngine", "DeviceParams", "B200DistributionEngine", "B200LikelihoodEngine",
           "B200DirectSamplerEngine", "Histogram", "Autocorrelation", "BayaderaError",
           "WalkerCountError", "AcorTooShortError", "ModelCompileError", "launch_count", "nccl_unique_id",
           "LIB_PATH"]
