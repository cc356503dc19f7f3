"""The reference's OWN kernel text, executed on the host (TEST INFRASTRUCTURE ONLY — see ref_shim/cl_shim.h).

`make -C oracle ref` (run by __graft_entry__.build() when /root/reference is present) compiles
/root/reference/src/device/uncomplicate/bayadera/internal/device/opencl/engines/amd-gcn-mcmc-stretch.cl together with a
distribution file into oracle/_ref/libref_<model>_d<DIM>.so; the .so travels to the GPU box, the sources do not.
`ReferenceTextStretch` drives those kernels with the host sequencing of the oracle (= GTXStretch / GCNStretch), so
that the C restatement in bayadera_oracle.c can be held against the real kernels step by step."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .oracle import OracleStretch

REF_DIR = Path(__file__).resolve().parent / "_ref"
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

# (library stem) -> what it was built from; keep in step with oracle/Makefile
BUILT = {"uniform_d1": ("uniform.cl", "uniform_logpdf", 1),
         "gaussian_d1": ("gaussian.cl", "gaussian_mcmc_logpdf", 1),
         "beta_binomial_d1": ("beta.cl + binomial.cl + the expanded posterior template", "beta_binomial_mcmc_logpdf", 1),
         "gaussian2_d2": ("ref_shim/gaussian2.cl (ours: a 2-D model to exercise DIM > 1)", "gaussian2_logpdf", 2)}


def available(stem: str) -> bool:
    return (REF_DIR / f"libref_{stem}.so").exists()


def load(stem: str) -> C.CDLL:
    lib = C.CDLL(str(REF_DIR / f"libref_{stem}.so"))
    lib.ref_dim.restype = C.c_int
    lib.ref_init_walkers.argtypes = [C.c_uint32, _f32, _f32, C.c_uint32]
    lib.ref_logfn.argtypes = [C.c_uint32, C.c_uint32, _f32, _f32, _f32, C.c_uint32]
    lib.ref_stretch_move_bare.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _f32, _f32, _f32, _f32,
                                          C.c_float, C.c_float, C.c_uint32, C.c_uint32]
    return lib


class ReferenceTextStretch(OracleStretch):
    """OracleStretch whose kernels are the reference's text.  NOTE the literal semantics of that text for DIM > 1
    (SURVEY Appendix B-1, B-2): the partner is an element offset and `logfn` fills only the first W entries."""

    def __init__(self, stem: str, model, seed: int, walkers: int, params, wgs: int = 256):
        super().__init__(model, seed, walkers, params, wgs=wgs, literal_partner=True)
        self.ref = load(stem)
        assert self.ref.ref_dim() == self.D

    def _logfn_all(self):
        self.ref.ref_logfn(self.data_len, self.params_len, self.params, self.xs, self.lp, self.W)

    def init_position(self, seed: int, limits: np.ndarray):
        lim = np.ascontiguousarray(limits, dtype=np.float32).reshape(-1)
        self.ref.ref_init_walkers(seed & 0xFFFFFFFF, lim, self.xs, self.W)
        self._logfn_all()
        self.iterations = 0
        return self

    def _half(self, half: int, seed: int, tag: int, step: int, a: float, beta: float, want_diag: bool = False):
        H, D = self.H, self.D
        act = self.xs[half * H * D:(half + 1) * H * D]
        cmp_ = self.xs[(1 - half) * H * D:(2 - half) * H * D]
        lp = self.lp[half * H:(half + 1) * H]
        self.ref.ref_stretch_move_bare(seed & 0xFFFFFFFF, tag, self.data_len, self.params_len, self.params, cmp_, act, lp,
                                       a, beta, step & 0xFFFFFFFF, H)
        return None, None
