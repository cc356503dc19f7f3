"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and NVRTC builds every model for sm_100a (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import bayadera_b200 as bb
from bayadera_b200 import _lib, models

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "bayadera_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bay_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for name in syms:
        assert hasattr(L, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def _cuda():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_cuda(), reason="this checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(bb.BayaderaError, match="no CPU fallback|CUDA"):
        bb.B200BayaderaFactory(device=0, wgs=256)


def _compile_check(model, wgs=256):
    L = _lib.load()
    srcs = (C.c_char_p * len(model.source))(*[s.encode() for s in model.source])
    n = C.c_int64()
    log = C.create_string_buffer(1 << 16)
    rc = L.bay_model_compile_check(srcs, len(model.source), model.mcmc_logpdf.encode(), model.dimension, wgs,
                                   model.flags, C.byref(n), log, len(log))
    return rc, n.value, log.value.decode("utf-8", "replace")


ALL_MODELS = list(models.DISTRIBUTIONS.values()) + [
    models.beta_binomial_posterior(),
    models.posterior_model(models.GAUSSIAN, "gg", models.GAUSSIAN),
    models.gaussian_mean_sd_posterior(),          # generic row-additive path (bay_rowadd_loglik)
    models.student_t_posterior(),
    models.therapeutic_touch_model(),
    models.logistic_regression_model(64),
    models.mvn_model(100),
]


@pytest.mark.parametrize("model", ALL_MODELS, ids=lambda m: m.name)
def test_nvrtc_builds_model_for_sm100a(model):
    rc, nbytes, log = _compile_check(model)
    assert rc == 0, log
    assert nbytes > 1000
    assert "sm_100a" in log                      # ptxas -v report names the target
    for kernel in ("bay_stretch_bare", "bay_stretch_accu", "bay_logfn"):
        assert kernel in log


@pytest.mark.parametrize("wgs", [32, 512, 1024])
@pytest.mark.parametrize("model", [models.GAUSSIAN, models.therapeutic_touch_model(), models.logistic_regression_model(64)],
                         ids=lambda m: m.name)
def test_nvrtc_builds_at_every_work_group_size(model, wgs):
    """The reference's default WGS is max-block-dim-x = 1024 (nvidia_gtx.clj:803-807): the accu kernels' staging tile
    must fit static shared memory there (it is capped at 8 dimensions per round)."""
    rc, nbytes, log = _compile_check(model, wgs=wgs)
    assert rc == 0, log
    assert "bay_stretch_accu" in log


def test_nvrtc_builds_wide_glm_model():
    """DIM = 192 > the tensor-core path's 128: the SIMT likelihood's row tile shrinks so that it still fits."""
    rc, _, log = _compile_check(models.logistic_regression_model(192))
    assert rc == 0, log
    assert "bay_glm_loglik" in log


def test_nvrtc_error_is_reported():
    bad = models.DeviceModel("bad", ("extern \"C\" { inline REAL bad_logpdf(int x) { return undefined_symbol; } }",),
                             "bad_logpdf")
    rc, _, log = _compile_check(bad)
    assert rc == _lib.ECOMPILE
    assert "undefined_symbol" in log


def test_wgs_validation():
    rc, _, log = _compile_check(models.GAUSSIAN, wgs=100)
    assert rc == _lib.EINVAL and "power of two" in log


def test_parameter_vectors():
    p = models.beta_params(3, 2)
    assert p.dtype == np.float32 and np.isclose(p[2], np.log(12.0), rtol=1e-6)   # -lbeta(3,2) = log 12
    v, mu, sigma = models.mvn_params(8, seed=1)
    u = v[8:].reshape(8, 8).astype(np.float64)
    assert np.allclose(u, np.triu(u))
    assert np.allclose(np.linalg.inv(u.T @ u), sigma, rtol=2e-3, atol=1e-3)


# ---- util.clj host helpers (T/util_test.clj goldens) ----------------------------------------------------------
def test_hdi_host_helpers_match_reference_goldens():
    import goldens as G
    from bayadera_b200 import util
    pdf = np.asarray(G.HDI_PDF, dtype=np.float32)
    rank = np.asarray(G.HDI_BIN_RANK, dtype=np.float32)
    asum = util.asum(pdf)
    for mass, scaled, want in G.HDI_RANK_COUNTS:
        assert util.hdi_rank_count(rank, pdf, mass / asum if scaled else mass) == want
    for cnt, want in G.HDI_BINS.items():
        assert util.hdi_bins(rank, cnt) == want
    for cnt, (want, tol) in G.HDI_REGIONS.items():
        got = util.hdi_regions(G.HDI_LIMITS, rank, cnt).reshape(-1)
        assert np.linalg.norm(got - np.asarray(want)) < tol
    assert util.bin_mapper(79, 1.0, 7.0)(0) == pytest.approx(1.0 + 0.5 * 6.0 / 79)
    assert util.range_mapper(0.0, 1.0, 10.0, 20.0)(0.25) == 12.5


def test_mix_schedules_expose_their_kind():
    from bayadera_b200 import mcmc
    assert mcmc.pow_n(0.5).pow_n_power == 0.5
    assert mcmc.pow_n(0.5)(9.0)(5) == 2.0 and mcmc.sqrt_n(9.0)(5) == 2.0 and mcmc.minus_n(9.0)(5) == 4.0


# ---- the Clojure binding (src/clojure) stays in step with the header ------------------------------
def test_clojure_binding_declares_only_header_symbols_and_covers_the_protocols():
    clj = (ROOT / "src" / "clojure" / "uncomplicate" / "bayadera" / "internal" / "device" / "b200.clj").read_text()
    declared = set(re.findall(r"\(\^\w+ (bay_[a-z0-9_]+) \[", clj))
    called = set(re.findall(r"\(\.(bay_[a-z0-9_]+) bay", clj))
    header = set(header_symbols())
    assert declared and declared <= header, sorted(declared - header)
    assert called <= declared, sorted(called - declared)
    # every protocol method of protocols.clj:20-138 that the GTX engines implement has its entry point bound
    for sym in ("bay_engine_create_current", "bay_model_compile", "bay_sampler_create_dev", "bay_init",
                "bay_init_position_uniform", "bay_init_position_from", "bay_burn_in", "bay_anneal", "bay_acc_rate",
                "bay_run_sampler", "bay_init_move", "bay_move", "bay_move_bare", "bay_set_temperature", "bay_sample",
                "bay_histogram", "bay_mean", "bay_variance", "bay_sd", "bay_info", "bay_dataset_mean",
                "bay_dataset_variance", "bay_dataset_histogram", "bay_acor", "bay_model_density_dev", "bay_model_evidence_dev",
                "bay_direct_sample", "bay_sampler_release", "bay_model_release",
                "bay_engine_release"):
        assert sym in called, sym
    for proto in ("MCMC", "MCMCStretch", "RandomSampler", "EstimateEngine", "Location", "Spread", "DatasetEngine",
                  "AcorEngine", "DensityEngine", "LikelihoodEngine", "RandomSamplerEngine", "SamplerFactory",
                  "EngineFactory", "ModelProvider", "Releaseable", "Info"):
        assert re.search(rf"^\s+{proto}\s*$", clj, flags=re.M), proto
    entry = (ROOT / "src" / "clojure" / "uncomplicate" / "bayadera" / "b200.clj").read_text()
    assert "with-default-bayadera" in entry and "b200-bayadera-factory" in entry
