"""GPU parity for the callers either side of the hot path: direct samplers against the
reference goldens and the oracle; density / evidence engines against closed forms and the reference's values
(T/internal/nvidia_gtx_test.clj:30-159)."""
import math

import numpy as np
import pytest

import goldens as G
import bayadera_b200 as bb
from bayadera_b200 import models
from oracle import oracle_rng

pytestmark = pytest.mark.gpu


def f32(x):
    return np.asarray(x, dtype=np.float32)


@pytest.fixture(scope="module")
def factory():
    f = bb.B200BayaderaFactory(device=0, wgs=G.WGS)
    yield f
    f.release()


def test_direct_uniform_bit_exact(factory):
    g = G.DIRECT_UNIFORM
    x = bb.B200DirectSamplerEngine(factory, "uniform").sample(G.SEED, g["params"], g["n"])
    assert np.array_equal(x[:4], f32(g["first4"])) and np.array_equal(x[-4:], f32(g["last4"]))
    assert x.max() == np.float32(g["max"]) and x.min() == np.float32(g["min"])
    assert np.array_equal(x, oracle_rng.direct_sample("uniform", g["n"], G.SEED, g["params"]))


@pytest.mark.parametrize("family", ["gaussian", "erlang", "exponential"])
def test_direct_samplers_vs_goldens_and_oracle(factory, family):
    g = G.DIRECT[family]
    x = bb.B200DirectSamplerEngine(factory, family).sample(G.SEED, g["params"], 10000)
    exact = int((x[:4] == f32(g["first4"])).sum() + (x[-4:] == f32(g["last4"])).sum())
    print(f"{family}: {exact}/8 pinned values bit-exact vs the reference's fast-math goldens")
    assert np.allclose(x[:4], f32(g["first4"]), rtol=2e-5) and np.allclose(x[-4:], f32(g["last4"]), rtol=2e-5)
    assert abs(float(x.max()) / g["max"] - 1) < 2e-5 and abs(float(x.min()) / g["min"] - 1) < 1e-3
    assert abs(float(x.astype(np.float64).mean()) / g["mean"] - 1) < 1e-5
    ref = oracle_rng.direct_sample(family, 10000, G.SEED, g["params"])
    scale = float(np.abs(f32(g["params"])).max())          # libm vs MUFU sin/cos/lg2: absolute error ~1e-6 * scale
    assert np.allclose(x, ref, rtol=5e-5, atol=5e-6 * max(scale, 1.0))


def test_direct_sampler_rejects_ragged_n(factory):
    with pytest.raises(bb.BayaderaError, match="multiple of 4"):
        bb.B200DirectSamplerEngine(factory, "uniform").sample(1, [0, 1], 10)


def test_gaussian_density_engine(factory):
    """nvidia_gtx_test.clj:111-131: pdf / logpdf of N(0,1) on 200 points vs the closed form."""
    eng = bb.B200DistributionEngine(factory, models.GAUSSIAN)
    x = f32(np.arange(-10, 10, 0.1))[:200]
    logp = eng.log_density(f32([0.0, 1.0]), x)
    pdf = eng.density(f32([0.0, 1.0]), x)
    want_log = -0.5 * x.astype(np.float64) ** 2 - 0.5 * math.log(2 * math.pi)
    assert np.linalg.norm(logp - want_log) < 1e-4
    assert np.linalg.norm(pdf - np.exp(want_log)) < 1e-5


def test_beta_binomial_posterior_density_and_evidence(factory):
    """nvidia_gtx_test.clj:133-159: posterior logpdf = Beta(18,37) logpdf - 32.61044; evidence 1.63574e-15."""
    n, z, a, b = 50, 15, 3, 2
    post = bb.B200DistributionEngine(factory, models.beta_binomial_posterior())
    beta = bb.B200DistributionEngine(factory, models.BETA)
    lik = bb.B200LikelihoodEngine(factory, models.BINOMIAL)
    x = f32(np.arange(0.001, 1, 0.001))[:200]
    params = np.concatenate([models.binomial_lik_params(n, z), models.beta_params(a, b)])
    post_log = post.log_density(params, x)
    beta_log = beta.log_density(models.beta_params(a + z, b + n - z), x)
    assert np.linalg.norm((beta_log - post_log) - 32.61044) < 1e-3 * math.sqrt(200) * 0.1
    ev = lik.evidence(models.binomial_lik_params(n, z), x)
    assert abs(ev / 1.6357453252754427e-15 - 1) < 1e-4
    lik_vals = lik.density(models.binomial_lik_params(n, z), x)
    assert abs(float(lik_vals.astype(np.float64).mean()) / ev - 1) < 1e-5
