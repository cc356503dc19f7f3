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


# ---- row f-4: HDI on the device, mix! in one call ---------------------------------------------------------------
def test_hdi_device_reference_goldens(factory):
    """T/util_test.clj:21-58 through bay_hdi_histogram (79 bins on [1, 7])."""
    from bayadera_b200 import util
    from bayadera_b200.engine import Histogram, hdi_histogram
    pdf, rank = f32(G.HDI_PDF), f32(G.HDI_BIN_RANK)
    h = Histogram(f32([G.HDI_LIMITS]), pdf[None, :], rank[None, :])
    asum = util.asum(pdf)
    for mass, scaled, want in G.HDI_RANK_COUNTS:
        (cnt, _), = hdi_histogram(factory, h, mass / asum if scaled else mass)
        assert cnt == want, (mass, scaled, cnt)
    for cnt, want in G.HDI_BINS.items():
        (got_cnt, regions), = hdi_histogram(factory, h, forced_counts=[cnt])
        assert got_cnt == cnt
        bw = (G.HDI_LIMITS[1] - G.HDI_LIMITS[0]) / 79
        bins = np.stack([(regions[:, 0] - 1.0) / bw, (regions[:, 1] - 1.0) / bw - 1.0], axis=1).reshape(-1)
        assert np.allclose(bins, want, atol=1e-4), (cnt, bins)
    for cnt, (want, tol) in G.HDI_REGIONS.items():
        (_, regions), = hdi_histogram(factory, h, forced_counts=[cnt])
        assert np.linalg.norm(regions.reshape(-1) - np.asarray(want)) < tol
        assert np.allclose(regions, util.hdi_regions(G.HDI_LIMITS, rank, cnt), atol=1e-6)


def test_hdi_of_sampler_histogram_matches_host_definition(factory):
    """B200Stretch.hdi (device, all dimensions at once) == util.clj's host definition applied to the same histogram."""
    from bayadera_b200 import util
    # the built-in Gaussian in 1-D plus the 30-D hierarchical model cover single- and many-dimension histograms
    for model, params, limits, walkers in [
            (models.GAUSSIAN, f32([3, 1]), f32([-7, 7]), 4096),
            (models.therapeutic_touch_model(), models.therapeutic_touch_data(),
             models.therapeutic_touch_model().limits_array(), 2048)]:
        s = factory.mcmc_factory(model).create_sampler(7, walkers, params).init_position(8, limits)
        s.burn_in(200, 2.0)
        h = s.histogram(4)
        got = s.hdi(0.95)
        assert len(got) == model.dimension
        for d, (cnt, regions) in enumerate(got):
            assert cnt == util.hdi_rank_count(h.bin_ranks[d], h.pdf[d], 0.95)
            assert np.allclose(regions, util.hdi(h, d, 0.95), rtol=1e-6, atol=1e-6)
            assert 1 <= cnt <= G.WGS and (regions[:, 0] < regions[:, 1]).all()
    # Gaussian(3,1): the 95 % HDI is about [3 - 1.96, 3 + 1.96]
    s = factory.mcmc_factory(models.GAUSSIAN).create_sampler(7, 8192, f32([3, 1])).init_position(8, f32([-7, 7]))
    s.burn_in(300, 2.0)
    s.histogram(16)
    (_, regions), = s.hdi(0.95)
    assert abs(regions[0, 0] - (3 - 1.96)) < 0.25 and abs(regions[-1, 1] - (3 + 1.96)) < 0.25


@pytest.mark.parametrize("schedule", ["minus_n", "sqrt_n", "pow_n"])
def test_mix_in_one_call_equals_protocol_loop(factory, schedule):
    """bay_mix == mcmc.mix (C/mcmc.clj:66-101 as ~70 protocol calls): same chain, same tuned a, same rates."""
    from bayadera_b200 import mcmc
    sched = {"minus_n": mcmc.minus_n, "sqrt_n": mcmc.sqrt_n, "pow_n": mcmc.pow_n(0.7)}[schedule]
    model, params = models.GAUSSIAN, f32([200, 1])
    out = []
    for one_call in (True, False):
        s = factory.mcmc_factory(model).create_sampler(3, 2 * G.W, params).init_position(4, f32([180, 220]))
        opts = {"step": 32, "a": 2.0, "cooling-schedule": sched}   # a = 2 accepts too often in 1-D: the tuner must move it
        res = s.mix(opts) if one_call else mcmc.mix(s, opts)
        out.append((res, s.get_state()))
    (ra, sa), (rb, sb) = out
    assert ra == rb, (ra, rb)
    assert ra["a"] > 2.0 and 0.2 <= ra["acc-rate"] < ra["acc-rate-2.0"], ra
    assert np.array_equal(sa["xs"], sb["xs"]) and np.array_equal(sa["logfn"], sb["logfn"])


def test_density_and_evidence_engines_on_device_blocks():
    """The reference hands cuda-float blocks to log-density / density / evidence (nvidia_gtx.clj:83-141): the _dev entry
    points must give what the host-array ones give."""
    import torch
    from bayadera_b200.engines import B200DistributionEngine, B200LikelihoodEngine
    with bb.B200BayaderaFactory(device=0, wgs=256) as factory:
        dist = B200DistributionEngine(factory, models.GAUSSIAN)
        params = np.asarray([1.5, 0.7], dtype=np.float32)
        x = np.linspace(-3, 5, 4001, dtype=np.float32)
        want_log, want_pdf = dist.log_density(params, x), dist.density(params, x)
        p_d, x_d = torch.from_numpy(params).cuda(), torch.from_numpy(x).cuda()
        out = torch.empty(x.size, dtype=torch.float32, device="cuda")
        dist.density_dev(p_d.data_ptr(), p_d.numel(), x_d.data_ptr(), x.size, out.data_ptr())
        factory.synchronize()
        assert np.array_equal(out.cpu().numpy(), want_log)
        dist.density_dev(p_d.data_ptr(), p_d.numel(), x_d.data_ptr(), x.size, out.data_ptr(), exponentiate=True)
        factory.synchronize()
        assert np.array_equal(out.cpu().numpy(), want_pdf)
        lik = B200LikelihoodEngine(factory, models.BINOMIAL)
        data = models.binomial_lik_params(50, 15)
        pts = np.random.default_rng(0).beta(3, 2, 8192).astype(np.float32)
        d_d, t_d = torch.from_numpy(data).cuda(), torch.from_numpy(pts).cuda()
        assert lik.evidence_dev(d_d.data_ptr(), d_d.numel(), t_d.data_ptr(), pts.size) == lik.evidence(data, pts)
        dist.release()
        lik.release()
