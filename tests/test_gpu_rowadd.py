"""GPU parity of the GENERIC row-additive dataset likelihood (BAY_MODEL_ROW_ADDITIVE): the reference's library
likelihoods gaussian_loglik / student_t_loglik (K/cuda/distributions/gaussian.cu:36-46, student-t.cu:40-53, K =
/root/reference/src/device/uncomplicate/bayadera/internal/device/cuda) evaluated by streaming the dataset once per 128
walkers through shared memory instead of once per walker.

Oracle: the model's own serial LOGFN (the posterior template around the likelihood's loop) run by the CPU oracle.
Tolerance (north_star): per-walker logpdf within 1e-5 relative, accept decisions equal except at near-ties.
"""
import numpy as np
import pytest

import bayadera_b200 as bb
from bayadera_b200 import mcmc, models
from oracle import oracle as orc
from test_gpu_parity import check_steplocked, f32, logpdf_close

pytestmark = pytest.mark.gpu
WGS = 256


@pytest.fixture(scope="module")
def factory():
    f = bb.B200BayaderaFactory(device=0, wgs=WGS)
    yield f
    f.release()


def gaussian_data(n, mu=2.5, sd=1.7, seed=3):
    rng = np.random.default_rng(seed)
    return np.concatenate([(mu + sd * rng.standard_normal(n)).astype(np.float32), f32([0.0, 5.0, 0.5])])


def gaussian_posterior_fp64(params, xs):
    """The model in fp64 closed form: the yardstick where the reference's own fp32 `REAL acc` loop (and so the serial
    oracle) is the less accurate party — 2e4 sequential fp32 additions at |acc| ~ 5e4 wander by ~1e-5 relative."""
    data, (m0, s0, lam) = params[:-3].astype(np.float64), params[-3:].astype(np.float64)
    mu, sd = xs[:, 0].astype(np.float64), xs[:, 1].astype(np.float64)
    ss = ((data[None, :] - mu[:, None]) ** 2).sum(axis=1)
    lik = -ss / (2.0 * sd * sd) + data.size * (-np.log(sd) - 0.9189385332046727)
    prior = -0.5 * ((mu - m0) / s0) ** 2 - np.log(s0) - 0.9189385332046727 + np.log(lam) - lam * sd
    return lik + prior


@pytest.mark.parametrize("n,walkers", [(1, 512), (7, 512), (1000, 1024), (4099, 1536)])
def test_gaussian_posterior_logdensity_and_moves_match_oracle(factory, n, walkers):
    model = models.gaussian_mean_sd_posterior()
    assert model.flags & models.ROW_ADDITIVE
    params = gaussian_data(n)
    lim = model.limits_array()
    gpu = factory.mcmc_factory(model).create_sampler(5, walkers, params).init_position(6, lim)
    cpu = orc.OracleStretch(model, 5, walkers, params, wgs=WGS).init_position(6, lim)
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs)
    assert logpdf_close(st["logfn"], cpu.lp, rtol=1e-5).all()
    check_steplocked(gpu, cpu, steps=3, a=2.0, tie_tol=5e-3, exact_lp=False)


@pytest.mark.parametrize("n", [50_000, 1_000_000])
def test_gaussian_posterior_large_dataset_against_fp64(factory, n):
    """10^6 data: the tiled path (fp32 inside a 1024-row tile, double across tiles) against the fp64 closed form at
    1e-6 relative — ten times inside the north-star tolerance, and closer than the reference-style serial loop."""
    model = models.gaussian_mean_sd_posterior()
    params = gaussian_data(n, seed=17)
    gpu = factory.mcmc_factory(model).create_sampler(5, 1024, params).init_position(6, model.limits_array())
    xs, lp64 = gpu.get_state64()
    want = gaussian_posterior_fp64(params, xs)
    assert np.abs(lp64 / want - 1.0).max() < 1e-6
    gpu.burn_in(3, 2.0)
    xs, lp64 = gpu.get_state64()
    assert np.abs(lp64 / gaussian_posterior_fp64(params, xs) - 1.0).max() < 1e-6


def test_row_additive_path_equals_the_serial_kernel(factory):
    """Same model with and without the row metadata: the tiled path against the reference-style per-thread loop."""
    n, walkers = 20_000, 2048
    params = gaussian_data(n, seed=8)
    tiled = models.gaussian_mean_sd_posterior()
    serial = models.gaussian_mean_sd_posterior(row_additive=False)
    assert not serial.flags & models.ROW_ADDITIVE
    lim = tiled.limits_array()
    a = factory.mcmc_factory(tiled).create_sampler(1, walkers, params).init_position(2, lim)
    b = factory.mcmc_factory(serial).create_sampler(1, walkers, params).init_position(2, lim)
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa["xs"], sb["xs"])
    # 2e4 rows: the serial kernel's fp32 accumulator is itself good to ~3e-5 only; both against fp64
    want = gaussian_posterior_fp64(params, sa["xs"])
    assert np.abs(sa["logfn"] / want - 1.0).max() < 1e-6
    assert np.abs(sb["logfn"] / want - 1.0).max() < 1e-4
    a.burn_in(4, 2.0)
    b.burn_in(4, 2.0)
    sa, sb = a.get_state(), b.get_state()
    same = np.all(sa["xs"] == sb["xs"], axis=1)
    assert same.mean() > 0.97                        # the serial kernel's summation noise flips near-ties
    assert np.abs(sa["logfn"] / gaussian_posterior_fp64(params, sa["xs"]) - 1.0).max() < 1e-6


def test_student_t_posterior_matches_oracle(factory):
    model = models.student_t_posterior()
    rng = np.random.default_rng(4)
    data = (1.0 + 0.8 * rng.standard_t(5, 3000)).astype(np.float32)
    params = np.concatenate([data, f32([0.0, 5.0, 0.5])])
    lim = model.limits_array()
    gpu = factory.mcmc_factory(model).create_sampler(9, 1024, params).init_position(10, lim)
    cpu = orc.OracleStretch(model, 9, 1024, params, wgs=WGS).init_position(10, lim)
    st = gpu.get_state()
    assert logpdf_close(st["logfn"], cpu.lp, rtol=2e-5).all()      # n * norm vs norm added per row: fp32 reassociation
    check_steplocked(gpu, cpu, steps=2, a=2.0, tie_tol=1e-2, exact_lp=False)


def test_gaussian_posterior_recovers_the_data_moments(factory):
    n = 200_000
    model = models.gaussian_mean_sd_posterior()
    params = gaussian_data(n, mu=-1.25, sd=0.6, seed=12)
    s = factory.mcmc_factory(model).create_sampler(21, 4096, params).init_position(22, model.limits_array())
    mcmc.mix(s)
    s.burn_in(300, 2.0)
    res = s.run_sampler(64, 2.0)
    assert 0.2 < res["acceptance-rate"] < 0.9
    x = s.sample().astype(np.float64)
    data = params[:-3].astype(np.float64)
    # posterior of (mu, sigma) concentrates at the sample mean / sd with width ~ sd / sqrt(n)
    assert abs(x[:, 0].mean() - data.mean()) < 0.01 and abs(x[:, 1].mean() - data.std()) < 0.01
    assert x[:, 0].std() < 0.01 and x[:, 1].std() < 0.01
    h = s.histogram(4)
    assert np.allclose(h.pdf.sum(axis=1) * (h.limits[:, 1] - h.limits[:, 0]) / WGS, 1.0, atol=1e-3)
