"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle and the reference's goldens.

Bar (north_star): Philox / proposal / init / histogram-count arithmetic bit-exact; per-walker logpdf within
1e-5 relative; accept decisions equal except at near-ties |u - q| <= tie_tol * q (fast-math exp/pow on the GPU vs
libm on the CPU); chain-level results statistical.
"""
import numpy as np
import pytest

import goldens as G
import bayadera_b200 as bb
from bayadera_b200 import mcmc, models
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

LOGPDF_RTOL = 1e-5       # north_star: per-walker logpdf within 1e-5 relative
TIE_TOL = 2e-4           # near-tie band on the accept ratio (fast-math exp: ~2 ulp * |arg|)


def f32(x):
    return np.asarray(x, dtype=np.float32)


@pytest.fixture(scope="module")
def factory():
    f = bb.B200BayaderaFactory(device=0, wgs=G.WGS)
    yield f
    f.release()


def make_pair(factory, model, seed, walkers, params, limits, wgs=G.WGS):
    sf = factory.mcmc_factory(model)
    gpu = sf.create_sampler(seed, walkers, params)
    cpu = orc.OracleStretch(model, seed, walkers, params, wgs=wgs)
    gpu.init(seed).init_position(seed, limits)
    cpu.init(seed).init_position(seed, limits)
    return sf, gpu, cpu


def logpdf_close(a, b, rtol=LOGPDF_RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & (a == b)
    fin = np.isfinite(a) & np.isfinite(b)
    ok = both_nan | same_inf | (fin & (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + 1e-6))
    return ok


# ------------------------------------------------------------------ reference goldens on the GPU --
def test_golden_uniform_positions(factory):
    sf, gpu, _ = make_pair(factory, models.UNIFORM, G.SEED, G.W, f32([-1, 2]), f32([-1, 2]))
    for want in G.UNIFORM_SAMPLES:
        got = gpu.sample()[:4, 0]
        assert np.array_equal(got, f32(want)), (got, want)
    assert gpu.info() == {"walker-count": G.W, "iteration-counter": 4}


def test_golden_raw_launches_and_accu_step(factory):
    sf, gpu, _ = make_pair(factory, models.UNIFORM, G.SEED, G.W, f32([-1, 2]), f32([-1, 2]))
    # raw launches with explicit step counters (nvidia_gtx_test.clj:217-235)
    for step, want in enumerate(G.UNIFORM_SAMPLES[:2]):
        gpu.set_state(bare_counter=step)
        gpu.move_bare_half(0).move_bare_half(1)
        assert np.array_equal(gpu.get_state()["xs"][:4, 0], f32(want))
    # one accu step with seeds (123, 124), tags 1111/2222, step 0 (nvidia_gtx_test.clj:237-258)
    gpu.set_state(move_seed=G.SEED - 2)
    gpu.init_move(G.A).move()
    assert np.array_equal(gpu.get_state()["xs"][:4, 0], f32(G.UNIFORM_ACCU_XS))
    accept, sums = gpu.accu_blocks()
    assert tuple(int(v) for v in accept[:10]) == G.UNIFORM_ACCU_ACCEPT
    assert np.array_equal(sums[0, :10], f32(G.UNIFORM_ACCU_BLOCK_SUMS))
    assert abs(float(sums.astype(np.float64).sum()) - G.UNIFORM_ACCU_TOTAL) < 2e-3
    assert abs(float(gpu.last_means(1)[0, 0]) - G.UNIFORM_ACCU_TOTAL / G.W) < 1e-6


def test_golden_gaussian_positions(factory):
    sf, gpu, _ = make_pair(factory, models.GAUSSIAN, G.SEED, G.W, f32([3, 1.0]), f32([-7, 7]))
    for want in G.GAUSSIAN_SAMPLES:
        got = gpu.sample()[:4, 0]
        assert np.array_equal(got, f32(want)), (got, want)


def test_golden_burn_in_summary(factory):
    sf = factory.mcmc_factory(models.GAUSSIAN)
    gpu = sf.create_sampler(G.SEED, G.W, f32([3, 1.0]))
    gpu.init_position(G.SEED, f32([-7, 7]))
    gpu.init(G.SEED + 1)
    gpu.burn_in(100, 1.5)
    ds = factory.dataset_engine()
    assert abs(float(ds.data_mean(gpu.sample())[0]) - G.BURN_IN_CUDA["mean"]) < 2e-3
    strided = gpu.sample()[::1500, 0]
    assert np.allclose(strided, f32(G.BURN_IN_CUDA["strided"]), atol=5e-3)
    assert abs(float(np.sqrt(ds.data_variance(gpu.sample())[0])) - G.BURN_IN_CUDA["sd"]) < 2e-3


def test_acc_rate_and_tau(factory):
    sf = factory.mcmc_factory(models.GAUSSIAN)
    gpu = sf.create_sampler(G.SEED, 2 * G.W, f32([200, 1]))
    gpu.init(G.SEED).init_position(G.SEED, f32([180.0, 220.0]))
    gpu.burn_in(5120, 8.0)
    assert abs(gpu.acc_rate(8.0) - G.ACC_RATE_OPENCL) < 1.5e-3
    res = gpu.run_sampler(63670, 8.0)
    assert abs(float(res["autocorrelation"].tau[0]) - G.TAU) < 0.5
    assert abs(float(res["autocorrelation"].mean[0]) - 200.0) < 0.01
    assert 0.48 < res["acceptance-rate"] < 0.49
    assert gpu.info()["iteration-counter"] == 5120 + 1 + 63670


@pytest.mark.parametrize("n", [67, 367, 112640])
def test_acor_fixtures(factory, n):
    series = G.acor_fixture(n)
    ac = factory.acor_engine().acor(series)
    tau, mean, sigma, lag = orc.acor(series, 2, n, G.WGS)
    assert ac.lag == lag and ac.steps == n
    assert np.allclose(ac.tau, tau, rtol=2e-5) and np.allclose(ac.sigma, sigma, rtol=2e-5)
    assert np.allclose(ac.mean, mean, rtol=1e-6, atol=1e-7)
    want = G.ACOR[n]
    if n == 112640:
        assert abs(float(ac.tau[1]) - want["tau"]) < 1e-3 and abs(float(ac.sigma[0]) - want["sigma"]) < 1e-3
    else:
        assert np.allclose(ac.tau, want["tau"], rtol=2e-5) and abs(float(ac.sigma[0]) - want["sigma"]) < 2e-5


def test_acor_too_short(factory):
    with pytest.raises(bb.AcorTooShortError, match="must not be less than 50"):
        factory.acor_engine().acor(np.zeros((40, 1), dtype=np.float32))


# --------------------------------------------------------------------------- kernel-level parity --
MODEL_CASES = [
    ("gaussian", models.GAUSSIAN, f32([3, 1.0]), f32([-7, 7])),
    ("student_t", models.STUDENT_T, f32([5, 1.0, 2.0, 0.0]), f32([-10, 12])),
    ("beta", models.BETA, models.beta_params(3, 2), f32([0, 1])),
    ("gamma", models.GAMMA, f32([2.0, 3.0, 0.0]), f32([0.01, 20])),
    ("erlang", models.ERLANG, f32([2.0, 3.0, 0.0]), f32([0.01, 10])),
    ("exponential", models.EXPONENTIAL, f32([4.0, np.log(4.0)]), f32([-0.5, 3])),   # part of the box is out of support
    ("beta_binomial", models.beta_binomial_posterior(),
     np.concatenate([models.binomial_lik_params(50, 15), models.beta_params(3, 2)]), f32([0, 1])),
]


@pytest.mark.parametrize("name,model,params,limits", MODEL_CASES, ids=[c[0] for c in MODEL_CASES])
def test_init_logfn_and_steplocked_moves(factory, name, model, params, limits):
    W = 4 * G.WGS * 8
    sf, gpu, cpu = make_pair(factory, model, 77, W, params, limits)
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs), "init_walkers must be bit-exact"
    assert logpdf_close(st["logfn"], cpu.lp).all(), "logfn kernel vs oracle"
    check_steplocked(gpu, cpu, steps=6, a=2.0)


def check_steplocked(gpu, cpu, steps, a, beta_t=1.0, tie_tol=TIE_TOL, exact_lp=True):
    """Per half-step: both engines start from the oracle's state; compare accept masks (modulo near-ties),
    positions of agreeing walkers bit-for-bit and the accepted proposals' log-densities.
    exact_lp=False for GLM samplers, which recompute their (double) log-densities on set_state."""
    D, H = cpu.D, cpu.H
    cpu.a_bare = a
    cpu.set_temperature(beta_t)
    gpu.set_a(a).set_temperature(beta_t)
    mism_total = 0
    for step in range(steps):
        for half in (0, 1):
            before_x, before_lp = cpu.xs.copy(), cpu.lp.copy()
            gpu.set_state(xs=before_x, logfn=before_lp, bare_counter=cpu.bare_counter)
            cpu.half_bare(half, want_diag=True)
            gpu.move_bare_half(half)
            g = gpu.get_state()
            gx, glp = g["xs"].reshape(-1, D), g["logfn"]
            cx = cpu.xs.reshape(-1, D)
            sl = slice(half * H, (half + 1) * H)
            other = slice((1 - half) * H, (2 - half) * H)
            assert np.array_equal(gx[other], cx[other]), "complementary half must be untouched"
            acc_cpu = cpu.diag["acc"].astype(bool)
            q, uz = cpu.diag["q"].astype(np.float64), cpu.diag["uz"].astype(np.float64)
            near_tie = np.abs(uz - q) <= tie_tol * np.maximum(q, 1e-30)
            same_pos = np.all(gx[sl] == cx[sl], axis=1)
            bad = ~same_pos & ~near_tie
            assert not bad.any(), (f"step {step} half {half}: {bad.sum()} walkers differ outside the near-tie band; "
                                   f"first k={np.flatnonzero(bad)[:5]}, q={q[bad][:5]}, uz={uz[bad][:5]}")
            mism_total += int((~same_pos).sum())
            agree_acc = same_pos & acc_cpu
            assert logpdf_close(glp[sl][agree_acc], cpu.lp[sl][agree_acc]).all(), "accepted log-density parity"
            kept = same_pos & ~acc_cpu
            if exact_lp:
                assert np.array_equal(glp[sl][kept], before_lp[sl][kept], equal_nan=True)
            else:
                assert logpdf_close(glp[sl][kept], before_lp[sl][kept]).all()
        cpu.bare_counter += 1
    print(f"step-locked: {mism_total} near-tie accept mismatches in {steps * 2 * H} walker-steps (band {tie_tol:g})")
    assert mism_total <= max(2, int(2e-4 * steps * 2 * H)), f"too many near-tie mismatches: {mism_total}"


def test_steplocked_annealed_and_large_a(factory):
    model, params = models.GAUSSIAN, f32([200, 1])
    sf, gpu, cpu = make_pair(factory, model, 5, 2048, params, f32([180, 220]))
    check_steplocked(gpu, cpu, steps=4, a=8.0, beta_t=7.0)
    check_steplocked(gpu, cpu, steps=4, a=1.5, beta_t=1.0)


def test_therapeutic_touch_d30(factory):
    model = models.therapeutic_touch_model()
    params = models.therapeutic_touch_data()
    sf, gpu, cpu = make_pair(factory, model, 11, 4096, params, model.limits_array())
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs)
    assert logpdf_close(st["logfn"], cpu.lp, rtol=2e-5).all()
    check_steplocked(gpu, cpu, steps=4, a=2.0, tie_tol=2e-3)   # z^(D-1) with D = 30 widens the fast-math band


def test_mvn_d100(factory):
    model = models.mvn_model(100)
    params, _, _ = models.mvn_params(100)
    sf, gpu, cpu = make_pair(factory, model, 3, 2048, params, model.limits_array())
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs)
    assert logpdf_close(st["logfn"], cpu.lp, rtol=5e-5).all()    # 5050-term fp32 sums, fma vs mul+add
    check_steplocked(gpu, cpu, steps=2, a=1.2, tie_tol=2e-3)


def test_logistic_regression_small(factory):
    d, rows = 8, 500
    rng = np.random.default_rng(2024)
    x = rng.standard_normal((rows, d)).astype(np.float32)
    theta = (rng.standard_normal(d) / np.sqrt(8)).astype(np.float32)
    y = (rng.random(rows) < 1 / (1 + np.exp(-(x @ theta)))).astype(np.float32)
    data = np.concatenate([y[:, None], x], axis=1).reshape(-1)
    model = models.logistic_regression_model(d)
    params = np.concatenate([data, f32([1.0 / (2 * 10.0 ** 2)])])
    sf, gpu, cpu = make_pair(factory, model, 9, 1024, params, model.limits_array())
    st = gpu.get_state()
    assert logpdf_close(st["logfn"], cpu.lp, rtol=2e-5).all()
    check_steplocked(gpu, cpu, steps=3, a=2.0, tie_tol=5e-3, exact_lp=False)


def test_uniform_chain_bit_exact_over_many_steps(factory):
    """The uniform model's accept test is exact (q is 0 or 1), so whole chains must agree bit-for-bit."""
    sf, gpu, cpu = make_pair(factory, models.UNIFORM, G.SEED, G.W, f32([-1, 2]), f32([-1, 2]))
    gpu.burn_in(200, 2.0)
    cpu.burn_in(200, 2.0)
    gpu.anneal(mcmc.minus_n(50.0), 50, 3.0)
    cpu.anneal(mcmc.minus_n(50.0), 50, 3.0)
    assert np.array_equal(gpu.get_state()["xs"].reshape(-1), cpu.xs)
    a, b = gpu.sample(3 * G.W + 100), cpu.sample(3 * G.W + 100)
    assert np.array_equal(a, b)
    assert gpu.info()["iteration-counter"] == cpu.iterations == 254


def test_run_sampler_means_and_accept_parity_uniform(factory):
    sf, gpu, cpu = make_pair(factory, models.UNIFORM, G.SEED, G.W, f32([-1, 2]), f32([-1, 2]))
    rg = gpu.run_sampler(64, 2.0)
    rc = cpu.run_sampler(64, 2.0)
    assert rg["acceptance-rate"] == rc["acceptance-rate"]            # integer counts: exact
    acc_g, sums_g = gpu.accu_blocks()
    assert np.array_equal(acc_g, cpu.accept)
    assert np.array_equal(sums_g.reshape(-1), cpu.blk_sums)          # block tree order: exact
    assert np.allclose(gpu.last_means(64), rc["means"], rtol=1e-6, atol=1e-7)
    ag, ac = rg["autocorrelation"], rc["autocorrelation"]
    assert ag.lag == ac["lag"] and np.allclose(ag.tau, ac["tau"], rtol=1e-3)


# --------------------------------------------------------------------------------- estimate engine --
def test_histogram_counts_bit_exact_with_cycles(factory):
    sf, gpu, cpu = make_pair(factory, models.UNIFORM, G.SEED, G.W, f32([-1, 2]), f32([-1, 2]))
    gpu.burn_in(20, 2.0)
    cpu.burn_in(20, 2.0)
    hg = gpu.histogram(5)
    hc = cpu.histogram(5)
    assert np.array_equal(hg.limits, hc["limits"])
    assert np.array_equal(gpu.histogram_counts(), hc["counts"])
    assert int(gpu.histogram_counts().sum()) == 5 * G.W
    assert np.array_equal(hg.pdf, hc["pdf"])
    assert np.array_equal(hg.bin_ranks, hc["bin-ranks"])
    assert gpu.info()["iteration-counter"] == 24


def test_histogram_and_moments_multidim(factory):
    model = models.mvn_model(100)
    params, mu, sigma = models.mvn_params(100)
    sf, gpu, cpu = make_pair(factory, model, 3, 8192, params, model.limits_array())
    st = gpu.get_state()
    hg = gpu.histogram(1)
    xs = st["xs"]
    limits = orc.min_max(xs, 100, 8192)
    assert np.array_equal(hg.limits.reshape(-1), limits)
    counts = orc.histogram_counts(xs, 100, 8192, G.WGS, limits).reshape(100, G.WGS)
    assert np.array_equal(gpu.histogram_counts(), counts)
    m, v = orc.mean_variance(xs, 100, 8192)
    assert np.allclose(gpu.mean(), m, rtol=1e-5, atol=1e-5)
    assert np.allclose(gpu.variance(), v, rtol=1e-5)
    assert np.allclose(gpu.sd(), np.sqrt(v), rtol=1e-5)


def test_dataset_engine_reference_case(factory):
    """T/core_test.clj:18-30: 22 x (31*2^16) random matrix; histogram integrates to 1, moments match."""
    n, m = 31 * 2 ** 16, 22
    rng = np.random.default_rng(7)
    data = rng.random((n, m), dtype=np.float32)
    ds = factory.dataset_engine()
    h, counts = ds.histogram(data, with_counts=True)
    assert abs(float(h.pdf[4].sum()) / G.WGS - 1.0) < 1e-3            # the reference's own check (limits ~ [0,1])
    lim = orc.min_max(data, m, n)
    assert np.array_equal(h.limits.reshape(-1), lim)
    want = orc.histogram_counts(data, m, n, G.WGS, lim).reshape(m, G.WGS)
    assert np.array_equal(counts, want)
    mean, var = ds.data_mean(data), ds.data_variance(data)
    assert abs(float((data.astype(np.float64).mean(axis=0) - mean).sum())) < 0.003
    assert abs(float((data.astype(np.float64).var(axis=0) - var).sum())) < 0.003
    om, ov = orc.mean_variance(data, m, n)
    assert np.allclose(mean, om, rtol=1e-6) and np.allclose(var, ov, rtol=1e-5)


@pytest.mark.parametrize("n,m", [(1, 1), (3, 2), (257, 5), (1000, 33)])
def test_dataset_engine_ragged_shapes(factory, n, m):
    rng = np.random.default_rng(n * 31 + m)
    data = (rng.standard_normal((n, m)) * 3 + 1).astype(np.float32)
    ds = factory.dataset_engine()
    om, ov = orc.mean_variance(data, m, n)
    assert np.allclose(ds.data_mean(data), om, rtol=1e-5, atol=1e-6)
    assert np.allclose(ds.data_variance(data), ov, rtol=1e-4, atol=1e-6)
    if n > 1:
        h, counts = ds.histogram(data, with_counts=True)
        lim = orc.min_max(data, m, n)
        assert np.array_equal(h.limits.reshape(-1), lim)
        assert np.array_equal(counts, orc.histogram_counts(data, m, n, G.WGS, lim).reshape(m, G.WGS))


@pytest.mark.parametrize("wgs", [64, 256])
@pytest.mark.parametrize("n,m", [(40_003, 100), (9_000, 64), (70_001, 3), (5_000, 36)])
def test_dataset_histogram_repeated_bins(n, m, wgs):
    """The lane-private histogram kernel adds, per lane, the number of its in-flight rows that share a bin: data with
    few distinct values (whole runs of rows in one bin), a constant column and a +-inf / NaN sprinkle must still give
    the oracle's integer counts."""
    rng = np.random.default_rng(n + m + wgs)
    data = rng.integers(0, 7, size=(n, m)).astype(np.float32)          # 7 distinct values per column
    data[:, 1] = 2.5                                                     # zero range
    data[::3, 2] = rng.standard_normal(len(data[::3, 2])).astype(np.float32)
    if m > 40:
        data[:, 40] = np.repeat(rng.standard_normal(n // 8 + 1), 8)[:n].astype(np.float32)   # runs of 8 equal rows
        data[5, 33], data[6, 34], data[7, 35] = np.inf, -np.inf, np.nan
    f = bb.B200BayaderaFactory(device=0, wgs=wgs)
    try:
        h, counts = f.dataset_engine().histogram(data, with_counts=True)
        lim = orc.min_max(data, m, n)
        assert np.array_equal(h.limits.reshape(-1), lim, equal_nan=True)
        want = orc.histogram_counts(data, m, n, wgs, lim).reshape(m, wgs)
        assert np.array_equal(counts, want)
        assert int(counts.sum()) == n * m
    finally:
        f.release()


def test_dataset_histogram_more_than_256_bins_uses_the_atomic_kernel():
    """WGS = 512 bins: the lane-private kernel covers up to 256, so this is k_hist_aos (shared-memory atomics)."""
    n, m, wgs = 30_011, 40, 512
    rng = np.random.default_rng(5)
    data = (rng.standard_normal((n, m)) * 2 + 1).astype(np.float32)
    data[:, 3] = rng.integers(0, 5, n).astype(np.float32)
    f = bb.B200BayaderaFactory(device=0, wgs=wgs)
    try:
        h, counts = f.dataset_engine().histogram(data, with_counts=True)
        lim = orc.min_max(data, m, n)
        assert np.array_equal(h.limits.reshape(-1), lim)
        assert np.array_equal(counts, orc.histogram_counts(data, m, n, wgs, lim).reshape(m, wgs))
    finally:
        f.release()


# ------------------------------------------------------------------------- statistical end-to-end --
def test_beta_binomial_posterior_matches_analytic(factory):
    """configs[1]: Beta(3,2) prior, N=50, z=15 -> Beta(18,37) (nvidia_gtx_test.clj:135-151) via mix!."""
    from scipy import stats
    model = models.beta_binomial_posterior()
    params = np.concatenate([models.binomial_lik_params(50, 15), models.beta_params(3, 2)])
    sf = factory.mcmc_factory(model)
    s = sf.create_sampler(1234, 2 ** 16, params)
    s.init_position(4321, f32([0, 1]))
    tuned = mcmc.mix(s)
    assert 0.2 <= tuned["acc-rate"] <= 0.75
    x = s.sample()[:, 0].astype(np.float64)
    post = stats.beta(18, 37)
    assert abs(x.mean() - post.mean()) < 0.01 * post.mean()           # T/library_test.clj: within 1 %
    assert abs(x.std() - post.std()) < 0.02 * post.std()
    assert stats.kstest(x[::16], post.cdf).statistic < 0.03
    h = s.histogram(4)
    centers = h.limits[0, 0] + (np.arange(G.WGS) + 0.5) * (h.limits[0, 1] - h.limits[0, 0]) / G.WGS
    assert np.abs(h.pdf[0] - post.pdf(centers)).max() < 0.35          # density-normalised histogram vs analytic pdf


def test_mvn_moments_after_burn_in(factory):
    d = 8
    model = models.mvn_model(d)
    params, mu, sigma = models.mvn_params(d, seed=5)
    sf = factory.mcmc_factory(model)
    s = sf.create_sampler(42, 2 ** 15, params)
    s.init_position(43, f32([[-30, 30]] * d))
    mcmc.mix(s)
    s.burn_in(300, 2.0)
    x = s.sample().astype(np.float64)
    assert np.abs(x.mean(axis=0) - mu).max() < 0.25
    assert np.allclose(np.cov(x.T), sigma, rtol=0.15, atol=0.5)


# ----------------------------------------------------------------------------------- errors / misc --
def test_walker_count_error(factory):
    sf = factory.mcmc_factory(models.GAUSSIAN)
    with pytest.raises(bb.WalkerCountError, match=r"Number of walkers \(300\) must be a multiple of 512\."):
        sf.create_sampler(1, 300, f32([0, 1]))


def test_state_roundtrip_and_init_position_from(factory):
    sf, gpu, cpu = make_pair(factory, models.GAUSSIAN, 1, 1024, f32([0, 1]), f32([-3, 3]))
    gpu.burn_in(10)
    st = gpu.get_state()
    other = sf.create_sampler(1, 1024, f32([0, 1]))
    other.init_position(gpu)
    st2 = other.get_state()
    assert np.array_equal(st["xs"], st2["xs"]) and logpdf_close(st["logfn"], st2["logfn"]).all()
    other.set_state(bare_seed=st["bare_seed"], move_seed=st["move_seed"], bare_counter=st["bare_counter"])
    gpu.burn_in(5)
    other.burn_in(5)
    assert np.array_equal(gpu.get_state()["xs"], other.get_state()["xs"])


def test_sample_small_n_and_processing_elements(factory):
    sf, gpu, cpu = make_pair(factory, models.UNIFORM, 9, 1024, f32([-1, 2]), f32([-1, 2]))
    assert np.array_equal(gpu.sample(10), cpu.sample(10))
    assert factory.processing_elements() % G.WGS == 0 and factory.processing_elements() >= 100 * G.WGS
    assert bb.launch_count() > 0


# ---- implementation variants must not change the chain ----------------------------------------------------------
def _chain_state(factory, model, params, limits, walkers):
    s = factory.mcmc_factory(model).create_sampler(21, walkers, params).init_position(22, limits)
    s.burn_in(19, 1.8)
    s.anneal(mcmc.minus_n(7.0), 7, 2.2)
    s.sample(walkers)                      # one more bare move
    st = s.get_state()
    r = s.run_sampler(65, 2.0)             # odd: exercises the double-buffered block sums of the run-sampler! loop
    return st, r["acceptance-rate"], s.accu_blocks(), s.last_means(65), r["autocorrelation"].tau, s.get_state()


VARIANT_CASES = [("gaussian-d1", lambda: (models.GAUSSIAN, f32([3, 1]), f32([-7, 7]), 4096)),
                 ("touch-d30", lambda: (models.therapeutic_touch_model(), models.therapeutic_touch_data(),
                                        models.therapeutic_touch_model().limits_array(), 2048)),
                 ("mvn-d100", lambda: (models.mvn_model(100), models.mvn_params(100)[0],
                                       models.mvn_model(100).limits_array(), 1024))]


@pytest.mark.parametrize("name,case", VARIANT_CASES, ids=[c[0] for c in VARIANT_CASES])
def test_persistent_loop_and_mirror_do_not_change_the_chain(factory, name, case, monkeypatch):
    """The persistent step loop (one cooperative launch for n steps), the AoS mirror of the ensemble and the
    constant-memory copy of the parameters are pure implementation choices: chains must be bit-identical with any of
    them switched off (BAY_LOOP=0, BAY_MIRROR=0, BAY_CPARAMS=0)."""
    model, params, limits, walkers = case()
    # a property of the generic program: quadratic-form models are kept off their tensor-core kernel here (that one is
    # held against the generic kernel and the oracle in test_quadform_tensor_core_move_*)
    monkeypatch.setenv("BAY_QUADFORM_TC", "0")
    base = _chain_state(factory, model, params, limits, walkers)
    for var in ("BAY_LOOP", "BAY_MIRROR", "BAY_CPARAMS"):
        monkeypatch.setenv(var, "0")
        other = _chain_state(factory, model, params, limits, walkers)     # the model is recompiled per sampler factory
        monkeypatch.delenv(var)
        assert np.array_equal(base[0]["xs"], other[0]["xs"]), (name, var)
        assert np.array_equal(base[0]["logfn"], other[0]["logfn"], equal_nan=True), (name, var)
        assert base[1] == other[1], (name, var)
        assert np.array_equal(base[2][0], other[2][0]) and np.array_equal(base[2][1], other[2][1]), (name, var)
        assert np.array_equal(base[3], other[3]), (name, var)                      # per-step ensemble means
        assert np.array_equal(base[4], other[4], equal_nan=True), (name, var)      # autocorrelation times
        assert np.array_equal(base[5]["xs"], other[5]["xs"]), (name, var)          # state after run-sampler!


# ---- quadratic-form models on the tensor cores (k_quadform_move_tc) ---------------------------------------------
@pytest.mark.parametrize("d,walkers", [(100, 2048), (8, 1024), (64, 1536), (128, 1024), (36, 4096 + 512)])
def test_quadform_tensor_core_move_vs_oracle(factory, d, walkers):
    """Step-locked against the oracle's serial LOGFN: proposals bit-identical, accepted log-densities within the
    north-star's 1e-5 relative (measured ~3e-7: fp16 hi/lo split with power-of-two row scales), accept masks equal
    except at near-ties.  Dimensions exercise one / two K chunks, N padding (36 -> 48) and ragged last tiles."""
    model = models.mvn_model(d)
    params, _, _ = models.mvn_params(d)
    sf, gpu, cpu = make_pair(factory, model, 3, walkers, params, model.limits_array())
    assert sf.uses_quadform()
    # band: the accept ratio uses the approximate __powf / __expf units like the reference's -use_fast_math build;
    # z^(D-1) = exp2((D-1) lg2 z) carries ~(D-1) * 2^-22 relative
    check_steplocked(gpu, cpu, steps=2, a=1.2, tie_tol=2e-3)
    # a few unlocked steps, then the log-densities the chain carries must be the model's at the positions it holds
    gpu.burn_in(5, 1.3)
    st = gpu.get_state()
    cpu.set_positions(st["xs"].reshape(-1))
    assert logpdf_close(st["logfn"], cpu.lp, rtol=1e-5).all()


def test_quadform_at_config5_size_sampled_against_the_oracle(factory):
    """BASELINE config 5 at its own size — 2^20 walkers, D = 100 — after tensor-core moves: the log-densities the chain
    carries, for a random sample of 8192 walkers, against the oracle's serial LOGFN at the positions it holds
    (north-star tolerance 1e-5 relative), and every position finite."""
    model = models.mvn_model(100)
    params, _, _ = models.mvn_params(100)
    W = 2 ** 20
    gpu = factory.mcmc_factory(model).create_sampler(31, W, params).init_position(32, model.limits_array())
    assert gpu.uses_quadform()
    gpu.burn_in(3, 1.25)
    st = gpu.get_state()
    assert np.all(np.isfinite(st["xs"])) and np.all(np.isfinite(st["logfn"]))
    pick = np.random.default_rng(0).choice(W, 8192, replace=False)
    cpu = orc.OracleStretch(model, 1, 8192, params, wgs=G.WGS)
    cpu.set_positions(st["xs"][pick].reshape(-1))
    assert logpdf_close(st["logfn"][pick], cpu.lp, rtol=1e-5).all()
    rate = gpu.acc_rate(1.25)
    assert 0.2 < rate < 0.8, rate
    gpu.release()


def test_quadform_tensor_core_equals_generic_kernel(factory, monkeypatch):
    """Same seeds, same proposals: the tensor-core move and the generic per-thread kernel may only part ways at
    near-ties of the accept test; where they agree the positions are bit-identical."""
    model = models.mvn_model(100)
    params, mu, sigma = models.mvn_params(100)
    lim = model.limits_array()
    tc = factory.mcmc_factory(model).create_sampler(8, 8192, params).init_position(9, lim)
    monkeypatch.setenv("BAY_QUADFORM_TC", "0")
    sfg = factory.mcmc_factory(model)
    monkeypatch.delenv("BAY_QUADFORM_TC")
    assert not sfg.uses_quadform()
    gen = sfg.create_sampler(8, 8192, params).init_position(9, lim)
    for s in (tc, gen):
        s.burn_in(3, 1.25)
    a, b = tc.get_state(), gen.get_state()
    same = np.all(a["xs"] == b["xs"], axis=1)
    assert same.mean() > 0.995, same.mean()
    assert logpdf_close(a["logfn"][same], b["logfn"][same], rtol=1e-5).all()
    # statistics of a longer run: the posterior mean and covariance the chain is supposed to sample
    tc.burn_in(400, 1.25)
    x = tc.sample().astype(np.float64)
    # every walker has moved and the ensemble contracts from the +-30 box towards N(mu, Sigma) (slow in D = 100)
    assert np.abs(x - mu).mean() < 12.0                  # started at mean |x - mu| = 15
    assert np.all(np.isfinite(tc.get_state()["logfn"]))
    rate = tc.acc_rate(1.25)
    assert 0.2 < rate < 0.7, rate


def test_constant_parameter_block_follows_the_sampler(factory):
    """Samplers of ONE compiled model share its __constant__ parameter block: interleaving two samplers with
    different parameters must give each the chain it has alone."""
    model = models.mvn_model(8)
    pa, _, _ = models.mvn_params(8, seed=1)
    pb, _, _ = models.mvn_params(8, seed=2)
    assert pa.size >= 64 and not np.array_equal(pa, pb)
    lim = model.limits_array()

    def alone(params):
        s = factory.mcmc_factory(model).create_sampler(5, 1024, params).init_position(6, lim)
        s.burn_in(6, 2.0)
        s.burn_in(5, 2.0)
        return s.get_state()

    sf = factory.mcmc_factory(model)
    a = sf.create_sampler(5, 1024, pa).init_position(6, lim)
    b = sf.create_sampler(5, 1024, pb).init_position(6, lim)
    a.burn_in(6, 2.0); b.burn_in(6, 2.0); a.burn_in(5, 2.0); b.burn_in(5, 2.0)
    for got, want in ((a.get_state(), alone(pa)), (b.get_state(), alone(pb))):
        assert np.array_equal(got["xs"], want["xs"]) and np.array_equal(got["logfn"], want["logfn"])
    assert sf.kernel_info("bay_stretch_bare")["registers"] > 0


def test_large_ensemble_bit_exact_and_index_arithmetic(factory):
    """Maximum-size edge: 2^24 walkers (the persistent loop does not fit, grids of 32 768 CTAs, 64 MB state) must
    still follow the oracle bit for bit, and a 2^21 x 12-D ensemble (offsets beyond 2^24 elements per half) must
    keep exact moments bookkeeping."""
    walkers = 1 << 24
    _, gpu, cpu = make_pair(factory, models.UNIFORM, 77, walkers, f32([-1, 2]), f32([-1, 2]))
    for s in (gpu, cpu):
        s.burn_in(2, 2.0)
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs)
    assert logpdf_close(st["logfn"], cpu.lp).all()          # fast-math log on the device
    h = gpu.histogram(1)
    counts = gpu.histogram_counts()
    assert int(counts.sum()) == walkers
    limits = orc.min_max(cpu.xs.reshape(walkers, 1), 1, walkers)
    assert np.array_equal(h.limits.reshape(-1), limits)
    assert np.array_equal(counts.reshape(-1), orc.histogram_counts(cpu.xs.reshape(walkers, 1), 1, walkers, G.WGS, limits))
    del gpu, cpu

    d, walkers = 12, 1 << 21
    model = models.mvn_model(d)
    params, mu, _ = models.mvn_params(d)
    s = factory.mcmc_factory(model).create_sampler(3, walkers, params).init_position(4, model.limits_array())
    s.burn_in(3, 2.0)
    x = s.sample(walkers)
    assert x.shape == (walkers, d) and np.isfinite(x).all()
    assert np.allclose(s.mean(), x.astype(np.float64).mean(axis=0), rtol=0, atol=2e-4)
