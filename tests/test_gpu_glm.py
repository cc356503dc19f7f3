"""GPU parity of the GLM (Bernoulli-logit) row-additive likelihood path — BASELINE config 4's hot path.

Oracle: the model's own serial LOGFN (double-accumulated row loop) run by the CPU oracle at sizes it finishes in
seconds; at larger sizes, size-independent properties (row-shard additivity, equality with the generic per-thread
path, permutation invariance over rows).  Tolerance (north_star): summed log-likelihood within 1e-5 relative.
"""
import dataclasses

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


def synth(rows, d, seed=2024):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((rows, d)).astype(np.float32)
    theta = (rng.standard_normal(d) / np.sqrt(8)).astype(np.float32)
    y = (rng.random(rows) < 1 / (1 + np.exp(-(x @ theta)))).astype(np.float32)
    data = np.concatenate([y[:, None], x], axis=1).reshape(-1)
    return np.concatenate([data, f32([1.0 / (2 * 10.0 ** 2)])]), theta


@pytest.mark.parametrize("rows,d,walkers", [(1, 4, 512), (33, 8, 512), (5000, 64, 1024), (20011, 64, 512)])
def test_glm_logdensity_matches_oracle(factory, rows, d, walkers):
    model = models.logistic_regression_model(d)
    params, _ = synth(rows, d)
    sf = factory.mcmc_factory(model)
    gpu = sf.create_sampler(5, walkers, params).init_position(6, model.limits_array())
    cpu = orc.OracleStretch(model, 5, walkers, params, wgs=WGS).init_position(6, model.limits_array())
    st = gpu.get_state()
    assert np.array_equal(st["xs"].reshape(-1), cpu.xs)
    assert logpdf_close(st["logfn"], cpu.lp, rtol=1e-5).all()


def test_glm_steplocked_vs_oracle(factory):
    model = models.logistic_regression_model(8)
    params, _ = synth(700, 8)
    sf = factory.mcmc_factory(model)
    gpu = sf.create_sampler(9, 1024, params).init_position(10, model.limits_array())
    cpu = orc.OracleStretch(model, 9, 1024, params, wgs=WGS).init_position(10, model.limits_array())
    check_steplocked(gpu, cpu, steps=3, a=2.0, tie_tol=5e-3, exact_lp=False)


def test_glm_equals_generic_path(factory):
    """The tiled path and the reference-style per-thread serial loop must give the same chain statistics and the
    same log-densities (1e-5 relative) — 'checked against the fp32 SIMT path'."""
    d, rows, walkers = 16, 3000, 1024
    glm = models.logistic_regression_model(d)
    generic = dataclasses.replace(glm, flags=models.FAST_MATH)
    params, theta = synth(rows, d)
    a = factory.mcmc_factory(glm).create_sampler(1, walkers, params).init_position(2, glm.limits_array())
    b = factory.mcmc_factory(generic).create_sampler(1, walkers, params).init_position(2, glm.limits_array())
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa["xs"], sb["xs"])
    assert logpdf_close(sa["logfn"], sb["logfn"], rtol=1e-5).all()
    a.burn_in(3, 2.0)
    b.burn_in(3, 2.0)
    sa, sb = a.get_state(), b.get_state()
    same = np.all(sa["xs"] == sb["xs"], axis=1)
    assert same.mean() > 0.995                       # only near-tie accept flips may differ
    assert logpdf_close(sa["logfn"][same], sb["logfn"][same], rtol=1e-5).all()


def test_glm_row_shard_additivity_and_permutation(factory):
    """Size-independent property at a larger size: log-likelihood over the rows = sum over row shards, and is
    invariant under a permutation of the rows (prior counted once)."""
    d, rows, walkers = 64, 200_000, 512
    model = models.logistic_regression_model(d)
    params, _ = synth(rows, d, seed=7)
    data, hyper = params[:-1].reshape(rows, d + 1), params[-1:]
    sf = factory.mcmc_factory(model)

    def logdens(block):
        p = np.concatenate([block.reshape(-1), hyper])
        s = sf.create_sampler(3, walkers, p).init_position(4, model.limits_array())
        out = s.get_state()
        s.release()
        return out["xs"], out["logfn"].astype(np.float64)

    xs, full = logdens(data)
    _, a = logdens(data[:70_001])
    _, b = logdens(data[70_001:])
    prior = -(float(hyper[0]) * (xs.astype(np.float64) ** 2).sum(axis=1))
    assert np.allclose(full, a + b - prior, rtol=1e-5)
    perm = np.random.default_rng(0).permutation(rows)
    _, shuffled = logdens(data[perm])
    assert np.allclose(full, shuffled, rtol=1e-6)


def test_glm_posterior_recovers_truth(factory):
    d, rows, walkers = 8, 20_000, 4096
    model = models.logistic_regression_model(d)
    params, theta = synth(rows, d, seed=11)
    s = factory.mcmc_factory(model).create_sampler(21, walkers, params).init_position(22, model.limits_array())
    mcmc.mix(s)
    s.burn_in(200, 2.0)
    res = s.run_sampler(64, 2.0)
    assert 0.1 < res["acceptance-rate"] < 0.8
    x = s.sample().astype(np.float64)
    # posterior sd ~ 2/sqrt(rows) per coefficient; the posterior mean must sit within a few sd of the truth
    assert np.abs(x.mean(axis=0) - theta).max() < 0.08
    assert x.std(axis=0).max() < 0.05
    assert np.all(np.isfinite(s.get_state()["logfn"]))


@pytest.mark.parametrize("rows,walkers,d", [(1024, 512, 64), (30_017, 1024, 64), (9_000, 1536, 64),
                                            (50_000, 2048, 64), (7_001, 512, 4), (12_345, 1024, 20),
                                            (20_000, 512, 48), (15_000, 512, 96), (9_999, 1024, 128), (4_500, 512, 72)])
def test_glm_tensor_core_path_matches_simt_and_oracle(factory, rows, walkers, d, monkeypatch):
    """DIM <= 128 runs the tcgen05 kernel (bf16 hi/lo split, fp32 accumulation in TMEM; DIM zero-padded to one or
    two 64-wide K chunks).  It must agree with the fp32 SIMT tiled kernel and with the oracle's serial model to 1e-5
    relative on the summed log-likelihood."""
    model = models.logistic_regression_model(d)
    params, _ = synth(rows, d, seed=rows)
    sf = factory.mcmc_factory(model)
    monkeypatch.setenv("BAY_GLM_TC", "1")
    tc = sf.create_sampler(5, walkers, params).init_position(6, model.limits_array())
    monkeypatch.setenv("BAY_GLM_TC", "0")
    simt = sf.create_sampler(5, walkers, params).init_position(6, model.limits_array())
    a, b = tc.get_state(), simt.get_state()
    assert np.array_equal(a["xs"], b["xs"])
    rel = np.abs(a["logfn"].astype(np.float64) - b["logfn"]) / np.abs(b["logfn"])
    assert rel.max() < 1e-5, rel.max()
    if rows <= 30_017:
        cpu = orc.OracleStretch(model, 5, walkers, params, wgs=WGS).init_position(6, model.limits_array())
        assert logpdf_close(a["logfn"], cpu.lp, rtol=1e-5).all()
    # a few moves: identical accept decisions except near-ties, so almost all walkers stay bit-identical
    tc.burn_in(2, 1.5)
    simt.burn_in(2, 1.5)
    a, b = tc.get_state(), simt.get_state()
    same = np.all(a["xs"] == b["xs"], axis=1)
    assert same.mean() > 0.99, same.mean()
    rel = np.abs(a["logfn"].astype(np.float64)[same] - b["logfn"][same]) / np.abs(b["logfn"][same])
    assert rel.max() < 1e-5


def test_glm_state64_roundtrip_continues_chain_bit_exactly(factory):
    """Checkpoint/resume through the C-ABI: positions + double log-densities + the Philox counters
    fully determine the chain — a restored sampler continues bit-for-bit."""
    d, rows, walkers = 64, 4096, 1024
    model = models.logistic_regression_model(d)
    params, _ = synth(rows, d, seed=5)
    sf = factory.mcmc_factory(model)
    a = sf.create_sampler(3, walkers, params).init_position(4, model.limits_array())
    b = sf.create_sampler(3, walkers, params).init_position(4, model.limits_array())
    a.burn_in(5, 1.5)
    b.burn_in(2, 1.5)
    xs, lp64 = b.get_state64()
    st = b.get_state()
    assert lp64.dtype == np.float64 and np.allclose(lp64, st["logfn"], rtol=1e-6)
    c = sf.create_sampler(99, walkers, params)           # different seed: everything must come from the state
    c.set_state64(xs, lp64)
    c.set_state(bare_seed=st["bare_seed"], move_seed=st["move_seed"], bare_counter=st["bare_counter"],
                move_counter=st["move_counter"])
    c.burn_in(3, 1.5)
    xa, la = a.get_state64()
    xc, lc = c.get_state64()
    assert np.array_equal(xa, xc) and np.array_equal(la, lc)


def synth_poisson(rows, d, seed=11):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, d)) / np.sqrt(d)).astype(np.float32)
    theta = (0.3 * rng.standard_normal(d)).astype(np.float32)
    y = rng.poisson(np.exp(x @ theta)).astype(np.float32)
    data = np.concatenate([y[:, None], x], axis=1).reshape(-1)
    return np.concatenate([data, f32([1.0 / (2 * 10.0 ** 2)])]), theta


@pytest.mark.parametrize("rows,walkers,d", [(700, 512, 8), (20_011, 1024, 16), (6_000, 512, 64), (5_000, 512, 100)])
def test_poisson_link_matches_oracle_and_simt(factory, rows, walkers, d, monkeypatch):
    """The log-link (Poisson) variant of the row-additive GLM path: tensor-core kernel (>= 1024 rows) and fp32 tiled
    kernel against the oracle's serial model, 1e-5 relative on the summed log-density; chains stay step-locked."""
    model = models.poisson_regression_model(d)
    params, _ = synth_poisson(rows, d)
    sf = factory.mcmc_factory(model)
    tc = sf.create_sampler(5, walkers, params).init_position(6, model.limits_array())
    monkeypatch.setenv("BAY_GLM_TC", "0")
    simt = sf.create_sampler(5, walkers, params).init_position(6, model.limits_array())
    monkeypatch.delenv("BAY_GLM_TC")
    cpu = orc.OracleStretch(model, 5, walkers, params, wgs=WGS).init_position(6, model.limits_array())
    a, b = tc.get_state(), simt.get_state()
    assert np.array_equal(a["xs"].reshape(-1), cpu.xs) and np.array_equal(a["xs"], b["xs"])
    assert logpdf_close(a["logfn"], cpu.lp, rtol=1e-5).all()
    assert logpdf_close(b["logfn"], cpu.lp, rtol=1e-5).all()
    tc.burn_in(2, 1.5)
    simt.burn_in(2, 1.5)
    a, b = tc.get_state(), simt.get_state()
    same = np.all(a["xs"] == b["xs"], axis=1)
    assert same.mean() > 0.99, same.mean()


def test_poisson_posterior_recovers_truth(factory):
    d, rows = 8, 50_000
    model = models.poisson_regression_model(d)
    params, theta = synth_poisson(rows, d, seed=3)
    s = factory.mcmc_factory(model).create_sampler(1, 2048, params).init_position(2, model.limits_array())
    s.burn_in(600, 2.0)
    x = s.sample().astype(np.float64)
    assert np.abs(x.mean(axis=0) - theta).max() < 0.08, np.abs(x.mean(axis=0) - theta).max()


# ---------------------------------------------------------------------------------------------------------------
# Parity at the size that is benchmarked: D = 64, 10^7 rows (BASELINE config 4).  Log-densities are ~ -5e6 there
# and the accept test consumes O(1) differences, so relative tolerances on the log-density say nothing; what is
# held against an fp64 evaluation of the SAME points is the Δlogp of (current, proposed) pairs and the accept mask.
# ---------------------------------------------------------------------------------------------------------------
def _device_rows(rows, d, seed):
    """bench.py's generator (X ~ N(0,1), theta* ~ N(0, 1/8), y ~ Bernoulli(sigmoid(X theta*))) on the device."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty(rows * (d + 1) + 1, dtype=torch.float32, device=dev)
    mat = out[:-1].view(rows, d + 1)
    mat[:, 1:].normal_(generator=g)
    theta = torch.randn(d, generator=g, device=dev) / (8.0 ** 0.5)
    p = torch.sigmoid(mat[:, 1:] @ theta)
    mat[:, 0] = (torch.rand(rows, generator=g, device=dev) < p).float()
    out[-1] = 1.0 / 200.0
    torch.cuda.synchronize()
    return out, theta.cpu().numpy().astype(np.float64)


def test_glm_delta_logp_and_accept_mask_at_bench_size(factory):
    d, rows, pairs = 64, 10 ** 7, 1024
    model = models.logistic_regression_model(d)
    data, theta = _device_rows(rows, d, 2024)
    sampler = factory.mcmc_factory(model).create_sampler(123, 2 * pairs, bb.DeviceParams.from_torch(data))
    # walkers at the posterior's own scale around the generating coefficients (sd ~ 1/sqrt(rows * 0.2) = 7e-4), the
    # regime the timed chain lives in; proposals are stretch moves between random pairs of them
    rng = np.random.default_rng(99)
    cur = (theta[None, :] + 7e-4 * rng.standard_normal((pairs, d))).astype(np.float32)
    other = (theta[None, :] + 7e-4 * rng.standard_normal((pairs, d))).astype(np.float32)
    a = 1.2
    u = rng.random(pairs)
    z = (((a - 1.0) * u + 1.0) ** 2 / a).astype(np.float32)              # g(z) ~ 1/sqrt(z) on [1/a, a]
    prop = (other + z[:, None] * (cur - other)).astype(np.float32)
    pts = np.concatenate([cur, prop])
    sums = {m: sampler.glm_loglik_probe(pts, m) for m in (0, 1, 2)}
    ref = sums[2][pairs:] - sums[2][:pairs]                                # fp64 Δ(sum softplus), proposed - current
    assert np.all(np.isfinite(ref)) and 0.05 < np.abs(ref).mean() < 1e4   # O(1)..O(100) differences of ~7e6 sums
    err_tc = np.abs((sums[0][pairs:] - sums[0][:pairs]) - ref)
    err_simt = np.abs((sums[1][pairs:] - sums[1][:pairs]) - ref)
    print(f"|dlogp_tc - dlogp_f64|: max {err_tc.max():.3e} mean {err_tc.mean():.3e};  "
          f"fp32 SIMT: max {err_simt.max():.3e} mean {err_simt.mean():.3e};  |dlogp| mean {np.abs(ref).mean():.3f}")
    # the fp32 SIMT traversal itself sits at ~2.5e-3 mean / 1e-2 max here (fp32 summation noise over 10^7 rows); the
    # tensor-core path must be of the same quality: 1e-2 on average, no outlier beyond a few times that
    assert err_tc.mean() < 1e-2 and err_tc.max() < 5e-2, (err_tc.mean(), err_tc.max())
    assert err_tc.mean() < 4.0 * err_simt.mean() + 1e-3, (err_tc.mean(), err_simt.mean())
    # the accept test z^(D-1) exp(Δlogp) >= u with Δlogp = sy.(θp - θc) - Δ(sum softplus) + Δprior: the first and last
    # terms are computed identically on both paths, so they are taken from fp64 host arithmetic here
    xf = data[:-1].view(rows, d + 1)
    sy = (xf[:, 1:].double() * xf[:, :1].double()).sum(dim=0).cpu().numpy()
    lin = (prop.astype(np.float64) - cur.astype(np.float64)) @ sy
    dprior = -(1.0 / 200.0) * ((prop.astype(np.float64) ** 2).sum(1) - (cur.astype(np.float64) ** 2).sum(1))
    uz = rng.random(pairs)

    def mask(dsp):
        return uz <= z.astype(np.float64) ** (d - 1) * np.exp(np.minimum(lin - dsp + dprior, 50.0))

    m64, mtc = mask(ref), mask(sums[0][pairs:] - sums[0][:pairs])
    assert 0.02 < m64.mean() < 0.98
    flips = int((m64 != mtc).sum())
    print(f"accept decisions that differ from the fp64 evaluation: {flips} of {pairs}")
    assert flips <= pairs // 100, flips                                  # only near-ties of the accept test may flip
    # absolute level: the summed log-partition itself, against fp64, well inside the north-star's 1e-5 relative
    lvl = np.abs(sums[0] - sums[2]).max() / np.abs(sums[2]).max()
    print(f"level error of the summed log-partition vs fp64: {lvl:.3e} relative")
    assert lvl < 1e-6
    sampler.release()
