"""The model library against closed forms — CPU only.

The SAME source strings NVRTC compiles for the GPU are compiled by g++ (oracle.compile_model) and evaluated
through the oracle's `logfn`; scipy.stats provides the closed forms.  The GPU side of the argument is
tests/test_gpu_parity.py::test_init_logfn_and_steplocked_moves (GPU == oracle for every model)."""
import dataclasses
import math

import numpy as np
import pytest
from scipy import stats

from bayadera_b200 import models
from oracle import oracle as orc


def f32(x):
    return np.asarray(x, dtype=np.float32)


def evaluate(model, fn_name, params, points, dimension=None, extra_source=""):
    m = dataclasses.replace(model, mcmc_logpdf=fn_name, source=model.source + ((extra_source,) if extra_source else ()),
                            dimension=dimension or model.dimension, name=f"{model.name}_{fn_name}")
    fn, _keep = orc.compile_model(m)
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, m.dimension)
    out = np.zeros(pts.shape[0], dtype=np.float32)
    params = f32(params).reshape(-1)
    data_len = max(0, params.size - model.params_size) if not extra_source else params.size
    orc.lib().orc_logfn(fn, pts.shape[0], m.dimension, data_len, model.params_size if not extra_source else 0,
                        params, pts.reshape(-1), out)
    return out.astype(np.float64)


def student_t_params(nu, mu, sigma):
    logscale = math.lgamma((nu + 1) / 2) - math.lgamma(nu / 2) - 0.5 * math.log(nu * math.pi) - math.log(sigma)
    return [nu, mu, sigma, logscale]


CASES = [
    ("uniform", models.UNIFORM, [-1.0, 2.0], np.linspace(-0.9, 1.9, 41), lambda x: stats.uniform(-1, 3).logpdf(x)),
    ("gaussian", models.GAUSSIAN, [1.5, 0.7], np.linspace(-3, 5, 41), lambda x: stats.norm(1.5, 0.7).logpdf(x)),
    ("student_t", models.STUDENT_T, student_t_params(4.0, 0.5, 2.0), np.linspace(-9, 9, 41),
     lambda x: stats.t(4.0, 0.5, 2.0).logpdf(x)),
    ("beta", models.BETA, models.beta_params(2.5, 4.0), np.linspace(0.02, 0.98, 41), lambda x: stats.beta(2.5, 4.0).logpdf(x)),
    ("exponential", models.EXPONENTIAL, [3.0, math.log(3.0)], np.linspace(0.01, 4, 41),
     lambda x: stats.expon(scale=1 / 3.0).logpdf(x)),
    ("erlang", models.ERLANG, [2.0, 3.0, 3 * math.log(2.0) - math.lgamma(3.0)], np.linspace(0.05, 6, 41),
     lambda x: stats.gamma(3.0, scale=1 / 2.0).logpdf(x)),
    ("gamma", models.GAMMA, [1.7, 2.4, -math.lgamma(2.4) - 2.4 * math.log(1.7)], np.linspace(0.05, 9, 41),
     lambda x: stats.gamma(2.4, scale=1.7).logpdf(x)),
    ("binomial", models.BINOMIAL, [20.0, 0.3], np.arange(0, 21, dtype=np.float64), lambda k: stats.binom(20, 0.3).logpmf(k)),
]


@pytest.mark.parametrize("name,model,params,xs,closed", CASES, ids=[c[0] for c in CASES])
def test_normalised_logpdf_matches_scipy(name, model, params, xs, closed):
    got = evaluate(model, model.logpdf, params, xs)
    want = closed(xs)
    assert np.allclose(got, want, rtol=2e-5, atol=2e-5), np.abs(got - want).max()


@pytest.mark.parametrize("name,model,params,xs,closed", CASES, ids=[c[0] for c in CASES])
def test_mcmc_logpdf_differs_from_logpdf_by_a_constant(name, model, params, xs, closed):
    """mcmc-logpdf may drop the normalisation (that is all the sampler needs): the difference must not depend on x."""
    a = evaluate(model, model.logpdf, params, xs)
    b = evaluate(model, model.mcmc_logpdf, params, xs)
    d = a - b
    if name == "binomial":
        # reference quirk kept on purpose (K/cuda/distributions/binomial.cu:17-20): its mcmc-logpdf drops the binomial
        # coefficient although that depends on k
        from scipy.special import gammaln
        d = d - (gammaln(21.0) - gammaln(xs + 1) - gammaln(20.0 - xs + 1))
    finite = np.isfinite(d)
    assert finite.sum() >= 20 and np.ptp(d[finite]) < 5e-5 * max(1.0, np.abs(a[finite]).max()), np.ptp(d[finite])


def _lik_wrapper(loglik):
    return ('extern "C" {\n    inline REAL lik_as_logpdf(const uint32_t data_len, const uint32_t params_len, const REAL* params,\n'
            "                              const uint32_t dim, const REAL* x) {\n"
            f"        return {loglik}(data_len, params, dim, x);\n    }}\n}}\n")


def test_likelihoods_match_scipy():
    rng = np.random.default_rng(5)
    data = rng.normal(1.0, 2.0, 57)
    pts = np.array([[1.0, 2.0], [0.3, 1.1], [2.2, 3.5]])
    got = evaluate(models.GAUSSIAN, "lik_as_logpdf", data, pts, dimension=2, extra_source=_lik_wrapper("gaussian_loglik"))
    want = [stats.norm(m, s).logpdf(data).sum() for m, s in pts]
    assert np.allclose(got, want, rtol=3e-5)
    pts3 = np.array([[5.0, 1.0, 2.0], [3.0, 0.4, 1.3]])
    got = evaluate(models.STUDENT_T, "lik_as_logpdf", data, pts3, dimension=3, extra_source=_lik_wrapper("student_t_loglik"))
    want = [stats.t(nu, m, s).logpdf(data).sum() for nu, m, s in pts3]
    assert np.allclose(got, want, rtol=3e-5)
    ps = np.array([[0.2], [0.5], [0.9]])
    got = evaluate(models.BINOMIAL, "lik_as_logpdf", [50.0, 15.0], ps, dimension=1, extra_source=_lik_wrapper("binomial_loglik"))
    want = [15 * math.log(p) + 35 * math.log(1 - p) for p in ps[:, 0]]          # the kernel without the binomial coefficient
    assert np.allclose(got, want, rtol=1e-5)
    # invalid parameters -> NaN (the reference's convention: such proposals are rejected by isfinite)
    bad = evaluate(models.GAUSSIAN, "lik_as_logpdf", data, np.array([[0.0, -1.0]]), dimension=2,
                   extra_source=_lik_wrapper("gaussian_loglik"))
    assert np.isnan(bad).all()


def test_posterior_template_adds_likelihood_and_normalised_prior():
    """device-posterior-model (models.clj:102-115): loglik(data) + prior logpdf(hyperparams)."""
    post = models.beta_binomial_posterior()
    params = np.concatenate([models.binomial_lik_params(50, 15), models.beta_params(3, 2)])
    ps = np.linspace(0.05, 0.95, 19)
    got = evaluate(post, post.mcmc_logpdf, params, ps)
    want = 15 * np.log(ps) + 35 * np.log(1 - ps) + stats.beta(3, 2).logpdf(ps)
    assert np.allclose(got, want, rtol=2e-5, atol=2e-5)
