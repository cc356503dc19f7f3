"""Device models: C source strings exposing ``logpdf`` / ``mcmc_logpdf`` / ``loglik``.

Mirror of the model layer the stretch engine consumes in the reference:
``DeviceDistributionModel`` / ``DeviceLikelihoodModel`` / ``device-posterior-model``
(/root/reference/src/clojure/uncomplicate/bayadera/internal/device/models.clj:46-115)
and the C-side contract of SURVEY.md Appendix C:

    REAL f(const uint32_t data_len, const uint32_t params_len, const REAL* params,
           const uint32_t dim, const REAL* x)                         # logpdf / mcmc_logpdf
    REAL f(const uint32_t data_len, const REAL* data, const uint32_t dim, const REAL* x)  # loglik

Sources are plain C in the CUDA dialect of the reference: no ``__device__`` (NVRTC
``-default-device``), ``REAL`` is a macro, out-of-support returns NaN or -inf.  The
same strings are compiled by NVRTC for sm_100a (product) and by g++ for the CPU
oracle (tests only) — one source of truth.

The built-in families keep the reference's function names, parameter vectors and
arithmetic order (K/cuda/distributions/*.cu, K = …/internal/device/cuda) so that
user code written against Bayadera's library keeps working; the text itself is
written for this engine.  Models for BASELINE.json configs 3-5 (not in the
reference tree) are authored here against the same contract.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

# flags understood by bay_model_compile (include/bayadera_b200.h)
FAST_MATH = 0x1
ROW_ADDITIVE = 0x2
GLM_LOGISTIC = 0x4
GLM_POISSON = 0x8
QUADFORM = 0x10

_SIG = ("const uint32_t data_len, const uint32_t params_len, const REAL* params, "
        "const uint32_t dim, const REAL* x")
_LIK_SIG = "const uint32_t data_len, const REAL* data, const uint32_t dim, const REAL* x"
# row decomposition of a likelihood (engine extension, ignored by the reference): one data row / the row-free rest
_ROW_SIG = "const REAL* row, const uint32_t dim, const REAL* x"
_ROWCONST_SIG = "const uint32_t n_rows, const uint32_t dim, const REAL* x"


def _fn(name: str, body: str, sig: str = _SIG) -> str:
    return f"    inline REAL {name}({sig}) {{\n        {body}\n    }}\n"


def _wrap(*parts: str) -> str:
    return 'extern "C" {\n#include <stdint.h>\n' + "\n".join(parts) + "}\n"


# --- uniform: params [a b]; K/cuda/distributions/uniform.cu -------------------------------
UNIFORM_SRC = _wrap(
    "    inline REAL uniform_density(const REAL lo, const REAL hi, const REAL x) {\n"
    "        const bool inside = (lo <= x) && (x <= hi);\n"
    "        return inside ? (1 / (hi - lo)) : 0.0f;\n"
    "    }\n",
    _fn("uniform_logpdf", "return log(uniform_density(params[0], params[1], x[0]));"),
)

# --- gaussian: params [mu sigma]; K/cuda/distributions/gaussian.cu -------------------------
GAUSSIAN_SRC = _wrap(
    "#ifndef BAY_LOG_SQRT_2PI\n#define BAY_LOG_SQRT_2PI 0.9189385332046727f\n#endif\n"
    "    inline REAL gaussian_kernel(const REAL mu, const REAL sigma, const REAL x) {\n"
    "        return (x - mu) * (x - mu) / (-2.0f * sigma * sigma);\n"
    "    }\n"
    "    inline REAL gaussian_norm(const REAL sigma) {\n"
    "        return - log(sigma) - BAY_LOG_SQRT_2PI;\n"
    "    }\n",
    _fn("gaussian_mcmc_logpdf", "return gaussian_kernel(params[0], params[1], x[0]);"),
    _fn("gaussian_logpdf", "return gaussian_kernel(params[0], params[1], x[0]) + gaussian_norm(params[1]);"),
    _fn("gaussian_loglik",
        "const REAL mu = x[0];\n"
        "        const REAL sigma = x[1];\n"
        "        if (!(0.0f < sigma)) return nanf(\"NaN\");\n"
        "        REAL acc = gaussian_kernel(mu, sigma, data[0]) + data_len * gaussian_norm(sigma);\n"
        "        for (uint32_t i = 1; i < data_len; i++) acc += gaussian_kernel(mu, sigma, data[i]);\n"
        "        return acc;", _LIK_SIG),
    # the same likelihood, stated as a sum over data rows for the engine's row-additive path (glm_program.inc)
    _fn("gaussian_rowlik", "return gaussian_kernel(x[0], x[1], row[0]);", _ROW_SIG),
    _fn("gaussian_rowlik_const",
        "return (0.0f < x[1]) ? n_rows * gaussian_norm(x[1]) : nanf(\"NaN\");", _ROWCONST_SIG),
)

# --- student-t: params [nu mu sigma logscale]; K/cuda/distributions/student-t.cu ------------
STUDENT_T_SRC = _wrap(
    "#ifndef BAY_LOG_SQRT_PI\n#define BAY_LOG_SQRT_PI 0.5723649429247f\n#endif\n"
    "    inline REAL student_t_kernel(const REAL nu, const REAL mu, const REAL sigma, const REAL x) {\n"
    "        const REAL t = (x - mu) / sigma;\n"
    "        return - (0.5f * (nu + 1.0f) * log(1.0f + t * t / nu));\n"
    "    }\n"
    "    inline REAL student_t_norm(const REAL nu, const REAL sigma) {\n"
    "        return lgamma(0.5f * (nu + 1.0f)) - lgamma(0.5f * nu)\n"
    "            - BAY_LOG_SQRT_PI - 0.5f * log(nu) - log(sigma);\n"
    "    }\n",
    _fn("student_t_mcmc_logpdf", "return student_t_kernel(params[0], params[1], params[2], x[0]);"),
    _fn("student_t_logpdf", "return student_t_kernel(params[0], params[1], params[2], x[0]) + params[3];"),
    _fn("student_t_loglik",
        "const REAL nu = x[0];\n"
        "        const REAL mu = x[1];\n"
        "        const REAL sigma = x[2];\n"
        "        if (!((0.0f < nu) && (0.0f < sigma))) return nanf(\"NaN\");\n"
        "        const REAL norm = student_t_norm(nu, sigma);\n"
        "        REAL acc = 0.0;\n"
        "        for (uint32_t i = 0; i < data_len; i++) acc += (student_t_kernel(nu, mu, sigma, data[i]) + norm);\n"
        "        return acc;", _LIK_SIG),
    _fn("student_t_rowlik", "return student_t_kernel(x[0], x[1], x[2], row[0]);", _ROW_SIG),
    _fn("student_t_rowlik_const",
        "return ((0.0f < x[0]) && (0.0f < x[2])) ? n_rows * student_t_norm(x[0], x[2]) : nanf(\"NaN\");",
        _ROWCONST_SIG),
)

# --- beta: params [a b -lbeta(a,b)]; K/cuda/distributions/beta.cu ---------------------------
BETA_SRC = _wrap(
    "    inline REAL beta_kernel(const REAL a, const REAL b, const REAL x) {\n"
    "        return (a - 1.0f) * log(x) + (b - 1.0f) * log(1 - x);\n"
    "    }\n",
    _fn("beta_mcmc_logpdf", "return beta_kernel(params[0], params[1], x[0]);"),
    _fn("beta_logpdf", "return beta_kernel(params[0], params[1], x[0]) + params[2];"),
)

# --- exponential: params [lambda log(lambda)]; K/cuda/distributions/exponential.cu ----------
EXPONENTIAL_SRC = _wrap(
    _fn("exponential_mcmc_logpdf", "return (0.0f < x[0]) ? (- params[0] * x[0]) : nanf(\"NaN\");"),
    _fn("exponential_logpdf", "return (0.0f < x[0]) ? (- params[0] * x[0]) + params[1] : nanf(\"NaN\");"),
)

# --- erlang: params [lambda k logscale]; K/cuda/distributions/erlang.cu ---------------------
ERLANG_SRC = _wrap(
    "    inline REAL erlang_kernel(const REAL lambda, const REAL k, const REAL x) {\n"
    "        return (k - 1) * log(x) - lambda * x;\n"
    "    }\n",
    _fn("erlang_mcmc_logpdf", "return erlang_kernel(params[0], params[1], x[0]);"),
    _fn("erlang_logpdf", "return erlang_kernel(params[0], params[1], x[0]) + params[2];"),
)

# --- gamma: params [theta k logscale]; K/cuda/distributions/gamma.cu ------------------------
GAMMA_SRC = _wrap(
    "    inline REAL gamma_kernel(const REAL theta, const REAL k, const REAL x) {\n"
    "        return (k - 1.0f) * log(x) - (x / theta);\n"
    "    }\n",
    _fn("gamma_mcmc_logpdf", "return gamma_kernel(params[0], params[1], x[0]);"),
    _fn("gamma_logpdf", "return gamma_kernel(params[0], params[1], x[0]) + params[2];"),
)

# --- binomial: params [n p] / likelihood data [n k]; K/cuda/distributions/binomial.cu -------
BINOMIAL_SRC = _wrap(
    "    inline REAL binomial_kernel(const REAL n, const REAL p, const REAL k) {\n"
    "        return (k * log(p)) + ((n - k) * log(1 - p));\n"
    "    }\n"
    "    inline REAL binomial_lbinco(const REAL n, const REAL k) {\n"
    "        return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1);\n"
    "    }\n",
    _fn("binomial_mcmc_logpdf", "return binomial_kernel(params[0], params[1], x[0]);"),
    _fn("binomial_logpdf",
        "return binomial_kernel(params[0], params[1], x[0]) + binomial_lbinco(params[0], x[0]);"),
    _fn("binomial_loglik", "return binomial_kernel(data[0], x[0], data[1]);", _LIK_SIG),
)


def posterior_source(name: str, loglik: str, prior_logpdf: str) -> str:
    """``device-posterior-model`` template (models.clj:102-115, K/cuda/distributions/posterior.cu:5-10):
    loglik over params[0:data_len] plus the prior's *normalised* logpdf over params[data_len:]."""
    body = (f"return {loglik}(data_len, params, dim, x) +\n"
            f"            {prior_logpdf}(data_len, params_len, &params[data_len], dim, x);")
    return _wrap(_fn(f"{name}_logpdf", body), _fn(f"{name}_mcmc_logpdf", body))


def distribution_source(name: str, body: str) -> str:
    """``distribution.cu`` template: wrap a user-supplied body into the LOGFN signature."""
    return _wrap(_fn(name, body))


@dataclass(frozen=True)
class DeviceModel:
    """What crosses the boundary for one model (SURVEY §8b 'Inputs crossing it')."""
    name: str
    source: Tuple[str, ...]            # (source model)
    mcmc_logpdf: str                   # (mcmc-logpdf model)
    dimension: int = 1                 # (dimension model)
    params_size: int = 1               # (params-size model)
    limits: Optional[np.ndarray] = None  # 2 x DIM column-major (lo, hi)
    logpdf: Optional[str] = None
    loglik: Optional[str] = None
    flags: int = FAST_MATH             # reference compiles with -use_fast_math (nvidia_gtx.clj:630-633)
    meta: dict = field(default_factory=dict, compare=False, hash=False)

    def limits_array(self) -> np.ndarray:
        if self.limits is None:
            raise ValueError(f"model {self.name} has no default limits")
        return np.ascontiguousarray(np.asarray(self.limits, dtype=np.float32).reshape(-1))


def _lim(*pairs: Sequence[float]) -> np.ndarray:
    return np.asarray(pairs, dtype=np.float32).reshape(-1)


UNIFORM = DeviceModel("uniform", (UNIFORM_SRC,), "uniform_logpdf", 1, 2, None, "uniform_logpdf")
GAUSSIAN = DeviceModel("gaussian", (GAUSSIAN_SRC,), "gaussian_mcmc_logpdf", 1, 2, None, "gaussian_logpdf",
                       "gaussian_loglik", meta={"rowlik": ("gaussian_rowlik", "gaussian_rowlik_const", 1)})
STUDENT_T = DeviceModel("student_t", (STUDENT_T_SRC,), "student_t_mcmc_logpdf", 1, 4, None,
                        "student_t_logpdf", "student_t_loglik",
                        meta={"rowlik": ("student_t_rowlik", "student_t_rowlik_const", 1)})
BETA = DeviceModel("beta", (BETA_SRC,), "beta_mcmc_logpdf", 1, 3, _lim((0.0, 1.0)), "beta_logpdf")
EXPONENTIAL = DeviceModel("exponential", (EXPONENTIAL_SRC,), "exponential_mcmc_logpdf", 1, 2, None,
                          "exponential_logpdf")
ERLANG = DeviceModel("erlang", (ERLANG_SRC,), "erlang_mcmc_logpdf", 1, 3, None, "erlang_logpdf")
GAMMA = DeviceModel("gamma", (GAMMA_SRC,), "gamma_mcmc_logpdf", 1, 3, None, "gamma_logpdf")
BINOMIAL = DeviceModel("binomial", (BINOMIAL_SRC,), "binomial_mcmc_logpdf", 1, 2, None, "binomial_logpdf",
                       "binomial_loglik")

DISTRIBUTIONS = {m.name: m for m in (UNIFORM, GAUSSIAN, STUDENT_T, BETA, EXPONENTIAL, ERLANG, GAMMA, BINOMIAL)}


def posterior_model(prior: DeviceModel, name: str, likelihood: DeviceModel, row_additive: bool = True) -> DeviceModel:
    """``(posterior-model prior name likelihood)``: dimension, params-size and limits come from
    the prior; sources = distinct(prior ∪ likelihood) + the generated pair (models.clj:102-115)."""
    if likelihood.loglik is None or prior.logpdf is None:
        raise ValueError("posterior needs a likelihood with loglik and a prior with logpdf")
    srcs = list(dict.fromkeys(prior.source + likelihood.source))
    srcs.append(posterior_source(name, likelihood.loglik, prior.logpdf))
    flags = prior.flags
    rowlik = likelihood.meta.get("rowlik") if row_additive else None
    if rowlik:
        # the likelihood states its row decomposition: the engine streams the dataset ONCE per 128 walkers through
        # shared memory instead of once per walker (the serial loop above stays: it is the oracle and the
        # density-engine path).  Macros only — the reference's compiler would simply ignore them.
        fn, const_fn, stride = rowlik
        srcs.append(f"#define BAY_ROW_STRIDE {stride}\n#define BAY_ROWLIK {fn}\n#define BAY_ROWLIK_CONST {const_fn}\n"
                    f"#define BAY_PRIOR {prior.logpdf}\n")
        flags |= ROW_ADDITIVE
    return DeviceModel(name, tuple(srcs), f"{name}_mcmc_logpdf", prior.dimension, prior.params_size,
                       prior.limits, f"{name}_logpdf", likelihood.loglik, flags)


# ------------------------------------------------------------------------------------------
# host-side parameter vectors (C/distributions.clj of the reference: beta-params :543-545,
# binomial-lik-params :77-79, gaussian :411-414 …)
# ------------------------------------------------------------------------------------------
def _lgamma(x: float) -> float:
    import math
    return math.lgamma(x)


def beta_params(a: float, b: float) -> np.ndarray:
    return np.asarray([a, b, -(_lgamma(a) + _lgamma(b) - _lgamma(a + b))], dtype=np.float32)


def binomial_lik_params(n: float, k: float) -> np.ndarray:
    return np.asarray([n, k], dtype=np.float32)


# ==========================================================================================
# Models for BASELINE.json configs 3-5 (authored here; not in the reference tree)
# ==========================================================================================

def mean_sd_prior() -> DeviceModel:
    """A 2-D prior for the parameters (mu, sigma) of ``gaussian_loglik``: mu ~ N(m0, s0), sigma ~ Exponential(lam),
    independent.  params = [m0 s0 lam]."""
    src = _wrap(_fn("mean_sd_logpdf",
                    "const REAL t = (x[0] - params[0]) / params[1];\n"
                    "        if (!(0.0f < x[1])) return nanf(\"NaN\");\n"
                    "        return -0.5f * t * t - log(params[1]) - 0.9189385332046727f + log(params[2]) - params[2] * x[1];"))
    return DeviceModel("mean_sd", (src,), "mean_sd_logpdf", 2, 3, _lim((-10.0, 10.0), (0.1, 10.0)), "mean_sd_logpdf")


def gaussian_mean_sd_posterior(row_additive: bool = True) -> DeviceModel:
    """Posterior of (mu, sigma) given i.i.d. Gaussian data: the reference's ``gaussian_loglik``
    (K/cuda/distributions/gaussian.cu:36-46) under the posterior template, with the row-additive metadata."""
    return posterior_model(mean_sd_prior(), "gauss_ms", GAUSSIAN, row_additive=row_additive)


def nu_mean_sd_prior() -> DeviceModel:
    """A 3-D prior for (nu, mu, sigma) of ``student_t_loglik``: nu - 1 ~ Exponential(1/29) (Kruschke), mu ~ N(m0, s0),
    sigma ~ Exponential(lam).  params = [m0 s0 lam]."""
    src = _wrap(_fn("nu_mean_sd_logpdf",
                    "const REAL t = (x[1] - params[0]) / params[1];\n"
                    "        if (!((1.0f < x[0]) && (0.0f < x[2]))) return nanf(\"NaN\");\n"
                    "        return -(x[0] - 1.0f) / 29.0f - 0.5f * t * t - log(params[1]) + log(params[2]) - params[2] * x[2];"))
    return DeviceModel("nu_mean_sd", (src,), "nu_mean_sd_logpdf", 3, 3, _lim((2.0, 40.0), (-10.0, 10.0), (0.1, 10.0)),
                       "nu_mean_sd_logpdf")


def student_t_posterior(row_additive: bool = True) -> DeviceModel:
    """Robust location/scale posterior: ``student_t_loglik`` (K/cuda/distributions/student-t.cu:40-53)."""
    return posterior_model(nu_mean_sd_prior(), "robust", STUDENT_T, row_additive=row_additive)


def beta_binomial_posterior() -> DeviceModel:
    """config 2: binomial likelihood x beta prior (T/internal/nvidia_gtx_test.clj:135-151,
    T/library_test.clj:123-151).  params = [N z | a b -lbeta(a,b)]."""
    return posterior_model(BETA, "beta_binomial", BINOMIAL)


def therapeutic_touch_model(subjects: int = 28) -> DeviceModel:
    """config 3: hierarchical therapeutic-touch model (Kruschke DBDA ch.9).

    x = [theta_1..theta_S, omega, kappa-2]; data = [z_1..z_S];
    hyperparams = [N, a_omega, b_omega, shape_kappa, rate_kappa].
      z_s ~ Binomial(N, theta_s);  theta_s ~ Beta(omega*(kappa-2)+1, (1-omega)*(kappa-2)+1)
      omega ~ Beta(a_omega, b_omega);  kappa-2 ~ Gamma(shape, rate)
    """
    S = subjects
    body = f"""
        const REAL omega = x[{S}];
        const REAL km2 = x[{S + 1}];
        const REAL* hyper = &params[data_len];
        const REAL N = hyper[0];
        if (!((0.0f < omega) && (omega < 1.0f) && (0.0f < km2))) return nanf("NaN");
        const REAL a = omega * km2 + 1.0f;
        const REAL b = (1.0f - omega) * km2 + 1.0f;
        REAL acc = {S}.0f * (lgamma(a + b) - lgamma(a) - lgamma(b));
        for (uint32_t s = 0; s < {S}; s++) {{
            const REAL th = x[s];
            if (!((0.0f < th) && (th < 1.0f))) return nanf("NaN");
            const REAL z = params[s];
            acc += (z + a - 1.0f) * log(th) + (N - z + b - 1.0f) * log(1.0f - th);
        }}
        acc += (hyper[1] - 1.0f) * log(omega) + (hyper[2] - 1.0f) * log(1.0f - omega);
        acc += (hyper[3] - 1.0f) * log(km2) - hyper[4] * km2;
        return acc;"""
    src = distribution_source("touch_mcmc_logpdf", body)
    lim = [(0.01, 0.99)] * (S + 1) + [(0.1, 50.0)]
    return DeviceModel("touch", (src,), "touch_mcmc_logpdf", S + 2, 5, _lim(*lim), "touch_mcmc_logpdf",
                       meta={"subjects": S})


def therapeutic_touch_data(subjects: int = 28, trials: int = 10, seed: int = 9,
                           omega: float = 0.44, kappa: float = 12.0) -> np.ndarray:
    """Synthetic data of SURVEY §8d: params vector [z_1..z_S | N 1 1 .01 .01]."""
    rng = np.random.default_rng(seed)
    a = omega * (kappa - 2.0) + 1.0
    b = (1.0 - omega) * (kappa - 2.0) + 1.0
    theta = rng.beta(a, b, size=subjects)
    z = rng.binomial(trials, theta).astype(np.float32)
    return np.concatenate([z, np.asarray([trials, 1.0, 1.0, 0.01, 0.01], dtype=np.float32)])


def logistic_regression_model(d: int = 64, prior_sd: float = 10.0) -> DeviceModel:
    """config 4: Bayesian logistic regression.  data rows are [y, x_1..x_d] (row stride d+1);
    hyperparams = [1/(2*prior_sd^2)].  logpdf = sum_rows (y*eta - softplus(eta)) - sum theta^2/(2 sd^2).

    Row-additive + GLM metadata lets the engine run the tiled / tensor-core likelihood; the
    serial body below is the reference-style per-thread loop (and the oracle)."""
    prior_body = f"""
        REAL ss = 0.0f;
        for (uint32_t i = 0; i < {d}; i++) ss += x[i] * x[i];
        return - params[0] * ss;"""
    body = f"""
        const uint32_t stride = {d + 1};
        const uint32_t rows = data_len / stride;
        double acc = 0.0;   /* 1e7 rows: the sum is ~ -5e6, beyond fp32 resolution */
        for (uint32_t r = 0; r < rows; r++) {{
            const REAL* row = &params[(size_t)r * stride];
            REAL eta = 0.0f;
            for (uint32_t i = 0; i < {d}; i++) eta += row[1 + i] * x[i];
            const REAL sp = fmaxf(eta, 0.0f) + log(1.0f + exp(-fabsf(eta)));
            acc += (double)(row[0] * eta - sp);
        }}
        return (REAL)(acc + (double)logreg_prior(data_len, params_len, &params[data_len], dim, x));"""
    # BAY_GLM_PRIOR names the prior for the engine's row-additive GLM path (glm_program.inc)
    src = _wrap("#define BAY_GLM_PRIOR logreg_prior\n", _fn("logreg_prior", prior_body),
                _fn("logreg_mcmc_logpdf", body))
    return DeviceModel("logreg", (src,), "logreg_mcmc_logpdf", d, 1, _lim(*([(-1.0, 1.0)] * d)),
                       "logreg_mcmc_logpdf", flags=FAST_MATH | ROW_ADDITIVE | GLM_LOGISTIC,
                       meta={"row_stride": d + 1, "prior_sd": prior_sd})


def poisson_regression_model(d: int = 16, prior_sd: float = 10.0) -> DeviceModel:
    """Bayesian Poisson regression with log link: data rows are [y, x_1..x_d] (y a count), hyperparams =
    [1/(2*prior_sd^2)];  logpdf = sum_rows (y*eta - exp(eta)) - sum theta^2/(2 sd^2)  (the log y! constant is dropped).

    Same row-additive GLM interface as ``logistic_regression_model`` with the other log-partition function: the
    engine streams the dataset once for all walkers (tensor cores for d <= 128); the serial body below is the
    reference-style per-thread loop (and the oracle)."""
    prior_body = f"""
        REAL ss = 0.0f;
        for (uint32_t i = 0; i < {d}; i++) ss += x[i] * x[i];
        return - params[0] * ss;"""
    body = f"""
        const uint32_t stride = {d + 1};
        const uint32_t rows = data_len / stride;
        double acc = 0.0;
        for (uint32_t r = 0; r < rows; r++) {{
            const REAL* row = &params[(size_t)r * stride];
            REAL eta = 0.0f;
            for (uint32_t i = 0; i < {d}; i++) eta += row[1 + i] * x[i];
            acc += (double)(row[0] * eta - exp(eta));
        }}
        return (REAL)(acc + (double)poisreg_prior(data_len, params_len, &params[data_len], dim, x));"""
    src = _wrap("#define BAY_GLM_PRIOR poisreg_prior\n", _fn("poisreg_prior", prior_body),
                _fn("poisreg_mcmc_logpdf", body))
    return DeviceModel("poisreg", (src,), "poisreg_mcmc_logpdf", d, 1, _lim(*([(-0.5, 0.5)] * d)),
                       "poisreg_mcmc_logpdf", flags=FAST_MATH | ROW_ADDITIVE | GLM_POISSON,
                       meta={"row_stride": d + 1, "prior_sd": prior_sd})


def mvn_model(d: int = 100) -> DeviceModel:
    """config 5: d-dimensional correlated Gaussian N(mu, Sigma).

    params = [mu (d) | U (d x d row-major, upper triangular, zeros below the diagonal)] with Sigma^-1 = U^T U,
    so logpdf = -0.5 * |U (x - mu)|^2.  d must be a multiple of 4: rows of U and the centred point are read as
    aligned float4; only the blocks on or right of the diagonal are visited (~d(d+8)/2 FMAs instead of d^2).
    Eight rows of U at a time: each 4 centred coordinates feed 32 independent ``fmaf`` chains (explicit, so the GPU
    and the CPU oracle contract the same way)."""
    if d % 4:
        raise ValueError("mvn_model needs a dimension that is a multiple of 4")
    q = d // 4

    def rows_block(nr: int, first_row: str, first_col: str) -> str:
        accs = ", ".join(f"r{i} = 0.0f" for i in range(nr))
        loads = "\n".join(f"                const float4 u{i} = U0[{i} * {q} + j];" for i in range(nr))
        fmas = "\n".join(
            f"                r{i} = fmaf(u{i}.w, v.w, fmaf(u{i}.z, v.z, fmaf(u{i}.y, v.y, fmaf(u{i}.x, v.x, r{i}))));"
            for i in range(nr))
        norm = " + ".join(f"r{i} * r{i}" for i in range(nr))
        return f"""{{
            REAL {accs};
            const float4* U0 = &U[({first_row}) * {q}];
#pragma unroll 2
            for (uint32_t j = {first_col}; j < {q}; j++) {{
                const float4 v = c[j];
{loads}
{fmas}
            }}
            acc += {norm};
        }}"""

    full, rest = d // 8, d % 8
    body = f"""
        float4 c[{q}];
        for (uint32_t i = 0; i < {q}; i++) {{
            c[i].x = x[4 * i] - params[4 * i];
            c[i].y = x[4 * i + 1] - params[4 * i + 1];
            c[i].z = x[4 * i + 2] - params[4 * i + 2];
            c[i].w = x[4 * i + 3] - params[4 * i + 3];
        }}
        const float4* U = (const float4*)&params[{d}];
        REAL acc = 0.0f;
        for (uint32_t b = 0; b < {full}; b++) {rows_block(8, "8 * b", "2 * b")}
        {rows_block(rest, str(8 * full), str(2 * full)) if rest else ""}
        return -0.5f * acc;"""
    src = distribution_source(f"mvn{d}_mcmc_logpdf", body)
    # QUADFORM: the engine may evaluate -1/2 |U (x - mu)|^2 for 128-walker tiles on the tensor cores (d <= 128)
    return DeviceModel(f"mvn{d}", (src,), f"mvn{d}_mcmc_logpdf", d, d + d * d,
                       _lim(*([(-30.0, 30.0)] * d)), f"mvn{d}_mcmc_logpdf", flags=FAST_MATH | QUADFORM)


def mvn_params(d: int = 100, seed: int = 5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """mu_i = i/10, Sigma = Q diag(logspace(0,2,d)) Q^T, Q from the QR of a
    seeded Gaussian matrix.  Returns (params vector, mu, Sigma)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    ev = np.logspace(0.0, 2.0, d)
    sigma = (q * ev) @ q.T
    prec = (q / ev) @ q.T
    u = np.linalg.cholesky(prec).T          # prec = U^T U, U upper triangular
    mu = np.arange(d, dtype=np.float64) / 10.0
    return np.concatenate([mu, np.triu(u).reshape(-1)]).astype(np.float32), mu, sigma
