"""The oracle against the REFERENCE'S OWN KERNEL TEXT executed on the host (CPU only).

oracle/_ref/libref_*.so are gcc builds of /root/reference/src/device/.../opencl/engines/amd-gcn-mcmc-stretch.cl (plus
a distribution file) behind a small OpenCL-C shim — `make -C oracle ref`, run by __graft_entry__.build() wherever
/root/reference exists; the built files travel to the GPU box, the sources never enter the repo.  This is the
strongest pin of the restatement in oracle/bayadera_oracle.c: the real `stretch_move_bare`, `logfn` and
`init_walkers` run step by step beside it.
"""
import numpy as np
import pytest

import goldens as G
from bayadera_b200 import models
from oracle import oracle as orc
from oracle import ref_text

pytestmark = pytest.mark.skipif(not all(ref_text.available(s) for s in ref_text.BUILT),
                                reason="oracle/_ref not built (no /root/reference on this machine)")


def f32(x):
    return np.asarray(x, dtype=np.float32)


def test_reference_text_reproduces_its_own_goldens():
    """nvidia_gtx_test.clj:200-212 / amd_gcn_test.clj:198-208 through the reference's kernels compiled here."""
    r = ref_text.ReferenceTextStretch("uniform_d1", models.UNIFORM, G.SEED, G.W, f32([-1, 2]), wgs=G.WGS)
    r.init(G.SEED)
    r.init_position(G.SEED, f32([-1, 2]))
    for want in G.UNIFORM_SAMPLES:
        assert np.array_equal(r.sample()[:4, 0], f32(want))


@pytest.mark.parametrize("stem,model,params,limits", [
    ("uniform_d1", models.UNIFORM, [-1, 2], [-1, 2]),
    ("gaussian_d1", models.GAUSSIAN, [3, 1], [-7, 7]),
])
def test_oracle_equals_reference_text_bit_for_bit(stem, model, params, limits):
    """57 steps (burn-in at a = 2, annealed steps at a = 1.5 with beta != 1) from the same start: every position and
    log-density identical.  init_walkers is compared separately: `u*hi + (1-u)*lo` may be contracted either way by
    a compiler (gcc picks the other FMA than the reference's device compiler, whose choice the goldens pin and the
    oracle follows), so it agrees to 2 ulp, not bit for bit, for general limits."""
    a = ref_text.ReferenceTextStretch(stem, model, 5, 4096, f32(params))
    b = orc.OracleStretch(model, 5, 4096, f32(params))
    a.init_position(6, f32(limits))
    b.init_position(6, f32(limits))
    ulp = np.abs(a.xs.view(np.int32).astype(np.int64) - b.xs.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
    a.set_positions(b.xs.copy())                      # logfn kernel of the reference text on the same points
    assert np.array_equal(a.lp, b.lp, equal_nan=True)
    for s in (a, b):
        s.burn_in(50, 2.0)
        s.anneal(lambda i: 8.0 - i, 7, 1.5)
    assert np.array_equal(a.xs, b.xs)
    assert np.array_equal(a.lp, b.lp, equal_nan=True)


def test_gaussian_goldens_from_the_oracle_start():
    """The Gaussian(3,1) goldens (nvidia_gtx_test.clj:272-277) come out of the reference text bit-exactly when it
    starts from the oracle's init_walkers — i.e. the oracle's FMA choice there is the device compiler's."""
    o = orc.OracleStretch(models.GAUSSIAN, G.SEED, G.W, f32([3, 1.0]), wgs=G.WGS)
    o.init(G.SEED)
    o.init_position(G.SEED, f32([-7, 7]))
    r = ref_text.ReferenceTextStretch("gaussian_d1", models.GAUSSIAN, G.SEED, G.W, f32([3, 1.0]), wgs=G.WGS)
    r.init(G.SEED)
    r.set_positions(o.xs.copy())
    for want in G.GAUSSIAN_SAMPLES:
        assert np.array_equal(r.sample()[:4, 0], f32(want))


def test_literal_partner_switch_is_the_reference_text_for_dim_2():
    """for DIM > 1 the reference's partner is an element offset.  The oracle's `literal_partner`
    mode must BE that text, step-locked, except where the window runs past the half (the reference reads out of
    bounds there; the oracle clamps): expected rate 1/(2K) per walker and half-step."""
    src = ("extern \"C\" {\n#include <stdint.h>\n"
           "    inline REAL gaussian2_logpdf(const uint32_t data_len, const uint32_t params_len, const REAL* params,\n"
           "                                 const uint32_t dim, const REAL* x) {\n"
           "        const REAL a = x[0] - params[0];\n        const REAL b = x[1] - params[1];\n"
           "        return (a * a + b * b) / (-2.0f * params[2] * params[2]);\n    }\n}\n")
    m2 = models.DeviceModel("gaussian2", (src,), "gaussian2_logpdf", 2, 3, f32([[-5, 5], [-5, 5]]), "gaussian2_logpdf")
    p2 = f32([1, -1, 1.5])
    a = ref_text.ReferenceTextStretch("gaussian2_d2", m2, 5, 4096, p2)
    b = orc.OracleStretch(m2, 5, 4096, p2, literal_partner=True)
    b.init_position(6, f32([-5, 5, -5, 5]))
    a.set_positions(b.xs.copy(), b.lp.copy())
    total = differing = 0
    for _ in range(20):
        for half in (0, 1):
            a.half_bare(half)
            b.half_bare(half)
            act = slice(half * a.H * 2, (half + 1) * a.H * 2)
            differing += int((a.xs[act].reshape(-1, 2) != b.xs[act].reshape(-1, 2)).any(axis=1).sum())
            total += a.H
            b.xs[:] = a.xs                               # stay step-locked
            b.lp[:] = a.lp
        a.bare_counter += 1
        b.bare_counter += 1
    expected = total / (2 * a.H)
    assert differing <= 3 * expected + 5, (differing, expected)
    # ... and the walker-aligned default deliberately is NOT the reference text for DIM > 1
    c = orc.OracleStretch(m2, 5, 4096, p2)
    c.set_positions(a.xs.copy(), a.lp.copy())
    c.bare_counter = a.bare_counter
    before = a.xs.copy()
    a.half_bare(0)
    c.half_bare(0)
    moved = (a.xs[:a.H * 2] != before[:a.H * 2]).reshape(-1, 2).any(axis=1)
    assert (a.xs[:a.H * 2].reshape(-1, 2)[moved] != c.xs[:a.H * 2].reshape(-1, 2)[moved]).any(axis=1).mean() > 0.3


# ---- the model library (row a6) against the reference's distribution files -----------------------------------
def _ref_logfn(stem, params, points, params_size):
    lib = ref_text.load(stem)
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1)
    params = f32(params).reshape(-1)
    out = np.zeros(pts.size, dtype=np.float32)
    lib.ref_logfn(max(0, params.size - params_size), params_size, params, pts, out, pts.size)
    return out


def _our_logfn(model, fn_name, params, points):
    import dataclasses
    m = dataclasses.replace(model, mcmc_logpdf=fn_name, name=f"{model.name}_{fn_name}_vs_ref")
    fn, _keep = orc.compile_model(m)
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1)
    params = f32(params).reshape(-1)
    out = np.zeros(pts.size, dtype=np.float32)
    orc.lib().orc_logfn(fn, pts.size, 1, max(0, params.size - model.params_size), model.params_size, params, pts, out)
    return out


_MODEL_CASES = [
    ("uniform_logpdf", models.UNIFORM, [-1.0, 2.0], (-1.5, 2.5)),
    ("gaussian_logpdf", models.GAUSSIAN, [1.5, 0.7], (-3, 5)),
    ("student_t_logpdf", models.STUDENT_T, [4.0, 0.5, 2.0, -1.67], (-9, 9)),
    ("beta_logpdf", models.BETA, list(models.beta_params(2.5, 4.0)), (0.01, 0.99)),
    ("exponential_logpdf", models.EXPONENTIAL, [3.0, 1.0986123], (-0.5, 4)),
    ("erlang_logpdf", models.ERLANG, [2.0, 3.0, 1.3862944], (0.05, 6)),
    ("gamma_logpdf", models.GAMMA, [1.7, 2.4, -1.49], (0.05, 9)),
    ("binomial_logpdf", models.BINOMIAL, [20.0, 0.3], (0, 20)),
]


@pytest.mark.parametrize("fn,model,params,span", _MODEL_CASES, ids=[c[0] for c in _MODEL_CASES])
def test_model_strings_are_the_reference_arithmetic(fn, model, params, span):
    """Every built-in family: OUR model source (what NVRTC compiles) and the REFERENCE'S distribution file, both
    compiled by gcc without FMA contraction and with the same libm, give bit-identical log-densities — the re-authored
    library kept the reference's arithmetic order."""
    stem = f"model_{fn}"
    if not ref_text.available(stem):
        pytest.skip("reference model text not built")
    xs = np.linspace(span[0], span[1], 257, dtype=np.float32)
    ours = _our_logfn(model, fn, params, xs)
    theirs = _ref_logfn(stem, params, xs, model.params_size)
    assert np.array_equal(ours, theirs, equal_nan=True), np.nanmax(np.abs(ours - theirs))


def test_posterior_template_and_chain_equal_reference_text():
    """beta-binomial posterior (BASELINE config 2): our posterior_model == the reference's beta.cl + binomial.cl under
    its posterior template — bit-identical log-densities without contraction; with the device-style contraction
    the reference text differs by <= 2 ulp in logfn and the CHAIN (every accept decision over 40 steps) is identical."""
    post = models.beta_binomial_posterior()
    params = f32(np.concatenate([models.binomial_lik_params(50, 15), models.beta_params(3, 2)]))
    xs = np.linspace(0.001, 0.999, 513, dtype=np.float32)
    if ref_text.available("model_beta_binomial_mcmc_logpdf"):
        ours = _our_logfn(post, post.mcmc_logpdf, params, xs)
        theirs = _ref_logfn("model_beta_binomial_mcmc_logpdf", params, xs, post.params_size)
        assert np.array_equal(ours, theirs)
    if not ref_text.available("beta_binomial_d1"):
        pytest.skip("reference stretch text for the posterior not built")
    a = ref_text.ReferenceTextStretch("beta_binomial_d1", post, 5, 4096, params)
    b = orc.OracleStretch(post, 5, 4096, params)
    b.init_position(6, f32([0, 1]))
    a.set_positions(b.xs.copy())
    ulp = np.abs(a.lp.view(np.int32).astype(np.int64) - b.lp.view(np.int32).astype(np.int64))
    assert ulp.max() <= 2
    for s in (a, b):
        s.burn_in(40, 2.0)
    assert np.array_equal(a.xs, b.xs)
