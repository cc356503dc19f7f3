"""Engine-against-engine self-checks that run wherever the engine runs (bench.py prints their verdicts as
``parity_check``; tests/multigpu_check.py asserts on them).  No oracle is involved: each check holds one path of
the engine against another path of the same engine that is known-good from the oracle tests —

* mode A (walker partition): an R-GPU chain must be BIT-IDENTICAL to the 1-GPU chain;
* mode B (row shards + all-reduce): replicas hold identical state, log-densities equal the unsharded sampler's to
  double rounding;
* GLM tensor-core path: the Δlogp of (current, proposed) pairs against the fp64 traversal of the same rows.

Every function returns ``{"ok": bool, ...numbers...}`` and never raises on a mismatch.
"""
from __future__ import annotations

import numpy as np

from . import mcmc, models


def f32(v):
    return np.asarray(v, dtype=np.float32)


def _all_equal_across_ranks(arr: np.ndarray) -> bool:
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def mode_a_bit_identity(multi, single, world: int, steps: int = 12) -> dict:
    """Partitioned vs 1-GPU chain of the D = 30 therapeutic-touch ensemble and of a D = 100 Gaussian: positions,
    log-densities, accept counts, block sums, step means, histogram counts — all bit for bit.  Collective."""
    wgs = multi.wgs
    touch, mvn = models.therapeutic_touch_model(), models.mvn_model(100)
    cases = [("touch-d30", touch, models.therapeutic_touch_data(), touch.limits_array(), 2 * wgs * world * 2, 2.0),
             ("mvn-d100", mvn, models.mvn_params(100)[0], mvn.limits_array(), 2 * wgs * world * 4, 1.25)]
    out = {"ok": True, "cases": []}
    for name, model, params, limits, walkers, a in cases:
        sa = multi.mcmc_factory(model).create_sampler(11, walkers, params).init_position(12, limits)
        sb = single.mcmc_factory(model).create_sampler(11, walkers, params).init_position(12, limits)
        for s in (sa, sb):
            s.burn_in(steps, a)
            s.anneal(mcmc.minus_n(4.0), 4, a)
        ta, tb = sa.get_state(), sb.get_state()
        same = bool(np.array_equal(ta["xs"], tb["xs"]) and np.array_equal(ta["logfn"], tb["logfn"], equal_nan=True))
        ra, rb = sa.run_sampler(50, a), sb.run_sampler(50, a)
        same &= ra["acceptance-rate"] == rb["acceptance-rate"]
        acc_a, sums_a = sa.accu_blocks()
        acc_b, sums_b = sb.accu_blocks()
        same &= bool(np.array_equal(acc_a, acc_b) and np.array_equal(sums_a, sums_b))
        same &= bool(np.array_equal(sa.last_means(50), sb.last_means(50)))
        sa.histogram(2), sb.histogram(2)
        same &= bool(np.array_equal(sa.histogram_counts(), sb.histogram_counts()))
        xa = sa.sample(walkers)
        same &= bool(np.array_equal(xa, sb.sample(walkers)))
        if world > 1:
            same &= _all_equal_across_ranks(xa)
        out["cases"].append({"model": name, "walkers": walkers, "bit_identical": bool(same),
                             "acceptance": ra["acceptance-rate"]})
        out["ok"] &= bool(same)
        sa.release()
        sb.release()
    return out


def mode_b_replicas(multi, single, rank: int, world: int, rows: int = 60_000, d: int = 64, walkers: int = 1024) -> dict:
    """Row-sharded GLM sampler against the unsharded one on the same rows.  Collective."""
    from .distributed import shard_rows
    rng = np.random.default_rng(2024)
    x = rng.standard_normal((rows, d)).astype(np.float32)
    theta = (rng.standard_normal(d) / np.sqrt(8)).astype(np.float32)
    y = (rng.random(rows) < 1 / (1 + np.exp(-(x @ theta)))).astype(np.float32)
    data = np.concatenate([y[:, None], x], axis=1)
    hyper = f32([1.0 / 200.0])
    glm = models.logistic_regression_model(d)
    b0, b1 = shard_rows(rows, world, rank)
    sharded = multi.mcmc_factory(glm).create_sampler(5, walkers, np.concatenate([data[b0:b1].reshape(-1), hyper]))
    whole = single.mcmc_factory(glm).create_sampler(5, walkers, np.concatenate([data.reshape(-1), hyper]))
    for s in (sharded, whole):
        s.init_position(6, glm.limits_array())
    xs_s, lp_s = sharded.get_state64()
    xs_w, lp_w = whole.get_state64()
    rel = float(np.abs(lp_s / lp_w - 1.0).max())
    ok = bool(np.array_equal(xs_s, xs_w)) and rel < 1e-7
    for s in (sharded, whole):
        s.burn_in(4, 1.5)
    xs_s, lp_s = sharded.get_state64()
    xs_w, _ = whole.get_state64()
    agree = float(np.all(xs_s == xs_w, axis=1).mean())
    ok &= agree > 0.995
    replicas = True
    if world > 1:
        replicas = _all_equal_across_ranks(xs_s) and _all_equal_across_ranks(lp_s)
    ok &= replicas
    sharded.release()
    whole.release()
    return {"ok": bool(ok), "logdensity_rel_vs_unsharded": rel, "chain_agreement": agree,
            "replicas_identical": bool(replicas), "rows": rows}


def mode_b_rowadd(multi, single, rank: int, world: int, n: int = 400_000, walkers: int = 1024) -> dict:
    """Generic row-additive posterior (gaussian_loglik x a 2-D prior) with its data rows sharded over the ranks against
    the unsharded sampler: the row-free part of the likelihood must use the GLOBAL row count.  Collective."""
    from .distributed import shard_rows
    rng = np.random.default_rng(7)
    data = (2.5 + 1.7 * rng.standard_normal(n)).astype(np.float32)
    hyper = f32([0.0, 5.0, 0.5])
    model = models.gaussian_mean_sd_posterior()
    b0, b1 = shard_rows(n, world, rank)
    b0, b1 = (b0 // 4) * 4, (b1 // 4) * 4 if rank < world - 1 else n      # shard starts stay 16-byte aligned
    sharded = multi.mcmc_factory(model).create_sampler(5, walkers, np.concatenate([data[b0:b1], hyper]))
    whole = single.mcmc_factory(model).create_sampler(5, walkers, np.concatenate([data, hyper]))
    for s in (sharded, whole):
        s.init_position(6, model.limits_array())
    xs_s, lp_s = sharded.get_state64()
    xs_w, lp_w = whole.get_state64()
    rel = float(np.abs(lp_s / lp_w - 1.0).max())
    ok = bool(np.array_equal(xs_s, xs_w)) and rel < 1e-6
    for s in (sharded, whole):
        s.burn_in(4, 2.0)
    xs_s, lp_s = sharded.get_state64()
    xs_w, _ = whole.get_state64()
    agree = float(np.all(xs_s == xs_w, axis=1).mean())
    ok &= agree > 0.99
    replicas = True
    if world > 1:
        replicas = _all_equal_across_ranks(xs_s) and _all_equal_across_ranks(lp_s)
    ok &= replicas
    sharded.release()
    whole.release()
    return {"ok": bool(ok), "logdensity_rel_vs_unsharded": rel, "chain_agreement": agree,
            "replicas_identical": bool(replicas), "data": n}


def glm_delta_logp(sampler, theta_center: np.ndarray, scale: float, pairs: int = 256, a: float = 1.2, seed: int = 99) -> dict:
    """Δ(sum of the log-partition) of stretch proposals between walkers scattered at `scale` around `theta_center`,
    by the sampler's own path (tensor cores when eligible) against the fp64 traversal (bay_glm_loglik_probe)."""
    d = sampler.DIM
    rng = np.random.default_rng(seed)
    cur = (theta_center[None, :] + scale * rng.standard_normal((pairs, d))).astype(np.float32)
    other = (theta_center[None, :] + scale * rng.standard_normal((pairs, d))).astype(np.float32)
    u = rng.random(pairs)
    z = (((a - 1.0) * u + 1.0) ** 2 / a).astype(np.float32)
    prop = (other + z[:, None] * (cur - other)).astype(np.float32)
    pts = np.concatenate([cur, prop])
    own, ref = sampler.glm_loglik_probe(pts, 0), sampler.glm_loglik_probe(pts, 2)
    d_own, d_ref = own[pairs:] - own[:pairs], ref[pairs:] - ref[:pairs]
    err = np.abs(d_own - d_ref)
    # an fp32 traversal of 10^7 rows carries ~2.5e-3 mean / 1e-2 max of summation noise itself (tests/test_gpu_glm.py)
    return {"ok": bool(err.mean() < 1e-2 and err.max() < 5e-2), "pairs": pairs, "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()),
            "mean_abs_dlogp": float(np.abs(d_ref).mean()),
            "level_rel_err": float(np.abs(own - ref).max() / np.abs(ref).max())}
