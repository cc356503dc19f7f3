"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Mode A (walker partition + all-gather): an R-GPU chain must be BIT-IDENTICAL to the 1-GPU chain — positions,
log-densities, accept counts, per-block sums, histogram counts.
Mode B (row-sharded GLM + all-reduce): every rank holds identical state; log-densities agree with the unsharded
sampler to double rounding, chains agree except at near-ties.
Prints MULTIGPU OK on rank 0.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import bayadera_b200 as bb  # noqa: E402
from bayadera_b200 import mcmc, models  # noqa: E402
from bayadera_b200.distributed import init_engine_comm, shard_rows  # noqa: E402


def f32(v):
    return np.asarray(v, dtype=np.float32)


def all_equal_across_ranks(arr: np.ndarray) -> bool:
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wgs = 256
    multi = bb.B200BayaderaFactory(device=local, wgs=wgs)
    init_engine_comm(multi, rank, world, torch.device("cuda", local))
    single = bb.B200BayaderaFactory(device=local, wgs=wgs)          # no communicator: the 1-GPU reference

    # ---------------- mode A: walker partition, bit-identical to one GPU ----------------
    touch = models.therapeutic_touch_model()
    cases = [
        ("uniform", models.UNIFORM, f32([-1, 2]), f32([-1, 2]), 2 * wgs * world * 4),
        ("gaussian", models.GAUSSIAN, f32([3, 1]), f32([-7, 7]), 2 * wgs * world * 8),
        ("touch-d30", touch, models.therapeutic_touch_data(), touch.limits_array(), 2 * wgs * world * 2),
    ]
    for name, model, params, limits, walkers in cases:
        a = multi.mcmc_factory(model).create_sampler(11, walkers, params).init_position(12, limits)
        b = single.mcmc_factory(model).create_sampler(11, walkers, params).init_position(12, limits)
        for s in (a, b):
            s.burn_in(15, 2.0)
            s.anneal(mcmc.minus_n(6.0), 6, 1.7)
        sa, sb = a.get_state(), b.get_state()
        assert np.array_equal(sa["xs"], sb["xs"]), f"{name}: positions differ from the 1-GPU chain"
        assert np.array_equal(sa["logfn"], sb["logfn"], equal_nan=True), f"{name}: log-densities differ"
        ra, rb = a.run_sampler(64, 2.0), b.run_sampler(64, 2.0)
        assert ra["acceptance-rate"] == rb["acceptance-rate"], name
        acc_a, sums_a = a.accu_blocks()
        acc_b, sums_b = b.accu_blocks()
        assert np.array_equal(acc_a, acc_b) and np.array_equal(sums_a, sums_b), name
        assert np.array_equal(a.last_means(64), b.last_means(64)), name
        assert np.array_equal(ra["autocorrelation"].tau, rb["autocorrelation"].tau, equal_nan=True), name
        ha, hb = a.histogram(3), b.histogram(3)
        assert np.array_equal(a.histogram_counts(), b.histogram_counts()), name
        assert np.array_equal(ha.limits, hb.limits) and np.array_equal(ha.pdf, hb.pdf), name
        xa, xb = a.sample(walkers + 100), b.sample(walkers + 100)
        assert np.array_equal(xa, xb), name
        assert all_equal_across_ranks(xa), f"{name}: ranks disagree"
        assert a.info() == b.info()
    # walker-count rule under a partition: multiple of 2*WGS*R
    try:
        multi.mcmc_factory(models.GAUSSIAN).create_sampler(1, 2 * wgs * world + 2 * wgs, f32([0, 1]))
        if world > 1 and (2 * wgs * world + 2 * wgs) % (2 * wgs * world) != 0:
            raise AssertionError("walker-count check missing")
    except bb.WalkerCountError:
        pass

    # ---------------- mode B: row-sharded GLM ----------------
    d, rows, walkers = 64, 40_000, 1024
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
    assert np.array_equal(xs_s, xs_w)
    # sharding changes which CTA sums which tiles (fp32 partial sums regroup): agreement to ~1e-8, bound 1e-7
    assert np.allclose(lp_s, lp_w, rtol=1e-7), np.abs(lp_s / lp_w - 1).max()
    assert all_equal_across_ranks(lp_s), "all-reduced log-densities must be bit-identical on every rank"
    for s in (sharded, whole):
        s.burn_in(4, 1.5)
    xs_s, lp_s = sharded.get_state64()
    xs_w, lp_w = whole.get_state64()
    same = np.all(xs_s == xs_w, axis=1)
    assert same.mean() > 0.995, same.mean()
    assert all_equal_across_ranks(xs_s) and all_equal_across_ranks(lp_s), "replicas diverged"
    r = sharded.run_sampler(64, 1.5)
    assert 0.0 < r["acceptance-rate"] < 1.0

    # the self-checks bench.py runs before it prints a multi-GPU line (bayadera_b200/selfcheck.py)
    from bayadera_b200 import selfcheck
    ra = selfcheck.mode_a_bit_identity(multi, single, world)
    assert ra["ok"], ra
    rb = selfcheck.mode_b_replicas(multi, single, rank, world)
    assert rb["ok"], rb
    rc = selfcheck.mode_b_rowadd(multi, single, rank, world)
    assert rc["ok"], rc

    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU OK world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
