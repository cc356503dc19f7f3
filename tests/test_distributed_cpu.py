"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, unique-id broadcast, and the
row-sharded likelihood protocol itself — partial log-likelihoods all-reduced, identical accept decisions on every
rank — executed with the CPU oracle in place of the CUDA kernels (SURVEY §8e mode B)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from bayadera_b200 import models
from bayadera_b200.distributed import broadcast_unique_id, shard_rows


def test_shard_rows_partition():
    for rows, world in [(10, 3), (10 ** 7, 8), (5, 8), (128, 2)]:
        cover = []
        for r in range(world):
            b, e = shard_rows(rows, world, r)
            assert 0 <= b <= e <= rows
            cover.extend(range(b, e)) if rows <= 1000 else None
            assert abs((e - b) - rows / world) < 1
        if rows <= 1000:
            assert cover == list(range(rows))
    assert shard_rows(10 ** 7, 8, 7)[1] == 10 ** 7
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc

        # 1. unique id broadcast: every rank ends up with rank 0's bytes
        uid = broadcast_unique_id(lambda: np.arange(128, dtype=np.uint8) * 2 % 251, rank)
        assert np.array_equal(uid, np.arange(128, dtype=np.uint8) * 2 % 251)

        # 2. row-sharded stretch moves with replicated walkers
        d, rows, walkers, wgs = 4, 600, 512, 256
        rng = np.random.default_rng(3)
        x = rng.standard_normal((rows, d)).astype(np.float32)
        theta = (rng.standard_normal(d) / 2).astype(np.float32)
        y = (rng.random(rows) < 1 / (1 + np.exp(-(x @ theta)))).astype(np.float32)
        data = np.concatenate([y[:, None], x], axis=1)
        hyper = np.asarray([1.0 / 200.0], dtype=np.float32)
        model = models.logistic_regression_model(d)
        b, e = shard_rows(rows, world, rank)
        zero_prior = np.asarray([0.0], dtype=np.float32)          # shards carry the likelihood only
        shard = orc.OracleStretch(model, 7, walkers, np.concatenate([data[b:e].reshape(-1), zero_prior]), wgs=wgs)
        prior_only = orc.OracleStretch(model, 7, walkers, hyper, wgs=wgs)          # no rows: prior term alone
        full = orc.OracleStretch(model, 7, walkers, np.concatenate([data.reshape(-1), hyper]), wgs=wgs)
        lim = model.limits_array()
        for s in (shard, prior_only, full):
            s.init_position(11, lim)

        def allreduced_logdensity(points):
            """what bay_glm does per half-step: local partial over the shard, all-reduce(sum), prior added once"""
            shard.set_positions(points)
            part = torch.from_numpy(shard.lp.astype(np.float64))
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
            prior_only.set_positions(points)
            return part.numpy() + prior_only.lp.astype(np.float64)

        lp = allreduced_logdensity(full.xs)
        assert np.allclose(lp, full.lp, rtol=2e-6, atol=1e-4)
        # every rank holds the same numbers bit-for-bit after the all-reduce -> identical accept decisions
        gathered = [torch.zeros(walkers, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(lp))
        assert all(torch.equal(gathered[0], g) for g in gathered)
        # three replicated moves driven by the all-reduced density reproduce the single-process chain
        full.burn_in(3, 2.0)
        ref_xs = full.xs.copy()
        chain = orc.OracleStretch(model, 7, walkers, np.concatenate([data.reshape(-1), hyper]), wgs=wgs)
        chain.init_position(11, lim)
        chain.set_positions(chain.xs, lp.astype(np.float32))
        chain.burn_in(3, 2.0)
        assert np.array_equal(chain.xs, ref_xs)
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_row_sharded_protocol_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["ok0", "ok1"]


def _partition_worker(rank, world, port, out_dir):
    """Mode A (walker partition): rank r updates only walkers [r*H/R, (r+1)*H/R) of the active half, the updated
    slices (positions + log-densities) are all-gathered, and the chain must equal the single-process chain bit for
    bit — the host protocol of engine.cu's `my_slice` / `exchange_half`, run over gloo with the oracle as the kernel."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        model = models.therapeutic_touch_model()
        params, lim = models.therapeutic_touch_data(), model.limits_array()
        walkers, wgs = 2 * 256 * world * 2, 256
        D, H = model.dimension, walkers // 2
        hs = H // world
        single = orc.OracleStretch(model, 21, walkers, params, wgs=wgs).init_position(22, lim)
        mine = orc.OracleStretch(model, 21, walkers, params, wgs=wgs).init_position(22, lim)
        for step in range(4):
            for half in (0, 1):
                before_x, before_lp = mine.xs.copy(), mine.lp.copy()
                mine.a_bare, mine.beta = 1.8, 1.0
                mine.half_bare(half)                              # the oracle computes the whole half ...
                act = slice(half * H, (half + 1) * H)
                xs_half = mine.xs.reshape(walkers, D)[act].copy()
                lp_half = mine.lp[act].copy()
                mine.xs[:], mine.lp[:] = before_x, before_lp      # ... but this rank only KEEPS its slice
                own = slice(rank * hs, (rank + 1) * hs)
                send = torch.from_numpy(np.concatenate([xs_half[own].reshape(-1), lp_half[own]]))
                got = [torch.zeros_like(send) for _ in range(world)]
                dist.all_gather(got, send)                        # exchange_half
                for r, t in enumerate(got):
                    t = t.numpy()
                    dst = slice(half * H + r * hs, half * H + (r + 1) * hs)
                    mine.xs.reshape(walkers, D)[dst] = t[:hs * D].reshape(hs, D)
                    mine.lp[dst] = t[hs * D:]
            mine.bare_counter += 1
            single.a_bare, single.beta = 1.8, 1.0
            single.move_bare()
            assert np.array_equal(mine.xs, single.xs), f"step {step}: partitioned chain left the single-process chain"
            assert np.array_equal(mine.lp, single.lp, equal_nan=True)
        (Path(out_dir) / f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_walker_partition_protocol_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_partition_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["ok0", "ok1"]
