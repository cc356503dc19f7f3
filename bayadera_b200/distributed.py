"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous, NCCL inside the engine for the data path.

The reference has no distributed code; this is new work following SURVEY §8e:

* mode B, row-sharded dataset (GLM / row-additive posteriors, BASELINE config 4): every rank holds the SAME
  walkers and its own contiguous shard of the dataset rows; per half-step the per-walker partial
  log-likelihood sums are all-reduced (ncclAllReduce, double, sum) inside ``libbayadera_b200`` on the
  sampler's stream, so all ranks take bit-identical accept decisions and the replicas cannot diverge.

The only thing torch.distributed does here is carry the 128-byte ncclUniqueId from rank 0 to the others.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_rows(rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row shard [begin, end) of ``rows`` for ``rank`` of ``world`` (first rows%world ranks
    get one extra row)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_unique_id(make_id, rank: int, device=None) -> np.ndarray:
    """Rank 0 calls ``make_id()`` (-> 128 uint8); the bytes are broadcast over the default process group.
    Works with gloo (CPU tensor) and nccl (tensor on ``device``)."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.from_numpy(np.ascontiguousarray(make_id(), dtype=np.uint8).copy())
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return buf.cpu().numpy()


def init_engine_comm(factory, rank: int, world: int, device=None) -> None:
    """Create the engine's NCCL communicator (``bay_engine_comm_init``) on every rank of the default group."""
    from .engine import nccl_unique_id

    if world == 1:
        return
    uid = broadcast_unique_id(nccl_unique_id, rank, device)
    factory.comm_init(uid, world, rank)
