"""Host helpers of ``uncomplicate.bayadera.util`` (/root/reference/src/clojure/uncomplicate/bayadera/util.clj:33-110),
re-stated over numpy: the consumers of a ``Histogram``'s ``bin-ranks`` / ``pdf`` / ``limits`` columns.

These are the reference's HOST-side definitions (it computes them on the JVM); the device version for a whole
sampler histogram is ``B200Stretch.hdi`` / ``engine.hdi_histogram`` (C ABI ``bay_hdi``, ``bay_hdi_histogram``), and
``tests/test_gpu_next.py`` holds the two against each other and against the reference's goldens (util_test.clj).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def range_mapper(start1: float, end1: float, start2: float = None, end2: float = None) -> Callable:
    """util.clj:33-39 — affine map of [start1, end1] onto [start2, end2] (target given now or per call)."""
    if start2 is not None:
        return lambda value: start2 + (end2 - start2) * ((value - start1) / (end1 - start1))
    return lambda value, s2, e2: s2 + (e2 - s2) * ((value - start1) / (end1 - start1))


def bin_mapper(bin_count: int, lower: float, upper: float, offset: float = 0.5) -> Callable[[int], float]:
    """util.clj:41-50 — bin index -> coordinate of the bin (its centre for the default offset)."""
    bin_width = (upper - lower) / bin_count
    return lambda i: lower + bin_width * (i + offset)


def asum(x: Sequence[float]) -> float:
    """Neanderthal ``asum`` of a float vector: an fp32 accumulation (BLAS sasum; its order is unspecified — here
    and in the device kernel ``k_hdi`` it is sequential)."""
    acc = np.float32(0.0)
    for v in np.abs(np.asarray(x, dtype=np.float32)):
        acc = np.float32(acc + v)
    return float(acc)


def hdi_rank_count(bin_rank: Sequence[float], pdf: Sequence[float], mass: float = 0.95) -> int:
    """util.clj:52-65 — the smallest number of ranked bins whose mass reaches ``mass * asum(pdf)``; sequential
    accumulation in double like the reference's ``loop``."""
    pdf = np.asarray(pdf, dtype=np.float32)
    rank = np.asarray(bin_rank)
    density = mass * asum(pdf)
    acc, i, n = 0.0, 0, rank.shape[0]
    while i < n and acc < density:
        acc += float(pdf[int(rank[i])])
        i += 1
    return i


def hdi_bins(bin_rank: Sequence[float], hdi_cnt: int) -> List[float]:
    """util.clj:67-83 — ``[start0 end0 start1 end1 ...]``: the first ``hdi_cnt`` ranked bins grouped into runs."""
    v = np.sort(np.asarray(bin_rank, dtype=np.float64)[:hdi_cnt])
    regions, last = [float(v[0])], float(v[0])
    for b in v[1:]:
        b = float(b)
        if 1.5 < b - last:
            regions += [last, b]
        last = b
    regions.append(float(v[hdi_cnt - 1]))
    return regions


def hdi_regions(limits: Sequence[float], bin_rank: Sequence[float], hdi_cnt: int) -> np.ndarray:
    """util.clj:85-100 — (k, 2) array of [lower, upper] coordinates of the runs of ``hdi_bins``."""
    lower, upper = float(limits[0]), float(limits[1])
    bin_width = (upper - lower) / len(bin_rank)
    v = hdi_bins(bin_rank, hdi_cnt)
    out = np.zeros((len(v) // 2, 2), dtype=np.float32)
    for i in range(len(v) // 2):
        out[i, 0] = lower + bin_width * v[2 * i]
        out[i, 1] = lower + bin_width * (v[2 * i + 1] + 1.0)
    return out


def hdi(histogram, index: int, mass: float = 0.95) -> np.ndarray:
    """util.clj:102-110 — HDI regions holding ``mass`` for dimension ``index`` of a ``Histogram``."""
    limits, rank, pdf = histogram.limits[index], histogram.bin_ranks[index], histogram.pdf[index]
    return hdi_regions(limits, rank, hdi_rank_count(rank, pdf, mass))
