"""The lane-private histogram kernel (k_hist_aos_lanes, kernels.cuh) finds the bin in fixed point — q = round(t' * 4096)
from one FFMA against 1.5 * 2^23 — and trusts it only when q is inside the limits and not a multiple of 4096; every
other value takes the reference's rule  t = fl(fl(d / range) * bins), bin = trunc(t)  (estimate.cu:60-75 as restated in
oracle/bayadera_oracle.c).  This test restates the device arithmetic in numpy and checks the claim the kernel rests on:
WHENEVER the fast path is taken its bin equals the reference's, also for values engineered to sit a few ulps from a bin
edge and with the reciprocal scale off by several ulps (the device uses the approximate __fdividef)."""
import numpy as np
import pytest

MAGIC = np.float32(12582912.0)


def reference_bin(d, rng, bins):
    t = (d / rng).astype(np.float32) * np.float32(bins)
    t = t.astype(np.float32)
    b = np.where(~(t > 0), 0, np.where(t >= bins, bins - 1, np.trunc(t))).astype(np.int64)
    return b


def fast_path(d, scale_fx, bins):
    u = (d.astype(np.float64) * np.float64(scale_fx) + np.float64(MAGIC)).astype(np.float32)   # FFMA: one rounding
    q = (u.view(np.uint32).astype(np.int64) - 0x4B400000) & 0xFFFFFFFF
    plain = (q < bins * 4096) & ((q & 4095) != 0)
    return plain, q >> 12


@pytest.mark.parametrize("bins", [32, 64, 256])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fast_path_bin_is_the_reference_bin(bins, seed):
    rng_ = np.random.default_rng(seed * 10 + bins)
    lo = np.float32(rng_.uniform(-50, 50))
    hi = np.float32(lo + rng_.uniform(1e-3, 200))
    rng = np.float32(hi - lo)
    n = 400_000
    x = rng_.uniform(lo, hi, n).astype(np.float32)
    # values a few ulps either side of every bin edge, the limits themselves, and values outside / non-finite
    edges = (lo + rng * (np.arange(bins + 1, dtype=np.float64) / bins)).astype(np.float32)
    near = np.concatenate([np.nextafter(edges, np.float32(np.inf)), np.nextafter(edges, np.float32(-np.inf)), edges])
    for _ in range(3):
        near = np.concatenate([near, np.nextafter(near, np.float32(np.inf)), np.nextafter(near, np.float32(-np.inf))])
    special = np.asarray([np.inf, -np.inf, np.nan, lo - 1, hi + 1, 3e38, -3e38], dtype=np.float32)
    x = np.concatenate([x, near, special])
    d = (x - lo).astype(np.float32)
    want = reference_bin(d, rng, bins)
    exact = np.float32(np.float64(bins) * 4096.0 / np.float64(rng))
    taken = 0
    with np.errstate(invalid="ignore", over="ignore"):
        for ulps in (-4, -2, 0, 2, 4):                      # __fdividef is good to 2 ulp
            scale = exact
            for _ in range(abs(ulps)):
                scale = np.nextafter(scale, np.float32(np.inf if ulps > 0 else -np.inf))
            plain, b = fast_path(d, scale, bins)
            assert np.array_equal(b[plain], want[plain]), (bins, seed, ulps)
            taken += int(plain.sum())
            # the slow path is rare: 1 value in 4096 among the uniformly drawn ones
            assert plain[:n].mean() > 0.999
    assert taken > 0
