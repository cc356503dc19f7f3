"""Golden vectors of the reference's own tests for the stretch-move hot path.

T/ = /root/reference/test/clojure/uncomplicate/bayadera/  (values copied as literals with file:line;
identical in T/internal/amd_gcn_test.clj unless noted).  Common setup of the stretch goldens:
W = 44*256 = 11264 walkers, WGS = 256, seed 123, a = 2.0 (T/internal/nvidia_gtx_test.clj:161-170).
"""
from pathlib import Path

import numpy as np

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

WGS = 256
W = 44 * 256
SEED = 123
A = 2.0

# Random123 published known-answer tests for philox4x32-10 (kat_vectors): (ctr, key) -> out
PHILOX_KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]

# direct uniform sampler, seed 123, params [-99.9 200.1], n = 10000  (nvidia_gtx_test.clj:38-54)
DIRECT_UNIFORM = dict(
    params=(-99.9, 200.1), n=10000,
    first4=(108.93402099609375, 139.30517578125, -82.35221862792969, 47.42252731323242),
    last4=(17.43636131286621, 151.42117309570312, 42.78262710571289, 107.87583923339844),
    max=200.0757293701172, min=-99.86572265625, mean=51.331178125)

# the other direct samplers, seed 123, n = 10000 (nvidia_gtx_test.clj:56-107); compiled with -use_fast_math in the
# reference, so they embed MUFU approximations: pinned with a relative tolerance, not bit-for-bit
DIRECT = {
    "gaussian": dict(params=(100, 200.1),
                     first4=(-27.02083969116211, 55.2710075378418, 185.74417114257812, 322.7049255371094),
                     last4=(175.25135803222656, 7.720771312713623, 126.18173217773438, -69.4984359741211),
                     max=868.6444702148438, min=-610.0802612304688, mean=95.13473125),
    "erlang": dict(params=(2, 3),
                   first4=(1.571028470993042, 1.4484456777572632, 0.798355758190155, 1.1712464094161987),
                   last4=(0.7548800110816956, 2.2858035564422607, 1.19755220413208, 1.3439300060272217),
                   max=7.1595940589904785, min=0.04825383424758911, mean=1.5008442260742187),
    "exponential": dict(params=(4,),
                        first4=(0.29777514934539795, 0.3990693986415863, 0.015068254433572292, 0.1688637137413025),
                        last4=(0.12403398752212524, 0.45463457703590393, 0.16137929260730743, 0.29489004611968994),
                        max=2.3556251525878906, min=2.8555818062159233E-5, mean=0.25263106842041017),
}

# Uniform(-1,2) stretch, limits [-1 2]: xs[0..3] after each of 4 sample! calls (nvidia_gtx_test.clj:200-212)
UNIFORM_SAMPLES = [
    (0.7279692888259888, 1.81407630443573, 0.040318019688129425, 0.4697103202342987),
    (1.040357232093811, 1.4457943439483643, 0.3761849105358124, 1.5768483877182007),
    (1.047235131263733, 1.1567966938018799, 0.6802869439125061, 1.6528078317642212),
    (0.9332534670829773, 1.9495460987091064, 0.5958949327468872, 1.6429908275604248),
]
# after re-init, 2 bare steps, then ONE accu step with seeds (123,124), tags (1111,2222), step 0
# (nvidia_gtx_test.clj:237-258)
UNIFORM_ACCU_XS = (1.011080265045166, 1.615005373954773, 0.3426262140274048, 1.4122663736343384)
UNIFORM_ACCU_BLOCK_SUMS = (269.26575, 286.3589, 288.09372, 240.0009, 265.76953,
                           274.17465, 257.67914, 302.7213, 244.6228, 277.85284)
UNIFORM_ACCU_ACCEPT = (423, 422, 424, 428, 414, 439, 428, 409, 429, 409)
UNIFORM_ACCU_TOTAL = 5822.918

# Gaussian(3,1) stretch, limits [-7 7]: xs[0..3] after each of 3 sample! calls (nvidia_gtx_test.clj:260-277)
GAUSSIAN_SAMPLES = [
    (2.7455878257751465, 4.16290807723999, -2.1451826095581055, -0.14135171473026276),
    (3.9621310234069824, 2.9586496353149414, -0.5778038501739502, 5.025292873382568),
    (3.8151602745056152, 2.415064573287964, 0.7977100610733032, 5.292686939239502),
]

# Gaussian(3,1) burn-in 100 @ a=1.5 after init-position!(123) and init!(124)
# CUDA (nvidia_gtx_test.clj:289-296) and OpenCL (amd_gcn_test.clj:284-291) differ at ~1e-4 (fast-math exp)
BURN_IN_CUDA = dict(mean=2.9549713134765625, sd=0.996131420135498,
                    strided=(3.3301031589508057, 2.3123116493225098, 3.5831196308135986, 3.3420889377593994,
                             4.830397605895996, 2.7044715881347656, 2.064502716064453, 3.4433465003967285))
BURN_IN_OPENCL = dict(mean=2.9549477100372314, sd=0.995916485786438,
                      strided=(3.3298428058624268, 2.3125054836273193, 3.583237409591675, 3.3418703079223633,
                               4.829986572265625, 2.704503059387207, 2.064406156539917, 3.443021535873413))

# Gaussian(200,1), W = 2*44*256, a = 8, limits [180 220], burn-in 5120:
# acc-rate 0.488 (CUDA, biased by memset 1: SURVEY App. B-3) / 0.485 (OpenCL); tau 12.5 +- 0.5
# (nvidia_gtx_test.clj:298-318, amd_gcn_test.clj:293-312)
ACC_RATE_OPENCL = 0.485
TAU = 12.5

# acor fixtures (nvidia_gtx_test.clj:320-358): row 0 = data, row 1 = f(data)
ACOR = {
    67: dict(row1=lambda d: 2 * d, tau=12.072992324829102, sigma=0.45055490732192993),
    367: dict(row1=lambda d: d + 1, tau=20.156665802001953, sigma=0.1837492138147354),
    112640: dict(row1=lambda d: 2 * d, tau=20.566, sigma=0.009),
}


def acor_fixture(n: int) -> np.ndarray:
    """dim x n column-major series (flat index 2*t + d) as the reference test builds it."""
    data = np.load(GOLDEN_DIR / "acor_fixtures.npz")[f"acor_{n}"].astype(np.float32)
    m = np.stack([data, ACOR[n]["row1"](data).astype(np.float32)], axis=1)  # n x 2
    return np.ascontiguousarray(m, dtype=np.float32)


# HDI helpers: T/util_test.clj:21-57 — 79 bins on [1, 7]
HDI_LIMITS = (1.0, 7.0)
HDI_PDF = (0.02, 0.01, 0.04, 0.05, 0.07, 0.01, 0.2, 0.1, 0.3, 0.1, 0.02) + (0.01,) * 68
HDI_BIN_RANK = (8, 6, 7, 9, 4, 3, 2, 0, 10, 1, 5) + tuple(range(11, 79))
# (mass, divide-by-asum?) -> hdi-rank-count   (util_test.clj:35-41)
HDI_RANK_COUNTS = [(0.1, False, 1), (0.3, True, 1), (0.3, False, 2), (0.7, True, 4), (0.86, True, 7), (0.9, True, 9),
                   (0.95, True, 14)]
# hdi-cnt -> hdi-bins   (util_test.clj:45-50)
HDI_BINS = {1: [8.0, 8.0], 2: [6.0, 6.0, 8.0, 8.0], 5: [4.0, 4.0, 6.0, 9.0], 6: [3.0, 4.0, 6.0, 9.0], 12: [0.0, 11.0],
            15: [0.0, 14.0]}
# hdi-cnt -> (regions column-major [lo0 hi0 lo1 hi1], nrm2 tolerance)   (util_test.clj:54-58)
HDI_REGIONS = {1: ([1.61, 1.68], 0.005), 2: ([1.46, 1.53, 1.61, 1.68], 0.007), 5: ([1.30, 1.38, 1.46, 1.76], 0.006),
               6: ([1.23, 1.38, 1.46, 1.76], 0.006), 12: ([1.00, 1.91], 0.002)}
