/*
 * bayadera_oracle_rng.c — CPU ORACLE (test infrastructure, NOT product code): direct samplers.
 *
 * Restatement of K/cuda/rng/{gaussian,exponential,erlang}-sampler.cu
 * (K = /root/reference/src/device/uncomplicate/bayadera/internal/device/cuda; the uniform sampler is
 * orc_direct_uniform in bayadera_oracle.c).  Philox key {seed, 0xdecafaaa}, counter {gid, 0xf00dcafe,
 * 0xdeadbeef, 0xbeeff00d} (erlang: counter word 3 = draw index), 4 variates per work-item.
 *
 * Parity status: PINNED WITHIN FAST-MATH TOLERANCE.  The reference compiles these kernels with -use_fast_math
 * (C/internal/device/nvidia_gtx.clj:635-637), so its goldens (T/internal/nvidia_gtx_test.clj:56-107) embed
 * MUFU sin/cos/lg2/sqrt approximations that libm cannot reproduce bit-for-bit; tests/test_oracle_golden.py pins
 * this file on them at 2e-5 relative (first/last 4 of 10 000, max, min, mean).
 */
#include <math.h>
#include <stdint.h>

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
float orc_u01(uint32_t i);

static void draw4(uint32_t gid, uint32_t c3, uint32_t seed, float u[4]) {
    const uint32_t ctr[4] = {gid, 0xf00dcafeu, 0xdeadbeefu, c3};
    const uint32_t key[2] = {seed, 0xdecafaaau};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    for (int c = 0; c < 4; c++) u[c] = orc_u01(r[c]);
}

/* gaussian-sampler.cu:16-23 (Box-Muller), :48-51 */
void orc_direct_gaussian(uint32_t n, uint32_t seed, float mu, float sigma, float *x) {
    const float two_pi = 6.2831855f;
    for (uint32_t g = 0; g * 4 < n; g++) {
        float u[4], z[4];
        draw4(g, 0xbeeff00du, seed, u);
        z[0] = sinf(two_pi * u[0]) * sqrtf(-2.0f * logf(u[1]));
        z[1] = cosf(two_pi * u[0]) * sqrtf(-2.0f * logf(u[1]));
        z[2] = sinf(two_pi * u[2]) * sqrtf(-2.0f * logf(u[3]));
        z[3] = cosf(two_pi * u[2]) * sqrtf(-2.0f * logf(u[3]));
        for (int c = 0; c < 4; c++)
            if (4 * g + c < n) x[4 * g + c] = fmaf(z[c], sigma, mu);
    }
}

/* exponential-sampler.cu:36-39 */
void orc_direct_exponential(uint32_t n, uint32_t seed, float lambda, float *x) {
    for (uint32_t g = 0; g * 4 < n; g++) {
        float u[4];
        draw4(g, 0xbeeff00du, seed, u);
        for (int c = 0; c < 4; c++)
            if (4 * g + c < n) x[4 * g + c] = -1.0f / lambda * logf(1.0f - u[c]);
    }
}

/* erlang-sampler.cu:30-50: sum of k log-uniforms (counter word 3 = draw index), divided by -lambda */
void orc_direct_erlang(uint32_t n, uint32_t seed, float lambda, float k, float *x) {
    const float neg_lambda = -lambda;
    for (uint32_t g = 0; g * 4 < n; g++) {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, u[4];
        for (uint32_t i = 0; (float)i < k; i++) {
            draw4(g, i, seed, u);
            for (int c = 0; c < 4; c++) acc[c] += logf(u[c]);
        }
        for (int c = 0; c < 4; c++)
            if (4 * g + c < n) x[4 * g + c] = acc[c] / neg_lambda;
    }
}
