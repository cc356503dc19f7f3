/*
 * bayadera_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's stretch-move hot path, used only as
 * the checker for the CUDA engine (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing under
 * bayadera_b200/ may import, link or call this file.
 *
 * Path aliases used in the citations below (same as SURVEY.md):
 *   K/cuda/…   = /root/reference/src/device/uncomplicate/bayadera/internal/device/cuda/…
 *   K/opencl/… = /root/reference/src/device/uncomplicate/bayadera/internal/device/opencl/…
 *   C/…        = /root/reference/src/clojure/uncomplicate/bayadera/…
 *   T/…        = /root/reference/test/clojure/uncomplicate/bayadera/…
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * every golden vector the reference tests hold for the path (Philox KATs,
 * direct-uniform goldens, stretch positions for Uniform(-1,2) and
 * Gaussian(3,1), per-block accept counts and mean sums, the three acor
 * fixtures, acc-rate 0.485).  Random123's Philox4x32-10 is a third-party
 * dependency that is NOT under /root/reference (shipped in
 * uncomplicate/neanderthal 0.25.7-SNAPSHOT, see C/internal/device/nvidia_gtx.clj:648-656);
 * its published algorithm is restated here and pinned by the public KATs.
 *
 * On top of the goldens, tests/test_oracle_reference_text.py runs this file beside the REFERENCE'S OWN KERNEL TEXT
 * (K/opencl/engines/amd-gcn-mcmc-stretch.cl compiled by gcc behind oracle/ref_shim/cl_shim.h into oracle/_ref/):
 * stretch_move_bare and logfn agree bit for bit over 57 steps, init_walkers to 2 ulp (FMA choice), and the
 * literal_partner mode is that text for DIM = 2.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the three FMA contractions that the
 * reference's nvcc build performs are written out explicitly with fmaf().
 *
 * Layout: everything here is the reference's own AoS layout — walker k of a
 * half-ensemble occupies X[k*D .. k*D+D) — because that is what the goldens
 * are expressed in.  The CUDA engine stores SoA internally and converts at
 * the C-ABI boundary.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* The model callback: the reference's LOGFN contract
 * (K/cuda/engines/nvidia-gtx-mcmc-stretch.cu:85, C/internal/device/models.clj:93-100). */
typedef float (*orc_logfn_t)(uint32_t data_len, uint32_t params_len,
                             const float *params, uint32_t dim, const float *x);

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Random123; published algorithm, Salmon et al. SC'11).      */
/* Call sites: mcmc-stretch.cu:52-62, 170-182; rng/uniform-sampler.cu:23-37. */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uint32 -> open interval (0,1): K/cuda/rng/uniform-sampler.cu:11-13. */
float orc_u01(uint32_t i) {
    return (0.5f + (float)(i >> 9)) * 1.1920928955078125e-7f;
}

/* Direct uniform sampler (used only to pin Philox + u01 on the goldens at
 * T/internal/nvidia_gtx_test.clj:44-54): K/cuda/rng/uniform-sampler.cu:15-44.
 * nvcc contracts u*range+lower into one fma. */
void orc_direct_uniform(uint32_t n, uint32_t seed, float lower, float upper, float *x) {
    const float range = upper - lower;
    for (uint32_t g = 0; g * 4 < n; g++) {
        const uint32_t ctr[4] = {g, 0xf00dcafeu, 0xdeadbeefu, 0xbeeff00du};
        const uint32_t key[2] = {seed, 0xdecafaaau};
        uint32_t r[4];
        orc_philox4x32_10(ctr, key, r);
        for (int c = 0; c < 4; c++)
            if (4 * g + c < n) x[4 * g + c] = fmaf(orc_u01(r[c]), range, lower);
    }
}

/* ------------------------------------------------------------------------- */
/* init_walkers: mcmc-stretch.cu:158-195.  n4 = W*D/4 work-items; limits is  */
/* 2 x D column-major (lo_d, hi_d).  fma contraction pinned by the goldens   */
/* (SURVEY Appendix A-3).                                                    */
/* ------------------------------------------------------------------------- */
void orc_init_walkers(uint32_t n4, uint32_t dim, uint32_t seed, const float *limits, float *xs) {
    #pragma omp parallel for schedule(static)
    for (uint32_t g = 0; g < n4; g++) {
        const uint32_t ctr[4] = {g, 0xf00dcafeu, 0xdeadbeefu, 0xbeeff00du};
        const uint32_t key[2] = {seed, 0xdecafaaau};
        uint32_t r[4];
        orc_philox4x32_10(ctr, key, r);
        for (uint32_t c = 0; c < 4; c++) {
            const uint32_t e = 4 * g + c;
            const float lo = limits[2 * (e % dim)];
            const float hi = limits[2 * (e % dim) + 1];
            const float u = orc_u01(r[c]);
            xs[e] = fmaf(u, hi, (1.0f - u) * lo);
        }
    }
}

/* logfn: mcmc-stretch.cu:197-206. */
void orc_logfn(orc_logfn_t f, uint32_t n, uint32_t dim, uint32_t data_len, uint32_t params_len,
               const float *params, const float *x, float *res) {
    #pragma omp parallel for schedule(static)
    for (uint32_t g = 0; g < n; g++)
        res[g] = f(data_len, params_len, params, dim, x + (size_t)dim * g);
}

/* The three proposal coefficients of z = A u^2 + B u + C, g(z) ∝ 1/sqrt(z) on
 * [1/a, a] (mcmc-stretch.cu:69-71), evaluated in IEEE fp32 in source order. */
void orc_stretch_coeffs(float a, float abc[3]) {
    const float inv = 1.0f / a;
    abc[0] = (a - 2.0f) + inv;
    abc[1] = 2.0f * (1.0f - inv);
    abc[2] = inv;
}

/* ------------------------------------------------------------------------- */
/* One half-ensemble stretch move: device fn stretch_move,                    */
/* mcmc-stretch.cu:35-99 (twin K/opencl/engines/amd-gcn-mcmc-stretch.cl:31-81)*/
/*                                                                           */
/* literal_partner = 0: partner is walker j = (uint)(u.x*K), window          */
/*   Scompl[j*D .. j*D+D)  (the intended semantics; identical to the         */
/*   reference for D = 1, which is all the reference tests exercise).        */
/* literal_partner = 1: the reference's literal element offset               */
/*   j0 = (uint)(u.x*K*D), window Scompl[j0 .. j0+D) clamped to the half     */
/*    — for D>1 decision-parity experiments only.       */
/*                                                                           */
/* Optional per-walker diagnostics (any may be NULL): acc[k] 0/1, ly[k] the  */
/* proposal's log-density, q[k] the acceptance ratio, uz[k] the uniform it   */
/* was compared with.  Returns the number of accepted walkers.               */
/* ------------------------------------------------------------------------- */
uint32_t orc_stretch_half(orc_logfn_t f, uint32_t K, uint32_t dim, uint32_t seed, uint32_t tag,
                          uint32_t step, uint32_t data_len, uint32_t params_len,
                          const float *params, const float *Scompl, float *X, float *logfn_X,
                          float a, float beta, int literal_partner,
                          uint8_t *acc, float *ly_out, float *q_out, float *uz_out) {
    float abc[3];
    orc_stretch_coeffs(a, abc);
    uint32_t total = 0;
    #pragma omp parallel reduction(+:total)
    {
        float *Y = (float *)malloc(sizeof(float) * (dim ? dim : 1));
        #pragma omp for schedule(static)
        for (uint32_t k = 0; k < K; k++) {
            const uint32_t ctr[4] = {k, step, tag, 0xbeeff00du};
            const uint32_t key[2] = {seed, 0xdecafbadu};
            uint32_t r[4];
            orc_philox4x32_10(ctr, key, r);
            const float ux = orc_u01(r[0]), uy = orc_u01(r[1]), uz = orc_u01(r[2]);

            /* contraction pinned by the goldens */
            const float z = fmaf(abc[0] * uy, uy, abc[1] * uy) + abc[2];

            size_t j0;
            if (literal_partner) {
                j0 = (size_t)(uint32_t)((ux * (float)K) * (float)dim);
                if (j0 + dim > (size_t)K * dim) j0 = (size_t)K * dim - dim;
            } else {
                uint32_t j = (uint32_t)(ux * (float)K);
                if (j >= K) j = K - 1;
                j0 = (size_t)j * dim;
            }
            const size_t k0 = (size_t)k * dim;
            for (uint32_t i = 0; i < dim; i++) {
                const float xj = Scompl[j0 + i];
                Y[i] = fmaf(z, X[k0 + i] - xj, xj); /* Appendix A-6 */
            }
            const float ly = f(data_len, params_len, params, dim, Y);
            const float q = isfinite(ly)
                ? powf(z, (float)(dim - 1)) * expf(beta * (ly - logfn_X[k]))
                : 0.0f;
            const int ok = uz <= q;
            if (ok) {
                memcpy(X + k0, Y, sizeof(float) * dim);
                logfn_X[k] = ly;
                total += 1;
            }
            if (acc) acc[k] = (uint8_t)ok;
            if (ly_out) ly_out[k] = ly;
            if (q_out) q_out[k] = q;
            if (uz_out) uz_out[k] = uz;
        }
        free(Y);
    }
    return total;
}

/* ------------------------------------------------------------------------- */
/* Block reductions.  The reference's shared-memory halving tree             */
/* (block_reduction_sum_uint, mcmc-stretch.cu:8-33; ClojureCUDA's            */
/* block_reduction_sum has the same shape — not in /root/reference, pinned   */
/* bit-exactly by the 10 block sums at T/internal/nvidia_gtx_test.clj:251).  */
/* v has wgs entries and is destroyed.                                       */
/* ------------------------------------------------------------------------- */
float orc_block_tree_sum(float *v, uint32_t wgs) {
    uint32_t i = wgs;
    while (i > 1) {
        const uint32_t odd = i & 1u;
        i >>= 1;
        for (uint32_t l = 0; l < i; l++) v[l] = v[l] + v[l + i];
        if (odd) v[i - 1] = v[i - 1] + v[2 * i]; /* include_odd lane */
    }
    return v[0];
}

/* The accu epilogue of stretch_move_accu (mcmc-stretch.cu:101-134) for one
 * launch: per block b of wgs walkers, accept[b] += #accepted and
 * blk_means[i*G + b] += sum over the block of the post-move X[k,i]. */
void orc_accu_epilogue(uint32_t K, uint32_t dim, uint32_t wgs, const float *X, const uint8_t *acc,
                       uint32_t *accept, float *blk_means) {
    const uint32_t G = (K + wgs - 1) / wgs;
    float *v = (float *)malloc(sizeof(float) * wgs);
    for (uint32_t b = 0; b < G; b++) {
        uint32_t cnt = 0;
        for (uint32_t l = 0; l < wgs; l++) {
            const uint32_t k = b * wgs + l;
            if (k < K && acc[k]) cnt++;
        }
        accept[b] += cnt;
        for (uint32_t i = 0; i < dim; i++) {
            for (uint32_t l = 0; l < wgs; l++) {
                const uint32_t k = b * wgs + l;
                v[l] = (k < K) ? X[(size_t)k * dim + i] : 0.0f;
            }
            blk_means[(size_t)i * G + b] += orc_block_tree_sum(v, wgs);
        }
    }
    free(v);
}

/* Ensemble mean of one accu step from the per-block sums:
 * sum_means_vertical (mcmc-stretch.cu:250-260) + scal! 0.5/(WGS*G)
 * (C/internal/device/nvidia_gtx.clj:457-460).  The multi-level order of
 * launch-reduce! is unpinned; a sequential fp32 sum over blocks
 * is the oracle's definition, compared against the engine with a tolerance. */
void orc_step_means(uint32_t G, uint32_t dim, uint32_t wgs, const float *blk_means, float *means) {
    for (uint32_t i = 0; i < dim; i++) {
        float s = 0.0f;
        for (uint32_t b = 0; b < G; b++) s += blk_means[(size_t)i * G + b];
        means[i] = s * (0.5f / ((float)wgs * (float)G));
    }
}

/* ------------------------------------------------------------------------- */
/* Estimate engine: K/cuda/engines/nvidia-gtx-estimate.cu                    */
/* data is the reference's AoS view data[offset + ld*col + row].             */
/* ------------------------------------------------------------------------- */

/* min_max_reduce / min_max_reduction (estimate.cu:48-115) with the OpenCL
 * twin's semantics (no (0,0) padding — SURVEY Appendix B-4).
 * limits out: 2 x dim column-major (min_d, max_d). */
void orc_min_max(uint32_t dim, uint64_t n, const float *data, uint64_t offset, uint64_t ld,
                 float *limits) {
    for (uint32_t d = 0; d < dim; d++) {
        float lo = INFINITY, hi = -INFINITY;
        for (uint64_t g = 0; g < n; g++) {
            const float x = data[offset + ld * g + d];
            lo = fminf(lo, x);
            hi = fmaxf(hi, x);
        }
        limits[2 * d] = lo;
        limits[2 * d + 1] = hi;
    }
}

/* histogram (estimate.cu:5-32): bin = min((uint)((x-lo)/(hi-lo)*WGS), WGS-1).
 * The oracle definition: IEEE fp32 sub, div, mul, then
 * floor, clamped to [0, WGS-1] (samples below lo in later histogram! cycles
 * saturate to bin 0 on NVIDIA hardware; NaN goes to bin 0 as well).
 * res is wgs x dim column-major (res[wgs*d + bin]) and is ACCUMULATED into. */
void orc_histogram(uint32_t dim, uint64_t n, uint32_t wgs, const float *limits, const float *data,
                   uint64_t offset, uint64_t ld, uint32_t *res) {
    #pragma omp parallel for schedule(static)
    for (uint32_t d = 0; d < dim; d++) {
        const float lo = limits[2 * d], hi = limits[2 * d + 1];
        const float range = hi - lo;
        for (uint64_t g = 0; g < n; g++) {
            const float x = data[offset + ld * g + d];
            const float t = ((x - lo) / range) * (float)wgs;
            uint32_t bin;
            if (!(t > 0.0f)) bin = 0;                 /* negative, -0, NaN */
            else if (t >= (float)wgs) bin = wgs - 1;  /* includes +inf */
            else bin = (uint32_t)t;
            res[(size_t)wgs * d + bin] += 1;
        }
    }
}

/* uint_to_real (estimate.cu:34-44) with alpha = WGS/n computed as the host
 * does (double division cast to float, nvidia_gtx.clj:507). */
void orc_uint_to_real(uint32_t wgs, uint32_t dim, uint64_t n, const float *limits,
                      const uint32_t *counts, float *pdf) {
    const float alpha = (float)((double)wgs / (double)n);
    for (uint32_t d = 0; d < dim; d++) {
        const float w = limits[2 * d + 1] - limits[2 * d];
        for (uint32_t b = 0; b < wgs; b++)
            pdf[(size_t)wgs * d + b] = (alpha / w) * (float)counts[(size_t)wgs * d + b];
    }
}

/* bitonic_local (estimate.cu:117-147): the exact compare-exchange network the
 * reference runs per dimension, simulated lane by lane, so that the order
 * inside groups of equal pdf is reproduced too.  out[wgs*d + r] = index of the
 * bin with rank r (decreasing mass). wgs must be a power of two. */
void orc_bin_ranks(uint32_t wgs, uint32_t dim, const float *pdf, float *out) {
    float *vx = (float *)malloc(sizeof(float) * wgs), *vy = (float *)malloc(sizeof(float) * wgs);
    float *nx = (float *)malloc(sizeof(float) * wgs), *ny = (float *)malloc(sizeof(float) * wgs);
    for (uint32_t d = 0; d < dim; d++) {
        for (uint32_t l = 0; l < wgs; l++) { vx[l] = (float)l; vy[l] = pdf[(size_t)wgs * d + l]; }
        for (uint32_t length = 1; length < wgs; length <<= 1) {
            for (uint32_t inc = length; inc > 0; inc >>= 1) {
                for (uint32_t l = 0; l < wgs; l++) {
                    const int direction = (l & (length << 1)) != 0;
                    const uint32_t j = l ^ inc;
                    const int smaller = (vy[l] < vy[j]) || (vy[j] == vy[l] && j < l);
                    const int swap = smaller ^ (j < l) ^ direction;
                    nx[l] = swap ? vx[j] : vx[l];
                    ny[l] = swap ? vy[j] : vy[l];
                }
                memcpy(vx, nx, sizeof(float) * wgs);
                memcpy(vy, ny, sizeof(float) * wgs);
            }
        }
        for (uint32_t l = 0; l < wgs; l++) out[(size_t)wgs * d + l] = vx[l];
    }
    free(vx); free(vy); free(nx); free(ny);
}

/* mean_reduce / variance_reduce (estimate.cu:149-187; host
 * nvidia_gtx.clj:161-198, 514-537): two passes, population variance.  The
 * reference accumulates in fp32 through an unpinned multi-level tree; the
 * oracle accumulates in double and the engine is compared with a tolerance. */
void orc_mean_variance(uint32_t dim, uint64_t n, const float *data, uint64_t offset, uint64_t ld,
                       float *mean, float *variance) {
    #pragma omp parallel for schedule(static)
    for (uint32_t d = 0; d < dim; d++) {
        double s = 0.0;
        for (uint64_t g = 0; g < n; g++) s += data[offset + ld * g + d];
        const float mu = (float)(s / (double)n);
        double v = 0.0;
        for (uint64_t g = 0; g < n; g++) {
            const float diff = data[offset + ld * g + d] - mu;
            v += (double)(diff * diff);
        }
        mean[d] = mu;
        if (variance) variance[d] = (float)(v / (double)n);
    }
}

/* ------------------------------------------------------------------------- */
/* acor: K/cuda/engines/nvidia-gtx-acor.cu:17-168 + host                     */
/* C/internal/device/nvidia_gtx.clj:230-278.  series is dim x n column-major */
/* (series[dim*t + d]) and is modified in place (mean subtraction, pairwise  */
/* sums) exactly like the reference's buffer.  Returns 0, or -1 when         */
/* 5*lag > n (the reference throws IllegalArgumentException).                */
/* ------------------------------------------------------------------------- */
static void acor_pass(uint32_t n, uint32_t wgs, uint32_t stride, uint32_t dim_id, uint32_t lag,
                      const float *x, float *c0_out, float *d_out) {
    /* acor_1d (acor.cu:44-108), launched with blocks of blk = min(WGS, n)
     * (acor.cu:124-126, 143-144).  Entry t contributes iff t+lag < n.  Its lag
     * window x[t+1..t+lag] is staged through shared memory: an element u that
     * sits in the SAME block as t was stored as (u+lag < n ? x[u] : 0), an
     * element in the NEXT block was loaded raw (the load_lag path, acor.cu:57-60).
     * That block-size dependence is part of the reference's observable result
     * and is kept.  Per-entry terms are fp32 (sequential window sum, as the
     * kernel's loop); the cross-entry sums use float atomics in the reference
     * (order unpinned) and double here. */
    const uint32_t blk = wgs < n ? wgs : n;
    double c0 = 0.0, dd = 0.0;
    for (uint32_t t = 0; t + lag < n; t++) {
        const float xt = x[(size_t)t * stride + dim_id];
        float xacc = 0.0f;
        for (uint32_t s = 1; s <= lag; s++) {
            const uint32_t u = t + s;
            const int same_block = (u / blk) == (t / blk);
            const float xu = (same_block && !(u + lag < n)) ? 0.0f : x[(size_t)u * stride + dim_id];
            xacc += xu;
        }
        c0 += (double)(xt * xt);
        dd += (double)(xt * (xt + 2.0f * xacc));
    }
    *c0_out = (float)c0;
    *d_out = (float)dd;
}

int orc_acor(uint32_t dim, uint32_t n, uint32_t wgs, float *series,
             float *tau, float *mean, float *sigma, uint32_t *lag_out) {
    const uint32_t min_fac = 5, win_mult = 5, min_lag = 10, max_lag = 64;
    uint32_t lag = n / min_fac;
    if (lag > wgs) lag = wgs;
    if (lag > max_lag) lag = max_lag;
    if (lag < min_lag) lag = min_lag;
    if (lag_out) *lag_out = lag;
    if ((uint64_t)lag * min_fac > n) return -1;

    for (uint32_t d = 0; d < dim; d++) {
        double s = 0.0;
        for (uint32_t t = 0; t < n; t++) s += series[(size_t)t * dim + d];
        const float mu = (float)(s / (double)n);
        mean[d] = mu;
        for (uint32_t t = 0; t < n; t++) series[(size_t)t * dim + d] -= mu;
    }
    for (uint32_t d = 0; d < dim; d++) {
        float c0v, dv;
        acor_pass(n, wgs, dim, d, lag, series, &c0v, &dv);
        const float c0 = c0v;              /* kept from the first pass (acor.cu:136) */
        float tau_d = dv / c0;
        uint32_t lag2 = lag, n2 = n, stride = 1;
        while (min_lag < lag2 && (float)lag2 < tau_d * (float)win_mult) {
            n2 /= 2;
            lag2 = (lag * win_mult < n2) ? lag : (n2 / win_mult > 10u ? n2 / win_mult : 10u);
            /* sum_pairwise (acor.cu:30-42) */
            for (uint32_t g = 0; g < n2; g++)
                series[(size_t)(2 * g) * stride * dim + d] +=
                    series[(size_t)(2 * g + 1) * stride * dim + d];
            stride *= 2;
            /* acor.cu:147 re-declares c0 inside the loop body (= c0acc of the PREVIOUS level),
             * and acor.cu:156 divides the new level's d by that shadowing value. */
            const float c0_prev = c0v;
            acor_pass(n2, wgs, stride * dim, d, lag2, series, &c0v, &dv);
            tau_d = dv / c0_prev;
        }
        const float scale = (float)stride * (float)(n2 - lag2);
        tau[d] = dv * (float)(n - lag) / (scale * c0);
        sigma[d] = sqrtf(dv / (scale * (float)n));
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Row-additive dataset likelihoods — the new engine's opt-in fast path      */
/* (no counterpart kernel in the reference: there every thread loops over    */
/* the whole dataset serially, e.g. K/cuda/distributions/gaussian.cu:40-42). */
/* The oracle for it is simply the serial model callback above.              */
/* ------------------------------------------------------------------------- */

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline asks for all host cores explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
