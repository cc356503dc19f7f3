// bayadera_b200 — model-independent CUDA kernels (compiled by nvcc for sm_100a).
//
// These replace K/cuda/engines/nvidia-gtx-estimate.cu (histogram, min/max,
// uint_to_real, bitonic_local, mean/variance), the init_walkers /
// sum_accept_* / sum_means_vertical kernels of nvidia-gtx-mcmc-stretch.cu and
// K/cuda/engines/nvidia-gtx-acor.cu of the reference
// (K/ = /root/reference/src/device/uncomplicate/bayadera/internal/device/cuda/).
//
// State layout: structure-of-arrays, element (dim d, sample w) at xs[d*pitch + w].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bay {

// ------------------------------------------------------------------ Philox --
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t i) {
    return __fmul_rn(__fadd_rn(0.5f, (float)(i >> 9)), 1.1920928955078125e-7f);
}

// init_walkers (mcmc-stretch.cu:158-195): flat AoS element e = w*D + d gets
// uniform #(e%4) of Philox counter e/4, mapped into limits[d]; stored SoA.
__global__ void k_init_walkers(uint32_t n4, uint32_t dim, uint32_t seed,
                               const float* __restrict__ limits, float* __restrict__ xs,
                               uint32_t pitch) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n4) return;
    uint32_t r[4];
    philox4x32_10(g, 0xf00dcafeu, 0xdeadbeefu, 0xbeeff00du, seed, 0xdecafaaau, r);
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        const uint32_t e = 4u * g + c;
        const uint32_t w = e / dim, d = e - w * dim;
        const float lo = limits[2 * d], hi = limits[2 * d + 1];
        const float u = u01(r[c]);
        xs[(size_t)d * pitch + w] = __fmaf_rn(u, hi, __fmul_rn(__fsub_rn(1.0f, u), lo));
    }
}

// ------------------------------------------------------------- transposes --
// SoA (dim x n, pitch) -> AoS column-major dim x n with leading dimension ld: out[w*ld + d]
__global__ void k_soa_to_aos(const float* __restrict__ xs, uint32_t pitch, uint32_t dim,
                             uint64_t n, float* __restrict__ out, uint32_t ld) {
    __shared__ float tile[32][33];
    const uint64_t w0 = (uint64_t)blockIdx.x * 32;
    const uint32_t d0 = blockIdx.y * 32;
    for (uint32_t r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint32_t d = d0 + r;
        const uint64_t w = w0 + threadIdx.x;
        if (d < dim && w < n) tile[r][threadIdx.x] = xs[(size_t)d * pitch + w];
    }
    __syncthreads();
    for (uint32_t r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint64_t w = w0 + r;
        const uint32_t d = d0 + threadIdx.x;
        if (d < dim && w < n) out[w * ld + d] = tile[threadIdx.x][r];
    }
}

// AoS view data[offset + ld*col + row] (row < dim) -> SoA
__global__ void k_aos_to_soa(const float* __restrict__ in, uint64_t offset, uint64_t ld,
                             uint32_t dim, uint64_t n, float* __restrict__ xs, uint64_t pitch) {
    __shared__ float tile[32][33];
    const uint64_t w0 = (uint64_t)blockIdx.x * 32;
    const uint32_t d0 = blockIdx.y * 32;
    for (uint32_t r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint64_t w = w0 + r;
        const uint32_t d = d0 + threadIdx.x;
        if (d < dim && w < n) tile[r][threadIdx.x] = in[offset + ld * w + d];
    }
    __syncthreads();
    for (uint32_t r = threadIdx.y; r < 32; r += blockDim.y) {
        const uint32_t d = d0 + r;
        const uint64_t w = w0 + threadIdx.x;
        if (d < dim && w < n) xs[(size_t)d * pitch + w] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------- min/max --
// Order-preserving float <-> uint32 map so that integer atomics give float min/max.
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void k_minmax_init(uint32_t dim, uint32_t* __restrict__ mm) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < dim) { mm[2 * d] = 0xffffffffu; mm[2 * d + 1] = 0u; }
}

// min_max_reduce (estimate.cu:100-115) over SoA rows; grid = (blocks, dim).
// Identities are +/-inf, NaNs are ignored.
__global__ void k_minmax_soa(const float* __restrict__ xs, uint64_t pitch, uint64_t n,
                             uint32_t* __restrict__ mm) {
    const uint32_t d = blockIdx.y;
    const float* row = xs + (size_t)d * pitch;
    float lo = INFINITY, hi = -INFINITY;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((pitch & 3) == 0) && ((((uintptr_t)xs) & 15) == 0);
    if (vec) {
        // every CTA streams ONE contiguous segment of the row (sequential 4 KB blocks: DRAM pages stay open)
        const uint64_t n4 = n >> 2, seg = (n4 + gridDim.x - 1) / gridDim.x;
        const uint64_t q0 = (uint64_t)blockIdx.x * seg, q1 = q0 + seg < n4 ? q0 + seg : n4;
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (uint64_t q = q0 + threadIdx.x; q < q1; q += 4 * blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
                v[r] = (q + r * blockDim.x < q1) ? __ldcs(row4 + q + r * blockDim.x) : make_float4(NAN, NAN, NAN, NAN);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                lo = fminf(fminf(lo, v[r].x), fminf(fminf(v[r].y, v[r].z), v[r].w));
                hi = fmaxf(fmaxf(hi, v[r].x), fmaxf(fmaxf(v[r].y, v[r].z), v[r].w));
            }
        }
        for (uint64_t q = (n4 << 2) + i; q < n; q += stride) { lo = fminf(lo, row[q]); hi = fmaxf(hi, row[q]); }
    } else {
        for (uint64_t q = i; q < n; q += stride) { lo = fminf(lo, row[q]); hi = fmaxf(hi, row[q]); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[32], shi[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
    __syncthreads();
    if (warp == 0) {
        lo = lane < nw ? slo[lane] : INFINITY;
        hi = lane < nw ? shi[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(&mm[2 * d], f2ord(lo));
            atomicMax(&mm[2 * d + 1], f2ord(hi));
        }
    }
}

__global__ void k_minmax_decode(uint32_t dim, const uint32_t* __restrict__ mm, float* __restrict__ limits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * dim) limits[i] = ord2f(mm[i]);
}

// -------------------------------------------------------------- histogram --
// histogram (estimate.cu:5-32).  bin = clamp(floor(((x-lo)/(hi-lo))*BINS), 0, BINS-1)
// in IEEE fp32 — integer counts, bit-exact against the oracle.
// Fast path: t' = (x-lo) * (BINS/range) differs from the defining expression by a few ulp, so floor(t') is the
// defining bin unless t' sits within 1e-6 relative of an integer; only then (and for NaN / out-of-range values)
// is the IEEE division evaluated.  The counts stay bit-exact while the common case costs a multiply.
__device__ __noinline__ uint32_t hist_bin_exact(float d, float range, float fbins, uint32_t bins) {
    const float t = __fmul_rn(__fdiv_rn(d, range), fbins);
    uint32_t b;
    if (!(t > 0.0f)) b = 0;
    else if (t >= fbins) b = bins - 1;
    else b = (uint32_t)t;
    return b;
}

// scale = BINS/range (any rounding), guard = BINS - 0.5
__device__ __forceinline__ uint32_t hist_bin(float x, float lo, float range, float fbins, uint32_t bins,
                                             float scale, float guard) {
    const float d = __fsub_rn(x, lo);
    const float ta = d * scale;
    const bool interior = ta > 0.5f && ta < guard && fabsf(ta - rintf(ta)) > fmaf(1e-6f, ta, 1e-6f);
    if (__builtin_expect(interior, 1)) return (uint32_t)ta;
    return hist_bin_exact(d, range, fbins, bins);     // kept out of line: rare, and it would be if-converted
}

// grid = (chunks, dim).  Shared memory: NSUB privatised sub-histograms of `bins`
// counters (one per warp group) to cut same-address atomic serialisation when
// the ensemble concentrates in a few bins; merged with one global atomic per
// non-empty bin per block.  counts is bins x dim column-major, accumulated.
__global__ void k_hist_soa(const float* __restrict__ xs, uint64_t pitch, uint64_t n,
                           const float* __restrict__ limits, uint32_t bins, uint32_t nsub,
                           uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t sh[];
    const uint32_t d = blockIdx.y;
    for (uint32_t i = threadIdx.x; i < bins * nsub; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const float lo = limits[2 * d], hi = limits[2 * d + 1];
    const float range = __fsub_rn(hi, lo), fbins = (float)bins;
    const float scale = __fdividef(fbins, range), guard = fbins - 0.5f;
    uint32_t* mine = sh + ((threadIdx.x >> 5) % nsub) * bins;
    const float* row = xs + (size_t)d * pitch;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((pitch & 3) == 0) && ((((uintptr_t)xs) & 15) == 0);
    if (vec) {
        // every CTA streams ONE contiguous segment of the row (sequential 4 KB blocks: DRAM pages stay open)
        const uint64_t n4 = n >> 2, seg = (n4 + gridDim.x - 1) / gridDim.x;
        const uint64_t q0 = (uint64_t)blockIdx.x * seg, q1 = q0 + seg < n4 ? q0 + seg : n4;
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (uint64_t q = q0 + threadIdx.x; q < q1; q += 2 * blockDim.x) {
            const bool two = q + blockDim.x < q1;
            const float4 v = __ldcs(row4 + q);
            const float4 w = two ? __ldcs(row4 + q + blockDim.x) : v;
            atomicAdd(&mine[hist_bin(v.x, lo, range, fbins, bins, scale, guard)], 1u);
            atomicAdd(&mine[hist_bin(v.y, lo, range, fbins, bins, scale, guard)], 1u);
            atomicAdd(&mine[hist_bin(v.z, lo, range, fbins, bins, scale, guard)], 1u);
            atomicAdd(&mine[hist_bin(v.w, lo, range, fbins, bins, scale, guard)], 1u);
            if (two) {
                atomicAdd(&mine[hist_bin(w.x, lo, range, fbins, bins, scale, guard)], 1u);
                atomicAdd(&mine[hist_bin(w.y, lo, range, fbins, bins, scale, guard)], 1u);
                atomicAdd(&mine[hist_bin(w.z, lo, range, fbins, bins, scale, guard)], 1u);
                atomicAdd(&mine[hist_bin(w.w, lo, range, fbins, bins, scale, guard)], 1u);
            }
        }
        for (uint64_t q = (n4 << 2) + i; q < n; q += stride)
            atomicAdd(&mine[hist_bin(row[q], lo, range, fbins, bins, scale, guard)], 1u);
    } else {
        for (uint64_t q = i; q < n; q += stride)
            atomicAdd(&mine[hist_bin(row[q], lo, range, fbins, bins, scale, guard)], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) {
        uint32_t c = 0;
        for (uint32_t s = 0; s < nsub; s++) c += sh[s * bins + b];
        if (c) atomicAdd(&counts[(size_t)bins * d + b], c);
    }
}

// ---------------------------------------------------------------------------
// The same three estimate passes over the reference's own layout: element (dimension d, sample w) at
// data[offset + ld*w + d] (column-major DIM x n, what GTXDatasetEngine receives, estimate.cu:22,109,165, and what
// the sampler's AoS mirror is).  A sample's coordinates are contiguous, so a warp reads them with lanes along d —
// 128-byte coalesced — and walks the samples; nothing is transposed first.  J = ceil(dims per CTA / 32) register
// slots per lane; blockIdx.y selects the chunk of 32*J dimensions.
// ---------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(256) k_minmax_aos(const float* __restrict__ data, uint64_t offset, uint64_t ld,
                                                    uint32_t dim, uint64_t n, uint32_t* __restrict__ mm) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t d0 = blockIdx.y * 32u * J;
    float lo[J], hi[J];
#pragma unroll
    for (int j = 0; j < J; j++) { lo[j] = INFINITY; hi[j] = -INFINITY; }
    const uint64_t stride = (uint64_t)gridDim.x * nw;
    const float* base = data + offset + d0 + lane;
    for (uint64_t w = (uint64_t)blockIdx.x * nw + warp; w < n; w += 4 * stride) {
        float v[4][J];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint64_t ww = w + r * stride;
#pragma unroll
            for (int j = 0; j < J; j++)
                v[r][j] = (ww < n && d0 + lane + 32u * j < dim) ? __ldcs(base + ld * ww + 32u * j) : NAN;
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int j = 0; j < J; j++) { lo[j] = fminf(lo[j], v[r][j]); hi[j] = fmaxf(hi[j], v[r][j]); }   // NaN ignored
    }
    __shared__ float slo[8][32 * J], shi[8][32 * J];
#pragma unroll
    for (int j = 0; j < J; j++) { slo[warp][lane + 32 * j] = lo[j]; shi[warp][lane + 32 * j] = hi[j]; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 32u * J; i += blockDim.x) {
        float a = INFINITY, b = -INFINITY;
        for (uint32_t q = 0; q < nw; q++) { a = fminf(a, slo[q][i]); b = fmaxf(b, shi[q][i]); }
        if (d0 + i < dim) { atomicMin(&mm[2 * (d0 + i)], f2ord(a)); atomicMax(&mm[2 * (d0 + i) + 1], f2ord(b)); }
    }
}

// shifted sums S1 = sum(x - p), S2 = sum((x - p)^2), pivot p = the dimension's first sample; fp32 over a strip of
// rows, double across strips / warps / CTAs (same arithmetic as k_moments_soa).  acc: dim x 2 doubles.
template <int J>
__global__ void __launch_bounds__(256) k_moments_aos(const float* __restrict__ data, uint64_t offset, uint64_t ld,
                                                     uint32_t dim, uint64_t n, double* __restrict__ acc) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t d0 = blockIdx.y * 32u * J;
    const float* base = data + offset + d0 + lane;
    float p[J];
    double s1[J], s2[J];
#pragma unroll
    for (int j = 0; j < J; j++) {
        p[j] = (d0 + lane + 32u * j < dim) ? __ldg(base + 32u * j) : 0.0f;
        s1[j] = 0.0; s2[j] = 0.0;
    }
    const uint64_t stride = (uint64_t)gridDim.x * nw;
    uint64_t w = (uint64_t)blockIdx.x * nw + warp;
    while (w < n) {
        float f1[J], f2[J];
#pragma unroll
        for (int j = 0; j < J; j++) { f1[j] = 0.f; f2[j] = 0.f; }
#pragma unroll 2
        for (int t = 0; t < 8 && w < n; t += 4) {
            float v[4][J];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint64_t ww = w + r * stride;
#pragma unroll
                for (int j = 0; j < J; j++)
                    v[r][j] = (ww < n && d0 + lane + 32u * j < dim) ? __ldcs(base + ld * ww + 32u * j) : p[j];
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int j = 0; j < J; j++) { const float a = v[r][j] - p[j]; f1[j] += a; f2[j] = fmaf(a, a, f2[j]); }
            w += 4 * stride;
        }
#pragma unroll
        for (int j = 0; j < J; j++) { s1[j] += (double)f1[j]; s2[j] += (double)f2[j]; }
    }
    __shared__ double a1[8][32 * J], a2[8][32 * J];
#pragma unroll
    for (int j = 0; j < J; j++) { a1[warp][lane + 32 * j] = s1[j]; a2[warp][lane + 32 * j] = s2[j]; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 32u * J; i += blockDim.x) {
        double x = 0.0, y = 0.0;
        for (uint32_t q = 0; q < nw; q++) { x += a1[q][i]; y += a2[q][i]; }
        if (d0 + i < dim) { atomicAdd(&acc[2 * (d0 + i)], x); atomicAdd(&acc[2 * (d0 + i) + 1], y); }
    }
}

// pivots for k_moments_finish when the data is AoS: xs[d*pitch] is replaced by data[offset + d]
__global__ void k_moments_finish_aos(uint32_t dim, uint64_t n, const float* __restrict__ data, uint64_t offset,
                                     const double* __restrict__ acc, int mode, float* __restrict__ out) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    const double p = (double)data[offset + d];
    const double m1 = acc[2 * d] / (double)n;
    double var = acc[2 * d + 1] / (double)n - m1 * m1;
    var = var < 0.0 ? 0.0 : var;
    out[d] = mode == 0 ? (float)(p + m1) : (mode == 1 ? (float)var : (float)sqrt(var));
}

// histogram over AoS data: the CTA keeps bins x (32*J) counters in (dynamic) shared memory — every dimension of its
// chunk at once — and walks the samples; one global atomic per non-empty bin per CTA at the end.  Same bin arithmetic
// as k_hist_soa (hist_bin): integer counts, bit-exact against the oracle.
template <int J>
__global__ void __launch_bounds__(512) k_hist_aos(const float* __restrict__ data, uint64_t offset, uint64_t ld,
                                                  uint32_t dim, uint64_t n, const float* __restrict__ limits, uint32_t bins,
                                                  uint32_t* __restrict__ counts) {
    // [bins][32*J]: the counter of (bin, local dimension dl) sits in bank dl % 32 = the lane that owns the dimension,
    // so the 32 shared-memory atomics of a warp instruction never collide, whatever the bins are
    extern __shared__ uint32_t sh[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t d0 = blockIdx.y * 32u * J;
    for (uint32_t i = threadIdx.x; i < 32u * J * bins; i += blockDim.x) sh[i] = 0;
    float lo[J], range[J], scale[J];
    const float fbins = (float)bins, guard = fbins - 0.5f;
#pragma unroll
    for (int j = 0; j < J; j++) {
        const uint32_t d = d0 + lane + 32u * j;
        lo[j] = d < dim ? limits[2 * d] : 0.f;
        range[j] = d < dim ? __fsub_rn(limits[2 * d + 1], lo[j]) : 1.f;
        scale[j] = __fdividef(fbins, range[j]);
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * nw;
    const float* base = data + offset + d0 + lane;
    for (uint64_t w = (uint64_t)blockIdx.x * nw + warp; w < n; w += 4 * stride) {
        float v[4][J];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint64_t ww = w + r * stride;
#pragma unroll
            for (int j = 0; j < J; j++)
                v[r][j] = (ww < n && d0 + lane + 32u * j < dim) ? __ldcs(base + ld * ww + 32u * j) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (w + r * stride >= n) break;
#pragma unroll
            for (int j = 0; j < J; j++)
                if (d0 + lane + 32u * j < dim)
                    atomicAdd(&sh[hist_bin(v[r][j], lo[j], range[j], fbins, bins, scale[j], guard) * (32u * J) + lane + 32u * j], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 32u * J * bins; i += blockDim.x) {
        const uint32_t b = i / (32u * J), dl = i - b * (32u * J), c = sh[i];
        if (c && d0 + dl < dim) atomicAdd(&counts[(size_t)bins * (d0 + dl) + b], c);
    }
}

// The same histogram WITHOUT shared-memory atomics.  ATOMS retires a conflict-free warp instruction only every ~24
// clocks, which capped k_hist_aos at a quarter of the HBM rate.  Here every WARP owns a private [bins][32] block of
// 16-bit counters and every LANE a fixed column of it, so an increment is a plain LDS / IADD / STS and no two threads
// ever touch the same counter.  R rows are in flight per lane: the R loads are issued back to back and each row stores
// count + (number of the R rows that fell into the same bin) — the same value from every row of an equal group, so
// the order of the stores does not matter.  16-bit counters double the resident warps (latency is what binds a
// kernel of 25 dependent instructions per row); the host keeps a CTA's rows <= 65535 so that none can overflow.
// A CTA works on ONE chunk of <= 32 dimensions (plan.first[c] .. plan.first[c+1] are the CTAs of chunk c, in numbers
// proportional to the chunk's cost).  A chunk narrower than 32 (the tail of D = 100, or a 1- or 2-dimensional model)
// packs 32 / tp rows into one warp instruction, tp = the next power of two >= its width.
struct HistPlan {
    uint32_t chunks;
    uint32_t first[34];     // CTA index where chunk c starts; first[chunks] = gridDim.x
};

// hist_bin in fixed point, for bins <= 256: q = round(t' * 4096) comes out of ONE FFMA against the rounding constant
// 1.5 * 2^23 (no F2I / FRND, which run on the quarter-rate XU pipe).  The fast path is taken when 0 < q < bins * 4096
// and q is not a multiple of 4096, i.e. t' is inside (0, bins) and at least 1 / 8192 = 1.2e-4 away from every
// integer — the approximate t' is within 1.8e-7 relative (4.6e-5 at the top bin) of the defining value, so
// floor(t') is the bin; everything else (1 value in 4096, NaN, out of range) takes the exact path
// (tests/test_hist_fixed_point.py restates this arithmetic in numpy and checks the claim).
constexpr float HIST_FX = 4096.f;
__device__ __forceinline__ bool hist_fx_plain(uint32_t q, uint32_t span) { return q < span && (q & 4095u) != 0u; }
__device__ __forceinline__ uint32_t hist_bin_fx(float x, float lo, float range, float fbins, uint32_t bins, float scale_fx) {
    const float d = __fsub_rn(x, lo);
    const uint32_t q = __float_as_uint(fmaf(d, scale_fx, 12582912.f)) - 0x4B400000u;
    if (__builtin_expect(hist_fx_plain(q, bins * 4096u), 1)) return q >> 12;
    return hist_bin_exact(d, range, fbins, bins);
}

__device__ __noinline__ uint32_t hist_bin_fx_call(float x, float lo, float range, float fbins, uint32_t bins, float scale_fx) {
    return hist_bin_fx(x, lo, range, fbins, bins, scale_fx);
}

template <int R, int G>   // R rows per shared-memory round, G rounds per load step: R * G rows in flight per lane
__global__ void __launch_bounds__(512) k_hist_aos_lanes(const float* __restrict__ data, uint64_t offset, uint64_t ld,
                                                        uint32_t dim, uint64_t n, const float* __restrict__ limits,
                                                        uint32_t bins, const __grid_constant__ HistPlan plan,
                                                        uint32_t* __restrict__ counts) {
    constexpr int RL = R * G;
    extern __shared__ uint32_t sh[];
    uint16_t* sh16 = reinterpret_cast<uint16_t*>(sh);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t chunk = 0;
    while (chunk + 1 < plan.chunks && blockIdx.x >= plan.first[chunk + 1]) chunk++;
    const uint32_t slot = blockIdx.x - plan.first[chunk], nslots = plan.first[chunk + 1] - plan.first[chunk];
    const uint32_t d0 = 32u * chunk, width = min(32u, dim - d0);
    uint32_t tp = 1;
    while (tp < width) tp <<= 1;
    const uint32_t rpw = 32u / tp, rsub = lane / tp, dl = lane & (tp - 1);
    const bool active = dl < width;
    {
        uint4* z = reinterpret_cast<uint4*>(sh);
        for (uint32_t i = threadIdx.x; i < nw * bins * 4u; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    const uint32_t d = d0 + (active ? dl : 0u);
    const float lo = limits[2 * d], range = __fsub_rn(limits[2 * d + 1], lo);
    const float fbins = (float)bins, scale_fx = __fdividef(fbins * HIST_FX, range);
    const uint32_t span = bins * 4096u;
    __syncthreads();
    uint16_t* mine = sh16 + warp * bins * 32u + lane;
    // rows of this CTA: [r_begin, r_begin + rows), rows <= 65535 and rows * ld * 4 < 2^32 (the host sees to both).
    // A warp step covers `step` = RL * rpw consecutive rows; every warp owns a CONTIGUOUS run of full steps (so that
    // it can prefetch its own stream), and the last warp also takes the ragged rest, row by row.
    const uint64_t r_begin = n * slot / nslots;
    const uint32_t rows = (uint32_t)(n * (slot + 1) / nslots - r_begin);
    const uint32_t step = (uint32_t)RL * rpw, full = rows / step;
    const uint32_t per_warp = (full + nw - 1) / nw;
    const uint32_t s_begin = min(full, warp * per_warp), s_end = min(full, s_begin + per_warp);
    const char* cta = reinterpret_cast<const char*>(data + offset + ld * r_begin + d0);   // row 0, first dimension of the chunk
    const uint32_t ldb = (uint32_t)ld * 4u;                            // bytes per row
    const uint32_t pitch = ldb * rpw, hop = ldb * step;                // a lane's next row / the next step
    auto count_rows = [&](const float* x) {     // R rows of this lane -> its private counters
        uint32_t q[R], b[R], c[R];
        bool plain = true;
#pragma unroll
        for (int k = 0; k < R; k++) {
            q[k] = __float_as_uint(fmaf(__fsub_rn(x[k], lo), scale_fx, 12582912.f)) - 0x4B400000u;
            plain = plain && hist_fx_plain(q[k], span);
            b[k] = (q[k] >> 12) * 32u;
        }
        if (__builtin_expect(!plain, 0)) {      // a value at a bin edge, outside the limits or NaN: the exact rule for all R
#pragma unroll
            for (int k = 0; k < R; k++) b[k] = hist_bin_fx_call(x[k], lo, range, fbins, bins, scale_fx) * 32u;
        }
#pragma unroll
        for (int k = 0; k < R; k++) c[k] = mine[b[k]];
#pragma unroll
        for (int k = 0; k < R; k++) {
            uint32_t same = 1;
#pragma unroll
            for (int j = 0; j < R; j++)
                if (j != k) same += (b[j] == b[k]) ? 1u : 0u;
            c[k] += same;
        }
#pragma unroll
        for (int k = 0; k < R; k++) mine[b[k]] = (uint16_t)c[k];
    };
    if (active && s_begin < s_end) {
        float v[RL], vn[RL];
        uint32_t off = hop * s_begin + ldb * rsub + 4u * dl;           // this lane's element of its first row
        const uint32_t off_last = hop * (s_end - 1) + ldb * rsub + 4u * dl;
        uint32_t pk[RL];                                                // the lane's RL rows inside a step
#pragma unroll
        for (int k = 0; k < RL; k++) pk[k] = pitch * k;
#pragma unroll
        for (int k = 0; k < RL; k++) v[k] = __ldcs(reinterpret_cast<const float*>(cta + (off + pk[k])));
        // The loads of one step ahead keep only a few KB in flight per SM, so the warp pulls its stream into L2 well
        // ahead: every PFB steps, one lane per row, the first and the last line of the segments of the next block.
        constexpr uint32_t PFB = 4;                                     // steps per prefetch block
        const uint32_t pf_rows = PFB * step;                            // rows per block
        const uint32_t pf_ahead = max(2u, 16u / rpw / PFB) * PFB;       // steps of lead
        uint32_t until_pf = 0;
        for (uint32_t s = s_begin; s < s_end; s++) {
            off = min(off + hop, off_last);                             // the last step re-reads itself (unused)
#pragma unroll
            for (int k = 0; k < RL; k++) vn[k] = __ldcs(reinterpret_cast<const float*>(cta + (off + pk[k])));
            if (until_pf == 0) {
                until_pf = PFB;
                if (s + pf_ahead + PFB <= s_end) {
                    const char* blk = cta + (uint64_t)hop * (s + pf_ahead);
                    for (uint32_t i = lane; i < pf_rows; i += 32) {
                        const char* q = blk + i * ldb;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 4u * (width - 1)));
                    }
                }
            }
            until_pf--;
#pragma unroll
            for (int g = 0; g < G; g++) count_rows(v + g * R);
#pragma unroll
            for (int k = 0; k < RL; k++) v[k] = vn[k];
        }
    }
    if (active && warp == nw - 1) {                                     // the ragged rest: fewer than `step` rows
        for (uint32_t r = full * step + rsub; r < rows; r += rpw) {
            const float x = __ldcs(reinterpret_cast<const float*>(cta + ((uint64_t)r * ldb + 4u * dl)));
            mine[hist_bin_fx(x, lo, range, fbins, bins, scale_fx) * 32u] += 1;
        }
    }
    __syncthreads();
    // fold the warps (and, in a narrow chunk, the lanes that served the same dimension); one global atomic per non-empty bin
    for (uint32_t i = threadIdx.x; i < bins * tp; i += blockDim.x) {
        const uint32_t bin = i / tp, x = i & (tp - 1);
        if (x >= width) continue;
        uint32_t total = 0;
        for (uint32_t w = 0; w < nw; w++)
            for (uint32_t q = 0; q < rpw; q++) total += sh16[(w * bins + bin) * 32u + q * tp + x];
        if (total) atomicAdd(&counts[(size_t)bins * (d0 + x) + bin], total);
    }
}

// uint_to_real (estimate.cu:34-44): pdf = (alpha / (hi-lo)) * count
__global__ void k_uint_to_real(uint32_t bins, uint32_t dim, float alpha,
                               const float* __restrict__ limits,
                               const uint32_t* __restrict__ counts, float* __restrict__ pdf) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t d = blockIdx.y;
    if (b < bins && d < dim) {
        const float w = __fsub_rn(limits[2 * d + 1], limits[2 * d]);
        pdf[(size_t)bins * d + b] = __fmul_rn(__fdiv_rn(alpha, w), (float)counts[(size_t)bins * d + b]);
    }
}

// Bin ranking (replaces bitonic_local, estimate.cu:117-147): bins of one dimension ordered by decreasing mass, the
// input of hdi-rank-count (C/util.clj:52-65).  ONE WARP per dimension keeps the whole column in registers: entry
// e = lane + 32*slot (slot < bins/32) lives in register `slot` of `lane`, so a comparator of span < 32 is a lane
// shuffle and one of span >= 32 a register-to-register exchange inside the lane — no shared memory, no block
// barrier (the reference stages the column in shared memory behind two __syncthreads per comparator level).
// The comparator schedule is the classic bitonic one (runs of length 2, 4, .., bins; spans run/2 .. 1), which is
// what fixes the order INSIDE groups of equal mass: lower entry p and upper entry q = p + span exchange iff
//     (mass[p] < mass[q]) != ascending(p, run),   ascending(p, run) = bit `run` of p,
// i.e. descending runs leave ties in place and ascending runs exchange them.  The oracle simulates the reference's
// network, so equal-mass groups (empty bins, mostly) come out in the reference's order.
template <int SLOTS>
__global__ void __launch_bounds__(32) k_bin_ranks(const float* __restrict__ pdf, float* __restrict__ ranks) {
    constexpr uint32_t bins = 32u * SLOTS;       // compile-time, so that every register index below is static
    const uint32_t lane = threadIdx.x;
    const float* col = pdf + (size_t)blockIdx.x * bins;
    float mass[SLOTS];
    uint32_t bin[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; s++) { bin[s] = lane + 32u * s; mass[s] = col[bin[s]]; }
#pragma unroll
    for (uint32_t run = 2; run <= bins; run <<= 1) {
#pragma unroll
        for (uint32_t span = run >> 1; span >= 1; span >>= 1) {
            if (span >= 32u) {
                const int ds = (int)(span >> 5);             // partner sits ds registers away in the same lane
#pragma unroll
                for (int s = 0; s < SLOTS; s++) {
                    if ((s & ds) || s + ds >= SLOTS) continue;   // visit each pair once, from its lower entry
                    const uint32_t p = lane + 32u * s;
                    const bool asc = (p & run) != 0u;
                    if ((mass[s] < mass[s + ds]) != asc) {
                        const float tm = mass[s]; mass[s] = mass[s + ds]; mass[s + ds] = tm;
                        const uint32_t tb = bin[s]; bin[s] = bin[s + ds]; bin[s + ds] = tb;
                    }
                }
            } else {
                const bool upper = (lane & span) != 0u;
#pragma unroll
                for (int s = 0; s < SLOTS; s++) {
                    const float om = __shfl_xor_sync(0xffffffffu, mass[s], span);
                    const uint32_t ob = __shfl_xor_sync(0xffffffffu, bin[s], span);
                    const uint32_t p = (lane & ~span) + 32u * s;           // the pair's lower entry
                    const bool asc = (p & run) != 0u;
                    const float lo_m = upper ? om : mass[s], hi_m = upper ? mass[s] : om;
                    if ((lo_m < hi_m) != asc) { mass[s] = om; bin[s] = ob; }
                }
            }
        }
    }
    float* out = ranks + (size_t)blockIdx.x * bins;
#pragma unroll
    for (int s = 0; s < SLOTS; s++) out[lane + 32u * s] = (float)bin[s];
}

// --------------------------------------------------------- mean / variance --
// mean_reduce + variance_reduce (estimate.cu:149-187) fused into ONE pass over
// the data: shifted sums S1 = sum(x-p), S2 = sum((x-p)^2) with pivot p = x[d,0]
// (the shift removes the cancellation of the naive one-pass formula), fp32
// inside a thread strip, double across threads/blocks.  acc: dim x 2 doubles.
__global__ void k_moments_soa(const float* __restrict__ xs, uint64_t pitch, uint64_t n,
                              double* __restrict__ acc) {
    const uint32_t d = blockIdx.y;
    const float* row = xs + (size_t)d * pitch;
    const float p = __ldg(row);
    double s1 = 0.0, s2 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((pitch & 3) == 0) && ((((uintptr_t)xs) & 15) == 0);
    if (vec) {
        // every CTA streams ONE contiguous segment of the row (sequential 4 KB blocks: DRAM pages stay open)
        const uint64_t n4 = n >> 2, seg = (n4 + gridDim.x - 1) / gridDim.x;
        const uint64_t q0 = (uint64_t)blockIdx.x * seg, q1 = q0 + seg < n4 ? q0 + seg : n4;
        const float4* row4 = reinterpret_cast<const float4*>(row);
        const float4 pad = make_float4(p, p, p, p);
        for (uint64_t q = q0 + threadIdx.x; q < q1; q += 8 * blockDim.x) {
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = (q + r * blockDim.x < q1) ? __ldcs(row4 + q + r * blockDim.x) : pad;
            float f1 = 0.f, f2 = 0.f;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const float a = v[r].x - p, b = v[r].y - p, c = v[r].z - p, e = v[r].w - p;
                f1 += (a + b) + (c + e);
                f2 += (a * a + b * b) + (c * c + e * e);
            }
            s1 += (double)f1; s2 += (double)f2;
        }
        for (uint64_t t = (n4 << 2) + i; t < n; t += stride) { const float a = row[t] - p; s1 += a; s2 += (double)a * a; }
    } else {
        for (uint64_t t = i; t < n; t += stride) { const float a = row[t] - p; s1 += a; s2 += (double)a * a; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ double a1[32], a2[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { a1[warp] = s1; a2[warp] = s2; }
    __syncthreads();
    if (warp == 0) {
        s1 = lane < nw ? a1[lane] : 0.0;
        s2 = lane < nw ? a2[lane] : 0.0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) { atomicAdd(&acc[2 * d], s1); atomicAdd(&acc[2 * d + 1], s2); }
    }
}

// mode 0: mean, 1: population variance, 2: sd
__global__ void k_moments_finish(uint32_t dim, uint64_t n, const float* __restrict__ xs, uint64_t pitch,
                                 const double* __restrict__ acc, int mode, float* __restrict__ out) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dim) return;
    const double p = (double)xs[(size_t)d * pitch];
    const double m1 = acc[2 * d] / (double)n;
    double var = acc[2 * d + 1] / (double)n - m1 * m1;
    var = var < 0.0 ? 0.0 : var;
    out[d] = mode == 0 ? (float)(p + m1) : (mode == 1 ? (float)var : (float)sqrt(var));
}

// ------------------------------------------------- accu-path finalisation --
// sum_means_vertical + scal! (mcmc-stretch.cu:250-260, nvidia_gtx.clj:457-460)
// folded into one launch per step: means[step*dim + i] = factor * sum_b blk[i*G + b].
// One warp per dimension, fixed lane-strided order + shuffle tree (deterministic).
__global__ void k_step_means(uint32_t dim, uint32_t G, const float* __restrict__ blk_sums,
                             float factor, float* __restrict__ means_step) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= dim) return;
    float s = 0.f;
    for (uint32_t b = lane; b < G; b += 32) s += blk_sums[(size_t)warp * G + b];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) means_step[warp] = s * factor;
}

// sum_accept_reduce / sum_accept_reduction (mcmc-stretch.cu:240-248): G uint32 -> 1 uint64
__global__ void k_accept_total(uint32_t G, const uint32_t* __restrict__ accept,
                               unsigned long long* __restrict__ total) {
    unsigned long long s = 0;
    for (uint32_t b = threadIdx.x; b < G; b += blockDim.x) s += accept[b];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ unsigned long long sm[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) sm[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < nw; w++) t += sm[w];
        *total = t;
    }
}

// ------------------------------------------------------------------- acor --
// Goodman's acor as the reference runs it (acor.cu:17-168; host
// nvidia_gtx.clj:230-278) without dynamic parallelism: one CTA per dimension
// performs mean subtraction, the (c0, d) pass and the halve-and-retry loop.
// series: dim x n column-major work copy (modified in place).
// The lag window follows the reference's shared-memory staging rule: a window
// element in the same WGS-block as t reads 0 when u+lag >= n, one in the next
// block is read raw (see oracle/bayadera_oracle.c acor_pass).
__device__ __forceinline__ double block_sum_d(double v, double* sm) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (uint32_t w = 0; w < nw; w++) t += sm[w];
    return t;
}

__global__ void k_acor(uint32_t dim, uint32_t n, uint32_t wgs, uint32_t lag, uint32_t min_lag,
                       uint32_t win_mult, float* __restrict__ series, float* __restrict__ tau,
                       float* __restrict__ mean, float* __restrict__ sigma) {
    __shared__ double sm[32];
    __shared__ double sm2[32];
    const uint32_t d = blockIdx.x;
    if (d >= dim) return;
    // mean and subtraction (sum_reduce_horizontal + subtract_mean, acor.cu:5-28)
    double s = 0.0;
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) s += (double)series[(size_t)t * dim + d];
    s = block_sum_d(s, sm);
    const float mu = (float)(s / (double)n);
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) series[(size_t)t * dim + d] -= mu;
    __syncthreads();

    uint32_t lag2 = lag, n2 = n, stride = 1;
    // c0: first level's (used by the final tau, acor.cu:160); c0_prev: the previous level's, which the
    // reference's loop body re-declares and divides the new level's d by (acor.cu:147, 156).
    float c0 = 0.f, c0_prev = 0.f, dv = 0.f, tau_d = 0.f;
    bool first = true;
    while (true) {
        const uint32_t blk = wgs < n2 ? wgs : n2;
        const size_t st = (size_t)stride * dim;
        double pc0 = 0.0, pd = 0.0;
        for (uint32_t t = threadIdx.x; t + lag2 < n2; t += blockDim.x) {
            const float xt = series[t * st + d];
            float xacc = 0.f;
            const uint32_t tb = t / blk;
            for (uint32_t q = 1; q <= lag2; q++) {
                const uint32_t u = t + q;
                const bool zero = ((u / blk) == tb) && !(u + lag2 < n2);
                xacc += zero ? 0.f : series[u * st + d];
            }
            pc0 += (double)__fmul_rn(xt, xt);
            pd += (double)__fmul_rn(xt, __fadd_rn(xt, __fmul_rn(2.0f, xacc)));
        }
        const float c0v = (float)block_sum_d(pc0, sm);
        dv = (float)block_sum_d(pd, sm2);
        if (first) { c0 = c0v; c0_prev = c0v; first = false; }
        tau_d = dv / c0_prev;
        c0_prev = c0v;
        if (!((min_lag < lag2) && ((float)lag2 < tau_d * (float)win_mult))) break;
        n2 /= 2;
        lag2 = (lag * win_mult < n2) ? lag : max(10u, n2 / win_mult);
        __syncthreads();
        for (uint32_t g = threadIdx.x; g < n2; g += blockDim.x)
            series[(size_t)(2 * g) * st + d] += series[(size_t)(2 * g + 1) * st + d];
        stride *= 2;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float scale = (float)stride * (float)(n2 - lag2);
        tau[d] = dv * (float)(n - lag) / (scale * c0);
        sigma[d] = sqrtf(dv / (scale * (float)n));
        mean[d] = mu;
    }
}

// ------------------------------------------------------ GLM data preparation --
// Rows [y, x_1..x_D] (stride D+1, as the model packs them into `params`) are split once
// into Xmat (rows x D row-major, 16-byte aligned rows for cp.async / TMA) and sy = X^T y.
__global__ void k_glm_repack(const float* __restrict__ data, uint64_t rows, uint32_t dim,
                             float* __restrict__ xmat) {
    const uint64_t total = rows * dim;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t r = e / dim;
        const uint32_t i = (uint32_t)(e - r * dim);
        xmat[e] = data[r * (dim + 1) + 1 + i];
    }
}

// sy[i] += sum_r y_r * x_{r,i}; one CTA per row chunk, thread = dimension (dim <= blockDim.x),
// double accumulation, one atomicAdd(double) per (chunk, dim).
// sx[i] += sum_r x_{r,i} (column sums; used by the tensor-core path, may be NULL).
__global__ void k_glm_xty(const float* __restrict__ data, uint64_t rows, uint32_t dim,
                          uint64_t rows_per_block, double* __restrict__ sy, double* __restrict__ sx) {
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_block;
    const uint64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
        double s = 0.0, c = 0.0;
        for (uint64_t r = r0; r < r1; r++) {
            const float* row = data + r * (dim + 1);
            const double x = (double)row[1 + i];
            s += (double)row[0] * x;
            c += x;
        }
        atomicAdd(&sy[i], s);
        if (sx) atomicAdd(&sx[i], c);
    }
}

// sp[k] = sum over row chunks of partial[c*H + k], fixed order (deterministic)
__global__ void k_glm_finish(uint32_t H, uint32_t chunks, const double* __restrict__ partial,
                             double* __restrict__ sp) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= H) return;
    double s = 0.0;
    for (uint32_t c = 0; c < chunks; c++) s += partial[(size_t)c * H + k];
    sp[k] = s;
}

// generic row-additive path: one warp per walker sums its row-chunk partials — lanes stride over the chunks, then a
// fixed shuffle tree (deterministic: the same order on every run and every rank)
__global__ void k_rowadd_finish(uint32_t H, uint32_t chunks, const double* __restrict__ partial, double* __restrict__ sp) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= H) return;
    double s = 0.0;
    for (uint32_t c = lane; c < chunks; c += 32) s += partial[(size_t)c * H + k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sp[k] = s;
}

// same with an explicit leading dimension of the partial matrix
__global__ void k_glm_finish_ld(uint32_t n, uint32_t chunks, uint32_t ldp, const double* __restrict__ partial,
                                double* __restrict__ sp) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s = 0.0;
    for (uint32_t c = 0; c < chunks; c++) s += partial[(size_t)c * ldp + k];
    sp[k] = s;
}

// ---------------------------------------------------------------------------
// HDI (C/util.clj:52-100): hdi-rank-count + hdi-bins + hdi-regions of every histogram column,
// one CTA per dimension.  pdf / ranks: bins x dim column-major, limits: 2 x dim.  The mass accumulation is the
// reference's sequential double loop (one thread; 256 shared-memory reads); asum is an fp32 accumulation like the
// reference's BLAS call (sequential order here).
//   counts[d]   = smallest number of ranked bins whose mass reaches mass * sum(pdf)   (or forced[d] if >= 0)
//   regions     = [lo0 hi0 lo1 hi1 ...] of the maximal runs of selected bins, 2 * max_regions floats per dimension
// dynamic shared memory: 3 * bins words
// ---------------------------------------------------------------------------
__global__ void k_hdi(uint32_t bins, const float* __restrict__ limits, const float* __restrict__ pdf,
                      const float* __restrict__ ranks, double mass, const int32_t* __restrict__ forced,
                      int32_t* __restrict__ counts, int32_t* __restrict__ nregions, float* __restrict__ regions,
                      uint32_t max_regions) {
    extern __shared__ float hdi_sm[];
    float* p = hdi_sm;
    uint32_t* rk = reinterpret_cast<uint32_t*>(hdi_sm + bins);
    uint32_t* sel = rk + bins;
    __shared__ uint32_t cnt_s;
    const uint32_t d = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x) {
        p[i] = pdf[(size_t)d * bins + i];
        const uint32_t r = (uint32_t)ranks[(size_t)d * bins + i];
        rk[i] = r < bins ? r : bins - 1;
        sel[i] = 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t cnt = 0;
        if (forced && forced[d] >= 0) {
            cnt = (uint32_t)forced[d] < bins ? (uint32_t)forced[d] : bins;
        } else {
            float total = 0.0f;   // Neanderthal asum of a float vector: fp32 accumulation (sequential here)
            for (uint32_t i = 0; i < bins; i++) total = __fadd_rn(total, fabsf(p[i]));
            const double density = mass * (double)total;
            double acc = 0.0;
            while (cnt < bins && acc < density) acc += (double)p[rk[cnt++]];
        }
        cnt_s = cnt;
        counts[d] = (int32_t)cnt;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cnt_s; i += blockDim.x) sel[rk[i]] = 1u;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double lower = (double)limits[2 * d], upper = (double)limits[2 * d + 1];
        const double bw = (upper - lower) / (double)bins;
        uint32_t n = 0;
        for (uint32_t b = 0; b < bins; b++) {
            if (!sel[b]) continue;
            const uint32_t start = b;
            while (b + 1 < bins && sel[b + 1]) b++;
            if (n < max_regions) {
                regions[((size_t)d * max_regions + n) * 2] = (float)(lower + bw * (double)start);
                regions[((size_t)d * max_regions + n) * 2 + 1] = (float)(lower + bw * ((double)b + 1.0));
            }
            n++;
        }
        nregions[d] = (int32_t)n;
    }
}

// ---------------------------------------------------------------------------
// Multi-GPU mode A over peer memory.  PeerTable mirrors bay_peers_t of the NVRTC program (stretch_program.inc).
// ---------------------------------------------------------------------------
struct PeerTable {
    unsigned long long base[8];
    unsigned long long lp_off, xa_off;
    uint32_t n, self;
};

// Barrier between half-steps: lane r publishes `epoch` into slot [rank] of rank r's flags (after a system-scope
// fence, so the stretch kernel's peer stores are ordered before it) and waits for slot [r] of its own flags.
__global__ void k_peer_barrier(PeerTable t, uint64_t flags_off, uint32_t rank, uint32_t epoch) {
    const uint32_t r = threadIdx.x;
    if (r >= t.n) return;
    __threadfence_system();
    uint32_t* remote = reinterpret_cast<uint32_t*>(t.base[r]) + flags_off + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(t.base[rank]) + flags_off + r;
    uint32_t seen;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    } while ((int32_t)(seen - epoch) < 0);
}

}  // namespace bay
