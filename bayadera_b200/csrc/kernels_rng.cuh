// bayadera_b200 — direct (i.i.d.) samplers and density post-processing.
//
// Replaces K/cuda/rng/{uniform,gaussian,exponential,erlang}-sampler.cu and the exp / evidence parts of
// K/cuda/engines/nvidia-gtx-{distribution,likelihood}.cu
// (K = /root/reference/src/device/uncomplicate/bayadera/internal/device/cuda).
// Philox key {seed, 0xdecafaaa}, counter {gid, 0xf00dcafe, 0xdeadbeef, 0xbeeff00d}; 4 variates per thread,
// stored as one float4.  The reference builds these kernels with -use_fast_math, so the transcendental steps use
// the same approximate units here (MUFU sin/cos/lg2/sqrt/rcp) — the reference's goldens embed them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace bay {

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

enum DirectFamily { DIRECT_UNIFORM = 0, DIRECT_GAUSSIAN = 1, DIRECT_EXPONENTIAL = 2, DIRECT_ERLANG = 3 };

// n4 = number of float4 outputs; p0, p1 = the family's parameters
__global__ void k_direct_sample(int family, uint32_t n4, uint32_t seed, float p0, float p1, float4* __restrict__ x) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n4) return;
    uint32_t r[4];
    float4 out;
    if (family == DIRECT_ERLANG) {
        // erlang-sampler.cu:30-50: sum of k log-uniforms, counter word 3 = draw index, divided by -lambda
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (uint32_t i = 0; (float)i < p1; i++) {
            philox4x32_10(gid, 0xf00dcafeu, 0xdeadbeefu, i, seed, 0xdecafaaau, r);
            a0 += __logf(u01(r[0])); a1 += __logf(u01(r[1])); a2 += __logf(u01(r[2])); a3 += __logf(u01(r[3]));
        }
        const float nl = -p0;
        out = make_float4(__fdividef(a0, nl), __fdividef(a1, nl), __fdividef(a2, nl), __fdividef(a3, nl));
    } else {
        philox4x32_10(gid, 0xf00dcafeu, 0xdeadbeefu, 0xbeeff00du, seed, 0xdecafaaau, r);
        const float u0 = u01(r[0]), u1 = u01(r[1]), u2 = u01(r[2]), u3 = u01(r[3]);
        if (family == DIRECT_UNIFORM) {
            // uniform-sampler.cu:32-40: u * (upper - lower) + lower, contracted
            const float range = __fsub_rn(p1, p0);
            out = make_float4(__fmaf_rn(u0, range, p0), __fmaf_rn(u1, range, p0), __fmaf_rn(u2, range, p0),
                              __fmaf_rn(u3, range, p0));
        } else if (family == DIRECT_GAUSSIAN) {
            // gaussian-sampler.cu:16-23 (Box-Muller) and :48-51 (x * sigma + mu, contracted)
            const float two_pi = 6.2831855f;
            const float ra = sqrt_approx(__fmul_rn(-2.0f, __logf(u1))), rb = sqrt_approx(__fmul_rn(-2.0f, __logf(u3)));
            const float ta = __fmul_rn(two_pi, u0), tb = __fmul_rn(two_pi, u2);
            out = make_float4(__fmaf_rn(__fmul_rn(__sinf(ta), ra), p1, p0), __fmaf_rn(__fmul_rn(__cosf(ta), ra), p1, p0),
                              __fmaf_rn(__fmul_rn(__sinf(tb), rb), p1, p0), __fmaf_rn(__fmul_rn(__cosf(tb), rb), p1, p0));
        } else {
            // exponential-sampler.cu:36-39: -1/lambda * log(1 - u)
            const float s = __fdividef(-1.0f, p0);
            out = make_float4(__fmul_rn(s, __logf(__fsub_rn(1.0f, u0))), __fmul_rn(s, __logf(__fsub_rn(1.0f, u1))),
                              __fmul_rn(s, __logf(__fsub_rn(1.0f, u2))), __fmul_rn(s, __logf(__fsub_rn(1.0f, u3))));
        }
    }
    x[gid] = out;
}

// pdf / lik kernels of the reference = exp of the log kernel (fast-math exp there)
__global__ void k_exp_inplace(uint32_t n, float* __restrict__ v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = __expf(v[i]);
}

// evidence_reduce + sum_reduction (likelihood.cu:24-36): sum_i exp(loglik_i) accumulated in double.  One partial per
// CTA (acc[blockIdx.x]), added up by the host in CTA order: the same bits on every run, no floating-point atomics.
__global__ void k_exp_sum(uint32_t n, const float* __restrict__ v, double* __restrict__ acc) {
    double s = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (double)__expf(v[i]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sm[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) sm[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (uint32_t w = 0; w < nw; w++) t += sm[w];
        acc[blockIdx.x] = t;
    }
}

}  // namespace bay
