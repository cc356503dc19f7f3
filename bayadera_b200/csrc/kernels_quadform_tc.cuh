// bayadera_b200 — stretch move of Gaussian / quadratic-form models on the tensor cores (tcgen05 / TMEM, sm_100a).
//
// For a model flagged BAY_MODEL_QUADFORM the log-density is  logp(x) = -1/2 |U (x - mu)|^2  with
// params = [mu (D) | U (D x D, row-major)].  The generic kernel (stretch_program.inc) evaluates it per thread:
// a D x D matrix-vector product with no operand reuse — on BASELINE config 5 (D = 100, 2^20 walkers) that is 5 200
// FMAs per walker fed by one uniform load per two FMAs, and the kernel reaches 0.19 of the HBM roofline although
// the ensemble stream (419 MB per half-step) is all that has to move.  Across a TILE of 128 walkers the same work is
// the dense contraction  R(128 x D) = C(128 x D) . U^T(D x D),  C = centred proposals — here on the 5th-generation
// tensor cores, so that HBM is the limiter again.  There is no counterpart in the reference (its LOGFN contract is
// one thread per walker, K/cuda/engines/nvidia-gtx-mcmc-stretch.cu:78-89); the arithmetic contract is the
// model's serial LOGFN in the oracle, the proposal arithmetic is bit-identical to bay_stretch_move.
//
// Layout: only the AoS mirror of the ensemble is touched (row k = walker k, DA = D rounded up to 4 floats): the
// 128 own rows of a tile are one contiguous block and a partner row is one contiguous 4*DA-byte read — from this
// GPU's mirror or, under the walker-partitioned multi-GPU mode, PULLED over NVLink from the rank that owns walker
// j (a warp reads the row as contiguous 16-byte pieces, so it crosses the link as full lines).  The SoA matrix is
// rebuilt lazily before a read-out (bay_sampler::soa_own_stale).
//
// fp32-level accuracy on fp16 tensor-core inputs: every centred proposal row is scaled by a power of two into
// [1/2, 1) (exact), split c = hi + lo into two fp16 values (22 significant bits), U likewise per row; the product is
// hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM (the dropped lo.lo is 2^-22 relative) and the scales are
// undone in the epilogue.  Measured against fp64 on config 5: 3e-7 relative on logp — tighter than the serial fp32
// evaluation itself (4e-7); a bf16 split gives 1e-5.
//
// One persistent CTA per SM, 512 threads = 16 warps, software-pipelined over tiles:
//     iteration t:  Philox(t+1) -> issue the global loads of tile t+1 (own rows + partner rows, 64 registers)
//                   epilogue(t): wait MMA(t), TMEM -> sum of squares -> accept -> write accepted rows back
//                   convert(t+1): proposal, centre, scale, split -> swizzled fp16 operand tiles in shared memory
//                   one thread issues MMA(t+1) (3 x D/16 tcgen05.mma into the other TMEM stage)
// so the HBM / NVLink latency of tile t+1 hides behind the epilogue of tile t.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "kernels_glm_tc.cuh"   // mbarrier / tcgen05 PTX helpers

namespace bay {
namespace qf {

constexpr int TILE = 128;                        // walkers per tile = MMA M
constexpr int KD = 64;                           // K extent of one operand tile (128 B of fp16 = one swizzle row)
constexpr int MAX_D = 128;                       // two K chunks, N <= 128
constexpr int THREADS = 512;
constexpr int ROWS_PER_WARP = TILE / (THREADS / 32);   // 8
constexpr uint32_t A_CHUNK_BYTES = TILE * KD * 2;      // 16 KB: one K chunk of one fp16 plane of the walker tile

struct Args {
    const float* mu;                 // D
    const float* U;                  // D x D row-major
    float* xa_active;                // this rank's mirror of the ACTIVE half: row k at xa_active + k*DA
    unsigned long long xa_compl[8];  // per rank: base of the COMPLEMENTARY half's mirror (row j at + j*DA)
    uint32_t hs;                     // walkers of a half owned by one rank (partner j lives on rank j / hs); K on one GPU
    float* lp_active;                // log-densities of the active half
    uint32_t K;                      // walkers per half (Philox partner range)
    uint32_t k_begin, k_end;         // this launch's walkers of the active half
    uint32_t D, DA;                  // dimension, mirror row stride (floats)
    uint32_t seed, tag, step;
    float cA, cB, cC, beta;
    unsigned long long* accepted;    // optional counter of accepted moves (diagnostics), may be NULL
};

__host__ __device__ constexpr uint32_t np_of(uint32_t D) { return (D + 15u) / 16u * 16u; }   // padded N = padded K

__host__ __device__ constexpr size_t smem_bytes(bool bulk) {
    // A: 2 planes x 2 chunks x 16 KB; B: 2 planes x 2 chunks x (128 rows x 128 B); per-walker arrays; slack
    // bulk: + the own rows of the next tile, one bulk copy per warp (128 rows x up to 512 B).  The walker-partitioned
    // instantiation does without: every KB of shared memory is taken from L1, and in-flight REMOTE loads are staged
    // there — with 205 KB of shared memory the same kernel pulled 16 % slower over NVLink (152 vs 128 us at 2 GPUs).
    return 1024 + 4 * (size_t)A_CHUNK_BYTES + 4 * (size_t)A_CHUNK_BYTES + 8192 + (bulk ? (size_t)TILE * MAX_D * 4 : 0);
}

// byte offset of element (row r, k' < 64) inside a K-major SWIZZLE_128B operand tile (8-row atoms of 1024 B)
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t k) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((((k >> 3) ^ r) & 7u) << 4) + (k & 7u) * 2u;
}

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Power of two s with |m| * s in [1/2, 1) and inv = 1 / s, branch-free on the exponent field e of m
// (m in [2^(e-127), 2^(e-126))).  e is clamped to [1, 252]: m = 0 gives a harmless huge s (0 * s = 0); an infinite,
// NaN or astronomically large m leaves the product non-finite or makes inv^2 overflow, so the move is rejected.
__device__ __forceinline__ void pow2_scale(float m, float* s, float* inv) {
    uint32_t e = (__float_as_uint(m) >> 23) & 0xffu;
    e = min(max(e, 1u), 252u);
    *s = __uint_as_float((253u - e) << 23);       // 2^(126-e)
    *inv = __uint_as_float((e + 1u) << 23);       // 2^(e-126)
}

// shared-memory accesses by 32-bit shared address (the operand tiles are addressed by computed byte offsets)
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t x) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(x) : "memory");
}
__device__ __forceinline__ void sts128z(uint32_t addr) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void stsf(uint32_t addr, float x) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}
__device__ __forceinline__ float ldsf(uint32_t addr) {
    float x;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr) : "memory");
    return x;
}
__device__ __forceinline__ float4 ldsf4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// BULK: the own rows of a tile travel by one cp.async.bulk per warp into shared memory (+2 % on one GPU).  The walker-
// partitioned sampler instantiates BULK = false — own rows prefetched into L2, read by ld.global.cg when converted —
// because the extra 64 KB of shared memory cost it NVLink throughput (see smem_bytes).
template <bool BULK>
__global__ void __launch_bounds__(THREADS, 1) k_quadform_move_tc(const __grid_constant__ Args a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // everything below is addressed by 32-bit SHARED addresses (no generic pointers: those cost a descriptor move per
    // access); the operand tiles need 1024-byte alignment
    const uint32_t smem = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t D = a.D, DA = a.DA, NP = np_of(D);
    const uint32_t ksteps = NP / 16u;
    const uint32_t b_chunk_bytes = NP * 128u;                 // NP rows of 128 B
    const uint32_t a_hi = smem;                               // [2 chunks][16 KB]
    const uint32_t a_lo = a_hi + 2 * A_CHUNK_BYTES;
    const uint32_t b_hi = a_lo + 2 * A_CHUNK_BYTES;           // [2 chunks][NP x 128 B]
    const uint32_t b_lo = b_hi + 2 * A_CHUNK_BYTES;
    const uint32_t part = b_lo + 2 * A_CHUNK_BYTES;           // [4][128] f32: partial sums of squares per column group
    const uint32_t uscale = part + 4 * TILE * 4;              // [128] f32: 1 / scale of U's row i
    const uint32_t mbar = uscale + TILE * 4;                  // [2] u64: MMA(stage) complete
    const uint32_t tmem_slot = mbar + 16;
    const uint32_t ybar = mbar + 32;                          // [16] u64: the own rows of warp w's 8 walkers have landed
    const uint32_t ybuf = part + 8192;                        // [128 rows][DA] f32, row r at ybuf + r * 4 * DA

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t nwalk = a.k_end - a.k_begin;
    const uint32_t n_tiles = (nwalk + TILE - 1) / TILE;
    const uint32_t my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t nvec = DA / 4u;                            // float4 pieces of a mirror row, <= 32
    const bool lane_on = lane < nvec;

    // ---- one-off: barriers, TMEM, zeroed operand tiles, U -> scaled fp16 hi/lo planes ----
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar + 8), "r"(1) : "memory");
        for (uint32_t w = 0; w < THREADS / 32; w++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ybar + 8u * w), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = tid; i < 8u * A_CHUNK_BYTES / 16u; i += THREADS) sts128z(smem + 16u * i);   // K / N padding reads as 0
    __syncthreads();
    for (uint32_t n = warp; n < D; n += THREADS / 32) {       // one warp per row of U
        const float* row = a.U + (size_t)n * D;
        float v[4];
        float m = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t k = lane + 32u * q;
            v[q] = k < D ? __ldg(row + k) : 0.0f;
            m = fmaxf(m, fabsf(v[q]));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s, inv;
        pow2_scale(m, &s, &inv);
        if (lane == 0) stsf(uscale + 4u * n, inv);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t k = lane + 32u * q;
            if (k >= D) continue;
            const float x = v[q] * s;
            const __half h = __float2half_rn(x);
            const __half l = __float2half_rn(x - __half2float(h));
            const uint32_t off = (k >> 6) * b_chunk_bytes + swz(n, k & 63u);
            sts16(b_hi + off, __half_as_ushort(h));
            sts16(b_lo + off, __half_as_ushort(l));
        }
    }
    for (uint32_t n = D + tid; n < (uint32_t)TILE; n += THREADS) stsf(uscale + 4u * n, 0.0f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    // kind::f16, D = f32, A = B = f16, both K-major, N = NP, M = 128
    const uint32_t idesc = (1u << 4) | ((NP >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

    // this lane's slice of mu; entries past D are 0, like the mirror's row padding, so padded columns centre to 0
    float4 mu4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_on) {
        const uint32_t k = 4u * lane;
        mu4.x = k < D ? __ldg(a.mu + k) : 0.f;
        mu4.y = k + 1 < D ? __ldg(a.mu + k + 1) : 0.f;
        mu4.z = k + 2 < D ? __ldg(a.mu + k + 2) : 0.f;
        mu4.w = k + 3 < D ? __ldg(a.mu + k + 3) : 0.f;
    }
    // offsets of this lane's 8-byte slot inside an operand row, and of this warp's rows: warp w owns rows 8w .. 8w+7
    const uint32_t kq = 4u * lane;
    const uint32_t row0 = warp * ROWS_PER_WARP;               // a multiple of 8: the rows of ONE swizzle atom
    const uint32_t a_off = (kq >> 6) * A_CHUNK_BYTES + (row0 >> 3) * 1024u + (kq & 7u) * 2u;
    const uint32_t kgrp = (kq & 63u) >> 3;                    // 16-byte group of the slot, XORed with the row below

    float4 Y[ROWS_PER_WARP];                                  // proposals of this warp's rows (kept until write-back)
    float4 Xj[ROWS_PER_WARP];                                 // partner rows of the NEXT tile, in flight
    // per-walker scalars live in lanes 0..7 (lane i <-> row 8w + i): current tile and next tile
    float z_cur = 1.f, u_cur = 2.f, lp_cur = 0.f, inv_cur = 1.f;
    float z_nxt = 1.f, u_nxt = 2.f, lp_nxt = 0.f, inv_nxt = 1.f;
    const float* pj_nxt = nullptr;                            // partner row of this lane's walker
    uint32_t valid_cur = 0, valid_nxt = 0;                    // walker exists (k < k_end)

    auto tile_k0 = [&](uint32_t it) { return a.k_begin + (blockIdx.x + it * gridDim.x) * (uint32_t)TILE; };
    const float dm1 = (float)(D - 1);

    // Philox draws for this warp's rows of tile `it` (lanes 0..7), then the partner rows start travelling
    auto draw_and_load = [&](uint32_t it) {
        const uint32_t k0 = tile_k0(it) + row0;
        {
            const uint32_t k = k0 + (lane & 7u);
            valid_nxt = k < a.k_end ? 1u : 0u;
            const uint32_t kc = valid_nxt ? k : a.k_end - 1;      // rows past the end repeat the last walker (never stored)
            uint32_t r[4];
            philox4x32_10(kc, a.step, a.tag, 0xbeeff00du, a.seed, 0xdecafbadu, r);
            const float ux = u01(r[0]), uy = u01(r[1]);
            u_nxt = valid_nxt ? u01(r[2]) : 2.0f;
            z_nxt = __fadd_rn(__fmaf_rn(__fmul_rn(a.cA, uy), uy, __fmul_rn(a.cB, uy)), a.cC);
            uint32_t j = (uint32_t)__fmul_rn(ux, (float)a.K);
            j = j < a.K ? j : a.K - 1;
            lp_nxt = __ldcg(a.lp_active + kc);
            pj_nxt = reinterpret_cast<const float*>(a.xa_compl[j / a.hs]) + (size_t)j * DA;
        }
        // the own rows of this warp — ONE contiguous block — travel into shared memory by a bulk copy (no registers in
        // flight, read back in the convert phase); rows past the end of the slice are left out
        if (BULK && lane == 0 && k0 < a.k_end) {
            const uint32_t bytes = min((uint32_t)ROWS_PER_WARP, a.k_end - k0) * 4u * DA;
            const uint32_t bar = ybar + 8u * warp;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(ybuf + row0 * 4u * DA), "l"(a.xa_active + (size_t)k0 * DA), "r"(bytes), "r"(bar) : "memory");
        }
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; i++) {
            const unsigned long long p = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)pj_nxt, i);
            Xj[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane_on) {
                if (!BULK) {   // own rows: only pulled into L2 here
                    const uint32_t k = min(k0 + (uint32_t)i, a.k_end - 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float4*>(a.xa_active + (size_t)k * DA) + lane));
                }
                // the random (possibly remote) partner row travels into registers now
                Xj[i] = __ldcg(reinterpret_cast<const float4*>(p) + lane);   // maybe another GPU's memory: never via L1
            }
        }
    };

    // proposal, centring, row scale, fp16 split -> operand tiles; keeps Y for the write-back
    auto convert_phase = [&](uint32_t it) {
        const uint32_t k0 = tile_k0(it) + row0;
        if (BULK && k0 < a.k_end) {                         // this warp's it-th bulk copy (warps past the end issue none)
            const uint32_t bar = ybar + 8u * warp, parity = it & 1u;
            uint32_t done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!done);
        }
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; i++) {          // own rows out of shared memory
            Y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane_on) {
                if (BULK) Y[i] = ldsf4(ybuf + (row0 + (uint32_t)i) * 4u * DA + 16u * lane);
                else Y[i] = __ldcg(reinterpret_cast<const float4*>(a.xa_active + (size_t)min(k0 + (uint32_t)i, a.k_end - 1) * DA) + lane);
            }
        }
        float inv_mine = 1.0f;
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; i++) {
            const float z = __shfl_sync(0xffffffffu, z_nxt, i);
            float4 y, c;
            y.x = __fmaf_rn(z, __fsub_rn(Y[i].x, Xj[i].x), Xj[i].x);
            y.y = __fmaf_rn(z, __fsub_rn(Y[i].y, Xj[i].y), Xj[i].y);
            y.z = __fmaf_rn(z, __fsub_rn(Y[i].z, Xj[i].z), Xj[i].z);
            y.w = __fmaf_rn(z, __fsub_rn(Y[i].w, Xj[i].w), Xj[i].w);
            Y[i] = y;
            c.x = y.x - mu4.x; c.y = y.y - mu4.y; c.z = y.z - mu4.z; c.w = y.w - mu4.w;
            float m = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w)));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s, inv;
            pow2_scale(m, &s, &inv);
            if (lane == (uint32_t)i) inv_mine = inv;
            if (lane_on) {
                const float x0 = c.x * s, x1 = c.y * s, x2 = c.z * s, x3 = c.w * s;
                const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
                const uint32_t off = a_off + (uint32_t)i * 128u + ((kgrp ^ (uint32_t)i) << 4);   // row 8w + i of the atom
                sts64(a_hi + off, *reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
                sts64(a_lo + off, *reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            }
        }
        inv_nxt = inv_mine;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
    };

    // 3 x ksteps MMAs of tile `it` into TMEM stage it & 1 (called by ONE thread, after a CTA barrier)
    auto mma_phase = [&](uint32_t it) {
        tc::tc_fence_after();
        const uint32_t d = tmem_base + (it & 1u) * (uint32_t)TILE;
        uint32_t acc = 0;
        for (uint32_t ks = 0; ks < ksteps; ks++) {
            const uint32_t kc = ks >> 2, kk = ks & 3u;
            const uint64_t ah = tc::make_desc(a_hi + kc * A_CHUNK_BYTES) + 2 * kk;
            const uint64_t al = tc::make_desc(a_lo + kc * A_CHUNK_BYTES) + 2 * kk;
            const uint64_t bh = tc::make_desc(b_hi + kc * b_chunk_bytes) + 2 * kk;
            const uint64_t bl = tc::make_desc(b_lo + kc * b_chunk_bytes) + 2 * kk;
            tc_mma_f16(d, ah, bh, idesc, acc);   // hi . hi
            acc = 1;
            tc_mma_f16(d, ah, bl, idesc, 1);     // hi . lo
            tc_mma_f16(d, al, bh, idesc, 1);     // lo . hi
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar + 8u * (it & 1u)) : "memory");
    };

    // epilogue of tile `it`: sum of squares out of TMEM, accept test, write-back of accepted rows
    auto epilogue_phase = [&](uint32_t it) {
        const uint32_t k0 = tile_k0(it) + row0;
        {
            const uint32_t bar = mbar + 8u * (it & 1u), parity = (it >> 1) & 1u;
            uint32_t done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!done);
        }
        tc::tc_fence_after();
        const uint32_t quarter = warp & 3u, grp = warp >> 2;     // TMEM lanes 32*quarter.., columns 32*grp..
        const uint32_t c0 = grp * 32u;
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + (it & 1u) * (uint32_t)TILE + c0;
        float s0 = 0.f, s1 = 0.f;
        uint32_t ra[16];
#pragma unroll 1
        for (uint32_t cc = c0; cc < c0 + 32u && cc < NP; cc += 16u) {   // warp-uniform trip count (0, 1 or 2)
            BAY_TMEM_LD16(ra, taddr + (cc - c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
                const float4 us = ldsf4(uscale + 4u * (cc + q));
                const float v0 = __uint_as_float(ra[q]) * us.x, v1 = __uint_as_float(ra[q + 1]) * us.y;
                const float v2 = __uint_as_float(ra[q + 2]) * us.z, v3 = __uint_as_float(ra[q + 3]) * us.w;
                s0 = fmaf(v0, v0, s0); s1 = fmaf(v1, v1, s1);
                s0 = fmaf(v2, v2, s0); s1 = fmaf(v3, v3, s1);
            }
        }
        stsf(part + 4u * (grp * TILE + quarter * 32u + lane), s0 + s1);
        tc::tc_fence_before();
        __syncthreads();
        // lanes 0..7 decide for the rows of this warp
        uint32_t acc = 0;
        if (lane < (uint32_t)ROWS_PER_WARP) {
            const uint32_t r = row0 + lane;
            const float ss = ((ldsf(part + 4u * r) + ldsf(part + 4u * (TILE + r))) +
                              (ldsf(part + 4u * (2 * TILE + r)) + ldsf(part + 4u * (3 * TILE + r)))) * inv_cur * inv_cur;
            const float ly = -0.5f * ss;
            // the accept ratio with the approximate units the reference's -use_fast_math build uses (G/:630-633)
            const float q = isfinite(ly) ? __powf(z_cur, dm1) * __expf(a.beta * (ly - lp_cur)) : 0.0f;
            acc = (valid_cur && u_cur <= q) ? 1u : 0u;
            if (acc) a.lp_active[k0 + lane] = ly;
        }
        const uint32_t accmask = __ballot_sync(0xffffffffu, acc != 0u);
        if (a.accepted && lane == 0 && accmask) atomicAdd(a.accepted, (unsigned long long)__popc(accmask));
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; i++) {
            if (lane_on && ((accmask >> i) & 1u))
                *(reinterpret_cast<float4*>(a.xa_active + (size_t)(k0 + i) * DA) + lane) = Y[i];
        }
    };

    auto advance = [&]() { z_cur = z_nxt; u_cur = u_nxt; lp_cur = lp_nxt; inv_cur = inv_nxt; valid_cur = valid_nxt; };

    if (my_tiles) {
        draw_and_load(0);
        convert_phase(0);
        advance();
        __syncthreads();
        if (tid == 0) mma_phase(0);
        for (uint32_t it = 0; it < my_tiles; it++) {
            const bool more = it + 1 < my_tiles;
            if (more) draw_and_load(it + 1);
            epilogue_phase(it);
            if (more) {
                convert_phase(it + 1);
                advance();
                __syncthreads();
                if (tid == 0) mma_phase(it + 1);
            }
        }
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

}  // namespace qf
}  // namespace bay
