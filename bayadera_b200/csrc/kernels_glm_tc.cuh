// bayadera_b200 — tensor-core (tcgen05 / TMEM / TMA, sm_100a) likelihood kernel of the GLM path.
//
// Computes, for up to 512 walkers per launch and ALL local dataset rows,
//     sp[w] = sum_rows softplus( x_row . theta_w )
// as the dense contraction  S(128 walkers x 128 rows) = Theta_blk(128 x 64) . X_tile(128 x 64)^T  on the
// 5th-generation tensor cores, followed by a thread-local softplus reduction:
//   * TMEM lanes = walkers, TMEM columns = dataset rows, so every epilogue thread owns one walker and sums
//     over the columns it reads with tcgen05.ld — no cross-lane reduction at all.
//   * fp32-level accuracy from bf16 inputs, fp32 accumulation in TMEM.  What the accept test needs is set by the
//     DATASET SIZE: at 10^7 rows the log-partition sum has a gradient of ~10^6 per coefficient, so an error d in how a
//     walker's coefficients are REPRESENTED shifts its log-density by ~10^6 * |d| — a two-piece bf16 split of theta
//     (residual 2^-17 relative) is off by 5..20 in a difference that must be good to ~1e-2 (measured against the fp64
//     traversal, tests/test_gpu_glm.py).  Theta is therefore split into THREE bf16 pieces (8+8+8 bits: exact), the
//     dataset into two (Xh, Xl: a fixed 2^-17 perturbation of the data, the same for every walker), and the product is
//     (th+tm+tl).Xh + (th+tm).Xl — five MMAs per K step (TERMS = 4 drops tm.Xl, ~2^-18 relative and zero-mean).
//     The dataset is split ONCE into its planes (same HBM bytes as the fp32 matrix), the walker block every half-step.
//   * TERMS = 3 — fp16 pieces (the default).  With the reference point (below) the contracted quantity delta = theta -
//     theta0 is small and bounded, so fp16's 11-bit pieces can be used safely: delta (scaled per column by the
//     dataset's power-of-two column scale and per call by a power of two that puts its largest entry at 2^13..2^14) in
//     TWO pieces carries 22 bits, the dataset columns (scaled into [-1, 1]) likewise, and hi.hi + hi.lo + lo.hi —
//     three MMAs — leaves 2^-22 relative on x.delta.  One MMA fewer than the bf16 scheme puts the kernel back on the
//     XU (MUFU) pipe instead of the tensor pipe.  The accumulator holds (x.delta) * scale; the epilogue's FFMA
//     a = acc / scale + eta0 undoes it.  TERMS = 4 / 5: the bf16 scheme above (BAY_GLM_TERMS=4).
//   * REFERENCE POINT.  The tensor core does not round its fp32 accumulation to nearest: addends are aligned to the
//     largest exponent and cut at ~2^-21 of it, so every row carries an error proportional to |eta| (measured: the
//     error against fp64 is unchanged with bf16-exact inputs and falls 10x when |eta| is small).  The kernel therefore
//     contracts only the SMALL part: theta = theta0 + delta, eta = x.theta0 + x.delta, with theta0 a per-sampler
//     reference point (the mean of the points of a recent call) whose eta0 = log2(e) * x.theta0 is computed once in
//     fp64 per refresh and streamed as one fp32 per row (+1.6 % HBM traffic); the planes hold the three pieces of
//     delta and the epilogue adds eta0 to the accumulator in true fp32.  Near the posterior |x.delta| ~ 1e-2 and the
//     tensor-core error drops below that of the fp32 SIMT traversal; far from it (early burn-in) it degrades
//     gracefully to the plain scheme, where differences of log-densities are huge anyway.
//   * Persistent, warp-specialised CTA (1 per SM): warp 16 = TMA producer (3-stage ring of 32 KB X tiles),
//     warp 17 = TMEM allocator + single-thread MMA issuer (4 accumulator stages of 128 columns),
//     warps 0-15 = four epilogue groups (4 warps = 128 TMEM lanes each), group g drains accumulator stage g.
//   * A CTA holds the three planes of 256 walkers (96 KB); a launch covers up to FOUR such walker groups: CTA c works
//     for group c % G on the row tiles c / G, c / G + gridDim / G, ...  The G CTAs of a slice stream the same row
//     tiles at about the same time, so the dataset comes from HBM once per launch and from L2 for the other G - 1.
//   * softplus(e) = max(e,0) + log(1 + exp(-|e|)); the log is taken of a running PRODUCT of 64 factors (1+t),
//     so the epilogue costs one MUFU.EX2 per element and one MUFU.LG2 per 64 (MUFU is the binding unit).
// There is no counterpart in the reference (its LOGFN loops over the dataset serially in every thread,
// e.g. K/cuda/distributions/gaussian.cu:40-42); the arithmetic contract is the oracle's serial model.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bay {
namespace tc {

constexpr int TILE = 128;                       // MMA M (walkers per block) and N (rows per tile)
constexpr int KD = 64;                          // K extent of one operand tile (128 B of bf16 = one swizzle row)
constexpr int MAX_KC = 2;                       // K chunks per dataset row: model dimension <= 128
constexpr int NSTAGE = 3;                       // X-tile ring
constexpr int NACC = 4;                         // TMEM accumulator stages (4 x 128 columns = 512)
constexpr int MAX_WB = 2;                         // walker blocks per CTA (256 walkers: three bf16 planes = 96 KB)
constexpr int MAX_GROUPS = 4;                     // walker groups per launch (CTA c serves group c % groups)
constexpr uint32_t TILE_BYTES = TILE * KD * 2;  // one bf16 plane of a tile: 16 KB
constexpr int NGRP = 4;                         // epilogue groups (4 warps each), one per accumulator stage
constexpr int ETA_RING = 8;                     // eta0 tiles in flight: producer <= NSTAGE tiles ahead of the MMA, MMA <=
                                                // NACC tiles ahead of the slowest epilogue group: 3 + 4 < 8
constexpr int THREADS = (4 * NGRP + 2) * 32;    // 16 epilogue warps + TMA warp + MMA warp
constexpr uint32_t IDESC =                      // kind::f16: D=f32, A=B=bf16, both K-major, N=128, M=128
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
constexpr uint32_t IDESC_F16 =                  // the same with A=B=f16
    (1u << 4) | ((uint32_t)(TILE >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);

__host__ __device__ constexpr size_t smem_bytes(int nwb, int nkc, int pieces = 3) {
    return 1024 /*alignment slack*/ + (size_t)nwb * nkc * pieces * TILE_BYTES + (size_t)NSTAGE * 2 * TILE_BYTES + 256 +
           (size_t)ETA_RING * TILE * 4;
}

// ------------------------------------------------------------------ PTX helpers --
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <uint32_t DESC>
__device__ __forceinline__ void tc_mma_f16kind(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(DESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    tc_mma_f16kind<IDESC>(tmem_d, adesc, bdesc, accumulate);
}
// K-major, SWIZZLE_128B operand tile (rows of 128 B, 8-row atoms of 1024 B): SBO = 1024 B, version 1, layout 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    const uint32_t lo = (smem_addr >> 4) & 0x3FFFu;
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lanes' worth of work) — the epilogue is bound by issue
// slots and the XU pipe, not by the FMA pipe.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// 2^-|a| for a PAIR on the FMA pipe instead of the XU pipe: x = max(-|a|, -126) = n + f with n = round(x), f in
// [-1/2, 1/2]; 2^f by a degree-5 minimax polynomial (relative error 7.5e-8, below MUFU.EX2's 2^-22), n added to the
// exponent field.  1.5 * 2^23 as the rounding constant: the low mantissa bits of x + magic hold n in two's complement.
// BAY_GLM_POLY8 = elements out of every 8 that take the polynomial (0, 2 or 4); the rest take MUFU.EX2.
// MEASURED (profiles/r02_c4_epilogue_variants.txt): with 2 of 8 the kernel needs 4.3 % fewer cycles (XU pipe 86 % ->
// 69 %) and the SM clock settles 4.4 % lower: the kernel runs at the board's power cap (sw_power_cap, ~1.5 of 1.97
// GHz), where cycles saved on one pipe by spending more instructions on another buy nothing.  Default: 0.
#ifndef BAY_GLM_POLY8
#define BAY_GLM_POLY8 0
#endif
__device__ __forceinline__ uint64_t ex2_neg_abs_poly2(float a0, float a1) {
    const float x0 = fmaxf(-fabsf(a0), -126.f), x1 = fmaxf(-fabsf(a1), -126.f);
    const uint64_t magic = pk2(12582912.f, 12582912.f);
    const uint64_t x = pk2(x0, x1);
    const uint64_t t = add2(x, magic);
    const uint64_t f = sub2(x, sub2(t, magic));
    uint64_t p = pk2(0x1.5c08b6p-10f, 0x1.5c08b6p-10f);
    p = fma2(p, f, pk2(0x1.3d0c4ap-7f, 0x1.3d0c4ap-7f));
    p = fma2(p, f, pk2(0x1.c6b6e6p-5f, 0x1.c6b6e6p-5f));
    p = fma2(p, f, pk2(0x1.ebf918p-3f, 0x1.ebf918p-3f));
    p = fma2(p, f, pk2(0x1.62e428p-1f, 0x1.62e428p-1f));
    p = fma2(p, f, pk2(0x1.000002p+0f, 0x1.000002p+0f));
    uint32_t p0, p1, n0, n1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(p));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(n0), "=r"(n1) : "l"(t));
    return pk2u(p0 + (n0 << 23), p1 + (n1 << 23));
}

#define BAY_TMEM_LD32(r, taddr)                                                                              \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"            \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),      \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),          \
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),          \
          "=r"(r[30]), "=r"(r[31])                                                                           \
        : "r"(taddr) : "memory")

#define BAY_TMEM_LD16(r, taddr)                                                                              \
    asm volatile(                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                     \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),      \
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])  \
        : "r"(taddr) : "memory")

// ------------------------------------------------------------------ the kernel --
// map_xh/map_xl: [rows][64*NKC] bf16 planes of the dataset; map_ah/map_am/map_al: [n_walkers][64*NKC] bf16 planes of
// the launch's walkers (three pieces of theta * log2 e), origin at the first walker of this launch.
// groups: walker groups of NWB*128 walkers in this launch; CTA c serves group c % groups on row-tile slice c / groups.
// partial: [NGRP * slices][ldp] doubles; entry (NGRP*slice + epilogue group, (g*NWB + wb)*128 + lane) = that
// epilogue thread's sum.
// NKC = 2 (64 < DIM <= 128): a row tile arrives as two 64-wide K chunks, each its own ring stage, and the MMAs of
// both accumulate into the same TMEM stage; the walker planes then take 96 KB per 128 walkers, so NWB = 1.
// LINK = 0: A(eta) = softplus(eta) (Bernoulli-logit); LINK = 1: A(eta) = exp(eta) (Poisson-log): one MUFU.EX2 and one
// FADD per element, nothing analytic left for the finish kernel.
template <int NWB, int NKC, int LINK, int TERMS>
__global__ void __launch_bounds__(THREADS, 1)  // (NWB, NKC) in {1, 2} x {1}, {1} x {2}
k_glm_loglik_tc(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_am,
                const __grid_constant__ CUtensorMap map_al, const float* __restrict__ eta0,
                const float* __restrict__ dscale, const uint32_t rows, const uint32_t n_tiles, const uint32_t groups,
                double* __restrict__ partial, const uint32_t ldp) {
    constexpr int PIECES = TERMS == 3 ? 2 : 3;              // planes of the walker block: fp16 (hi, lo) or bf16 (hi, mid, lo)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_hi = smem;                                   // NWB x NKC tiles, tile (wb, kc) at wb*NKC + kc
    uint8_t* a_mi = a_hi + NWB * NKC * TILE_BYTES;          // (absent with two pieces)
    uint8_t* a_lo = a_hi + (PIECES - 1) * NWB * NKC * TILE_BYTES;
    uint8_t* b_base = a_lo + NWB * NKC * TILE_BYTES;        // NSTAGE x (hi, lo)
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + NSTAGE * 2 * TILE_BYTES);
    uint64_t* full_bar = bars;                              // [NSTAGE] TMA -> MMA
    uint64_t* empty_bar = bars + NSTAGE;                    // [NSTAGE] MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * NSTAGE;                // [NACC]   MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * NSTAGE + NACC;        // [NACC]   epilogue -> MMA
    uint64_t* a_bar = bars + 2 * NSTAGE + 2 * NACC;         // walker block landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * NACC + 1);
    const uint32_t eta_s = smem_u32(b_base + NSTAGE * 2 * TILE_BYTES + 256);   // [ETA_RING][TILE] fp32, 16-byte aligned

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x % groups, slice = blockIdx.x / groups, slices = gridDim.x / groups;
    const uint32_t my_tiles = (n_tiles > slice) ? (n_tiles - slice + slices - 1) / slices : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < NACC; i++) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        mbar_init(a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4 * NGRP + 1) {   // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * NGRP) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(a_bar, NWB * NKC * PIECES * TILE_BYTES);
            for (int wb = 0; wb < NWB; wb++)
                for (int kc = 0; kc < NKC; kc++) {
                    const int wrow = (int)((g * NWB + wb) * TILE);
                    tma_load_2d(a_hi + (wb * NKC + kc) * TILE_BYTES, &map_ah, kc * KD, wrow, a_bar);
                    if (PIECES == 3) tma_load_2d(a_mi + (wb * NKC + kc) * TILE_BYTES, &map_am, kc * KD, wrow, a_bar);
                    tma_load_2d(a_lo + (wb * NKC + kc) * TILE_BYTES, &map_al, kc * KD, wrow, a_bar);
                }
            for (uint32_t it = 0; it < my_tiles; it++) {
                const int row0 = (int)((slice + it * slices) * TILE);
                for (uint32_t kc = 0; kc < NKC; kc++) {
                    const uint32_t st = it * NKC + kc, s = st % NSTAGE, ph = (st / NSTAGE) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], 2 * TILE_BYTES + (kc == 0 ? TILE * 4u : 0u));
                    if (kc == 0)   // the tile's eta0 values (eta0 is padded to a whole tile): one 512-byte bulk copy
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(eta_s + (it % ETA_RING) * TILE * 4u), "l"(eta0 + (size_t)row0), "r"(TILE * 4u),
                                       "r"(smem_u32(&full_bar[s])) : "memory");
                    tma_load_2d(b_base + (2 * s) * TILE_BYTES, &map_xh, (int)(kc * KD), row0, &full_bar[s]);
                    tma_load_2d(b_base + (2 * s + 1) * TILE_BYTES, &map_xl, (int)(kc * KD), row0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 4 * NGRP + 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            mbar_wait(a_bar, 0);
            uint32_t item = 0;
            for (uint32_t it = 0; it < my_tiles; it++) {
                uint32_t stage[NKC];
                uint64_t bh[NKC], bl[NKC];
#pragma unroll
                for (int kc = 0; kc < NKC; kc++) {          // all K chunks of the row tile must have landed
                    const uint32_t st = it * NKC + kc, s = st % NSTAGE, ph = (st / NSTAGE) & 1u;
                    mbar_wait(&full_bar[s], ph);
                    stage[kc] = s;
                    bh[kc] = make_desc(smem_u32(b_base + (2 * s) * TILE_BYTES));
                    bl[kc] = make_desc(smem_u32(b_base + (2 * s + 1) * TILE_BYTES));
                }
                tc_fence_after();
#pragma unroll
                for (int wb = 0; wb < NWB; wb++, item++) {
                    const uint32_t a = item % NACC, aph = (item / NACC) & 1u;
                    mbar_wait(&tempty_bar[a], aph ^ 1u);
                    tc_fence_after();
                    const uint32_t d = tmem_base + a * TILE;
#pragma unroll
                    for (int kc = 0; kc < NKC; kc++) {
                        const uint64_t ah = make_desc(smem_u32(a_hi + (wb * NKC + kc) * TILE_BYTES));
                        const uint64_t am = make_desc(smem_u32(a_mi + (wb * NKC + kc) * TILE_BYTES));
                        const uint64_t al = make_desc(smem_u32(a_lo + (wb * NKC + kc) * TILE_BYTES));
                        // smallest products first: they enter the fp32 accumulator before the leading term does
                        if (TERMS == 3) {   // fp16 pieces: lo.hi + hi.lo + hi.hi
#pragma unroll
                            for (int k = 0; k < KD / 16; k++) tc_mma_f16kind<IDESC_F16>(d, al + 2 * k, bh[kc] + 2 * k, kc > 0 || k > 0);
#pragma unroll
                            for (int k = 0; k < KD / 16; k++) tc_mma_f16kind<IDESC_F16>(d, ah + 2 * k, bl[kc] + 2 * k, 1);
#pragma unroll
                            for (int k = 0; k < KD / 16; k++) tc_mma_f16kind<IDESC_F16>(d, ah + 2 * k, bh[kc] + 2 * k, 1);
                            continue;
                        }
#pragma unroll
                        for (int k = 0; k < KD / 16; k++) tc_mma_bf16(d, al + 2 * k, bh[kc] + 2 * k, kc > 0 || k > 0);   // lo*hi
                        if (TERMS >= 5) {
#pragma unroll
                            for (int k = 0; k < KD / 16; k++) tc_mma_bf16(d, am + 2 * k, bl[kc] + 2 * k, 1);             // mid*lo
                        }
#pragma unroll
                        for (int k = 0; k < KD / 16; k++) tc_mma_bf16(d, am + 2 * k, bh[kc] + 2 * k, 1);                 // mid*hi
#pragma unroll
                        for (int k = 0; k < KD / 16; k++) tc_mma_bf16(d, ah + 2 * k, bl[kc] + 2 * k, 1);                 // hi*lo
#pragma unroll
                        for (int k = 0; k < KD / 16; k++) tc_mma_bf16(d, ah + 2 * k, bh[kc] + 2 * k, 1);                 // hi*hi
                    }
                    tc_commit(&tfull_bar[a]);
                }
#pragma unroll
                for (int kc = 0; kc < NKC; kc++) tc_commit(&empty_bar[stage[kc]]);   // chunks free once their MMAs completed
            }
        }
    } else {
        // ===================== epilogue groups =====================
        // Group e drains accumulator stage e, i.e. items e, e+4, e+8, ... (item = tile*NWB + wb), which for
        // NWB in {1,2} all belong to ONE walker block wb = e % NWB: a thread owns exactly one walker.
        // The loops are kept rolled on purpose: fully unrolled, the epilogue was ~100 KB of SASS and the
        // kernel stalled on instruction fetch (ncu: stall_no_inst).
        const uint32_t grp = warp >> 2;
        const uint32_t quarter = warp & 3;               // TMEM lanes 32*quarter .. +31
        const uint32_t wb_mine = grp % NWB;
        const uint32_t tbase = tmem_base + ((quarter * 32u) << 16) + grp * TILE;
        const uint32_t n_items = my_tiles * NWB;
        // The dataset planes are pre-scaled by log2(e), so the accumulators hold a = eta*log2(e) and
        //   softplus(eta) = max(eta,0) + ln2*log2(1 + 2^-|a|),   sum_rows max(eta,0) = (sum eta + sum |eta|)/2,
        // where sum_rows eta = (sum_rows x_row) . theta is added analytically by k_glm_finish_tc.  Per element that
        // leaves MUFU.EX2(-|a|), one FFMA on the running product and one FADD on sum|a|.
        // Rows that TMA zero-filled past the end of the dataset give a = 0 exactly: no |a|, factor 2 -> corrected
        // by subtracting (TILE - valid) from the log2 sum.
        const float inv_scale = TERMS == 3 ? __ldg(dscale + 1) : 1.0f;   // fp16 pieces: accumulators hold (x.delta) * scale
        const uint64_t inv_scale2 = pk2(inv_scale, inv_scale);
        float hi = 0.f, lo = 0.f;                        // two-float (compensated) sum over this thread's items
        for (uint32_t item = grp; item < n_items; item += NACC) {
            const uint32_t it = item / NWB;
            const uint32_t row0 = (slice + it * slices) * TILE;
            const uint32_t valid = min((uint32_t)TILE, rows - row0);
            mbar_wait(&tfull_bar[grp], (item / NACC) & 1u);
            tc_fence_after();
            // four independent (product, |a|-sum) chains; the next 16 columns are in flight while 16 are reduced
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
            uint64_t pa = pk2(1.f, 1.f), pb = pa;        // four product chains in two packed pairs
            uint32_t ra[16], rb[16];
            const uint32_t eta_tile = eta_s + (it % ETA_RING) * TILE * 4u;
            // a = eta0[row] + (x . delta): the reference part is added here, in round-to-nearest fp32
            auto reduce = [&](const uint32_t (&r)[16], const uint32_t c) {
                if (LINK == 1) {   // sum of 2^a = exp(eta)
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 eta;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(eta.x), "=f"(eta.y), "=f"(eta.z), "=f"(eta.w) : "r"(eta_tile + (c + j) * 4u));
                        const float a0 = fmaf(__uint_as_float(r[j]), inv_scale, eta.x), a1 = fmaf(__uint_as_float(r[j + 1]), inv_scale, eta.y);
                        const float a2 = fmaf(__uint_as_float(r[j + 2]), inv_scale, eta.z), a3 = fmaf(__uint_as_float(r[j + 3]), inv_scale, eta.w);
                        m0 += ex2_approx(a0); m1 += ex2_approx(a1);
                        m2 += ex2_approx(a2); m3 += ex2_approx(a3);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        uint64_t eta[4], t[4];
                        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(eta[0]), "=l"(eta[1]) : "r"(eta_tile + (c + j) * 4u));
                        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(eta[2]), "=l"(eta[3]) : "r"(eta_tile + (c + j + 4) * 4u));
                        float a[8];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            upk2(fma2(pk2u(r[j + 2 * q], r[j + 2 * q + 1]), inv_scale2, eta[q]), a[2 * q], a[2 * q + 1]);
                        m0 += fabsf(a[0]); m1 += fabsf(a[1]); m2 += fabsf(a[2]); m3 += fabsf(a[3]);
                        m0 += fabsf(a[4]); m1 += fabsf(a[5]); m2 += fabsf(a[6]); m3 += fabsf(a[7]);
#pragma unroll
                        for (int q = 0; q < 4; q++) {   // the LAST BAY_GLM_POLY8 / 2 pairs go to the FMA pipe
                            if (q >= 4 - BAY_GLM_POLY8 / 2) t[q] = ex2_neg_abs_poly2(a[2 * q], a[2 * q + 1]);
                            else t[q] = pk2(ex2_approx(-fabsf(a[2 * q])), ex2_approx(-fabsf(a[2 * q + 1])));
                        }
                        pa = fma2(pa, t[0], pa); pb = fma2(pb, t[1], pb);
                        pa = fma2(pa, t[2], pa); pb = fma2(pb, t[3], pb);
                    }
                }
            };
            BAY_TMEM_LD16(ra, tbase);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll 1
            for (uint32_t c = 0; c < TILE; c += 32) {
                BAY_TMEM_LD16(rb, tbase + c + 16);
                reduce(ra, c);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 32 < TILE) {
                    BAY_TMEM_LD16(ra, tbase + c + 32);
                } else {
                    // every column of the stage is in registers: hand the accumulator back before the last reduction
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[grp]);
                }
                reduce(rb, c + 16);
                if (c + 32 < TILE) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            // each chain holds 32 factors <= 2.  item value (in units of ln2): sum log2(1+t) + sum|a|/2
            // zero-filled rows past the end of the dataset contribute exactly 1 each (2^0, or one factor 2)
            float x;
            if (LINK == 1) {
                x = ((m0 + m1) + (m2 + m3)) - (float)(TILE - valid);
            } else {
                float p0, p1, p2, p3;
                upk2(pa, p0, p1);
                upk2(pb, p2, p3);
                const float lg = (lg2_approx(p0 * p1) + lg2_approx(p2 * p3)) - (float)(TILE - valid);
                x = fmaf(0.5f, (m0 + m1) + (m2 + m3), lg);
            }
            const float s = hi + x;                      // Knuth two-sum: (hi, lo) += x without fp64 (DADD is slow here)
            const float bp = s - hi;
            lo += (hi - (s - bp)) + (x - bp);
            hi = s;
        }
        const double acc64 = ((double)hi + (double)lo) * (LINK == 1 ? 1.0 : 0.6931471805599453);
        const size_t base = (size_t)(NGRP * slice + grp) * ldp + (size_t)g * NWB * TILE + quarter * 32 + lane;
#pragma unroll
        for (int wb = 0; wb < NWB; wb++) partial[base + wb * TILE] = (wb == (int)wb_mine) ? acc64 : 0.0;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4 * NGRP + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

constexpr float LOG2E = 1.4426950408889634f;

// points (SoA, dim x n, pitch) -> three bf16 planes [n][kdp] (row = walker; kdp = 64 or 128) of theta - theta0, the
// A operand: x = hi + mid + lo EXACTLY (8 + 8 + 8 significant bits; both subtractions are exact in fp32)
__global__ void k_glm_split_points(const float* __restrict__ pts, uint32_t pitch, uint32_t n, uint32_t dim,
                                   uint32_t kdp, const float* __restrict__ theta0, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;    // e = k*kdp + i, coalesced stores
    if (e >= (uint64_t)n * kdp) return;
    const uint32_t k = (uint32_t)(e / kdp), i = (uint32_t)(e % kdp);
    // dim < kdp: the K extent is zero-padded.  The coefficients enter UNSCALED: the factor log2(e) the epilogue wants
    // lives in the dataset planes (k_glm_split_rows) — rounding theta * log2(e) to fp32 would move every walker by up
    // to 2^-24 relative, which 10^7 rows amplify to ~0.1 in the log-density (measured), while the same rounding
    // applied to the dataset is one fixed perturbation shared by all walkers.
    // delta = theta - theta0: exact whenever the two are within a factor of two of each other, else rounded at
    // 2^-24 |delta| — far from the reference point, where log-density differences are large
    const float x = i < dim ? __fsub_rn(pts[(size_t)i * pitch + k], theta0[i]) : 0.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(h));
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    hi[e] = h;
    mid[e] = m;
    lo[e] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(m)));
}

// The likelihood launches of one call (up to 1024 walkers each) leave their partial sums in consecutive regions of
// `partial`; ONE finish launch reduces them all.
struct FinishPlan {
    uint32_t n_launch;
    uint32_t k0[8], cnt[8], chunks[8], ldp[8];   // first walker, walkers, partial rows, leading dimension per launch
    unsigned long long off[8];                   // offset of the launch's region in `partial` (doubles)
};

// sp[k] = sum over the kernel's partial rows + (1/2) * sx . theta_k: the analytic sum_rows eta term of
// sum max(eta,0) = (sum eta + sum |eta|) / 2.  sx = column sums of the LOCAL rows.  One warp per walker: lanes
// stride over the partial rows, then a fixed shuffle tree (deterministic: same order on every run and rank).
__global__ void k_glm_finish_tc(const __grid_constant__ FinishPlan plan, uint32_t kbase, uint32_t n,
                                const double* __restrict__ partial, const double* __restrict__ sx,
                                const float* __restrict__ pts, uint32_t pitch, uint32_t dim, uint32_t link,
                                double* __restrict__ sp) {
    // walkers kbase .. kbase + n - 1 (those the plan's launches cover)
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint32_t k = kbase + w;
    uint32_t l = 0;
    while (l + 1 < plan.n_launch && k >= plan.k0[l + 1]) l++;
    const double* part = partial + plan.off[l];
    const uint32_t kk = k - plan.k0[l], chunks = plan.chunks[l], ldp = plan.ldp[l];
    double s = 0.0;
    for (uint32_t c = lane; c < chunks; c += 32) s += part[(size_t)c * ldp + kk];
    if (link == 0)   // the analytic sum_rows eta / 2 of the softplus split; exp has no such term
        for (uint32_t i = lane; i < dim; i += 32)
            s += 0.5 * sx[i] * (double)pts[(size_t)i * pitch + k];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sp[k] = s;
}

// Reference point of the contraction, ONE launch (grid = dim CTAs).
// Step 1, every CTA: mean and rms = sqrt(mean square) of its coordinate over the n points (double accumulation).
// Step 2, the CTA that finishes last: theta0 = the mean snapped to a coarse grid, q * rint(mean / q) with q = 2^-3 of
// the largest coordinate rms (mean AND spread: while the ensemble is still wide the grid is coarse and the reference
// stays put; once it has contracted the grid follows the size of the coefficients) rounded down to a power of two.
// A PURE function of the points (so a chain restored from a checkpoint recomputes the same reference and continues
// bit for bit), yet stable: once the ensemble has settled the snapped mean changes only when a coordinate crosses a
// grid line, and only then (flags[0] = 1) is eta0 recomputed.  |theta - theta0| <= q/2 + the ensemble's spread keeps
// the contracted part >= 10x smaller than eta itself, which puts the tensor-core accumulation error at the level of
// an fp32 traversal's summation noise.   stats: dim means then dim rms;  flags: [0] changed, [1] arrival ticket (0),
// [2] reset to 0 for the fp16 path's delta maximum.
__global__ void k_glm_reference(const float* __restrict__ pts, uint32_t pitch, uint32_t n, float* __restrict__ stats,
                                float* __restrict__ theta0, int* __restrict__ flags, int force) {
    __shared__ double sm[32], sq[32];
    __shared__ float smax[32];
    __shared__ int last, sdiff;
    const uint32_t dim = gridDim.x;
    const float* row = pts + (size_t)blockIdx.x * pitch;
    double s = 0.0, s2 = 0.0;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) { const double v = (double)row[k]; s += v; s2 += v * v; }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, t2 = 0.0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; w++) { t += sm[w]; t2 += sq[w]; }
        stats[blockIdx.x] = (float)(t / (double)n);
        stats[dim + blockIdx.x] = (float)sqrt(t2 / (double)n);
        __threadfence();
        last = atomicAdd(&flags[1], 1) == (int)dim - 1 ? 1 : 0;
        sdiff = force;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float m = 0.0f;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) m = fmaxf(m, __ldcg(stats + dim + i));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
    __syncthreads();
    m = 0.0f;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; w++) m = fmaxf(m, smax[w]);
    // power of two not above m, times 2^-3; a non-finite or zero scale falls back to the origin
    const bool usable = m > 1e-30f && m < 1e30f;
    const float q = usable ? __uint_as_float((__float_as_uint(m) & 0x7f800000u) - (3u << 23)) : 1.0f;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
        const float v = usable ? __fmul_rn(q, rintf(__fdiv_rn(__ldcg(stats + i), q))) : 0.0f;
        if (v != theta0[i]) { theta0[i] = v; sdiff = 1; }
    }
    __syncthreads();
    if (threadIdx.x == 0) { flags[0] = sdiff; flags[1] = 0; flags[2] = 0; }   // [2]: this call's delta maximum (float bits)
}

// eta0[r] = log2(e) * (x_r . theta0), fp64 accumulation; a warp takes FOUR dataset rows [y, x_1..x_dim] at a time so
// that enough loads are in flight to stream the rows at HBM speed; entries past the last row (eta0 is padded to a
// whole tile) stay zero
__global__ void k_glm_eta0(const float* __restrict__ data, uint64_t rows, uint32_t dim, const float* __restrict__ theta0,
                           const int* __restrict__ changed, float* __restrict__ eta0) {
    if (*changed == 0) return;   // same reference point as the last call: eta0 is current
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t stride = dim + 1;
    for (uint64_t r0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 4; r0 < rows; r0 += warps * 4) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (uint32_t i = lane; i < dim; i += 32) {
            const double t = (double)theta0[i];
            float x[4];
#pragma unroll
            for (int q = 0; q < 4; q++) x[q] = (r0 + q < rows) ? __ldcs(data + (r0 + q) * stride + 1 + i) : 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++) s[q] += (double)x[q] * t;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        }
        if (lane < 4 && r0 + lane < rows) {
            const double v = lane == 0 ? s[0] : (lane == 1 ? s[1] : (lane == 2 ? s[2] : s[3]));
            eta0[r0 + lane] = (float)(v * 1.4426950408889634);
        }
    }
}

// ---- fp16 pieces (TERMS = 3) ----------------------------------------------------------------------------------------
// column scales: colmax[i] = max_rows |x_ri| as float bits (atomicMax on the non-negative pattern)
__global__ void k_glm_colmax(const float* __restrict__ data, uint64_t rows, uint32_t dim, uint32_t* __restrict__ colmax) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    float m[4] = {0.f, 0.f, 0.f, 0.f};                      // dims lane, lane + 32, ... (dim <= 128)
    for (uint64_t r = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < rows; r += warps) {
        const float* row = data + r * (dim + 1) + 1;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (lane + 32u * q < dim) m[q] = fmaxf(m[q], fabsf(__ldcs(row + lane + 32u * q)));
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (lane + 32u * q < dim && m[q] > 0.f) atomicMax(&colmax[lane + 32u * q], __float_as_uint(m[q]));
}

// cscale[i] = the power of two that brings column i of x * log2(e) into [-1, 1] (1 for an all-zero column); stored as
// (cscale, 1 / cscale) pairs
__global__ void k_glm_colscale(uint32_t dim, const uint32_t* __restrict__ colmax, float* __restrict__ cscale) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    const float m = __uint_as_float(colmax[i]) * LOG2E;
    uint32_t e = (__float_as_uint(m) >> 23) & 0xffu;        // m in [2^(e-127), 2^(e-126))
    if (!(m > 0.f) || e == 0u || e > 250u) { cscale[2 * i] = 1.0f; cscale[2 * i + 1] = 1.0f; return; }
    cscale[2 * i] = __uint_as_float((253u - e) << 23);      // 2^(126-e): m * cscale in [1/2, 1)
    cscale[2 * i + 1] = __uint_as_float((e + 1u) << 23);    // 2^(e-126)
}

// dataset rows -> fp16 hi/lo planes [rows][kdp] of x * log2(e) * cscale[i], zero-padded past dim
__global__ void k_glm_split_rows_f16(const float* __restrict__ data, uint64_t rows, uint32_t dim, uint32_t kdp,
                                     const float* __restrict__ cscale, __half* __restrict__ hi, __half* __restrict__ lo) {
    const uint64_t total = rows * kdp;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t r = e / kdp;
        const uint32_t i = (uint32_t)(e % kdp);
        const float x = i < dim ? __fmul_rn(__fmul_rn(data[r * (dim + 1) + 1 + i], LOG2E), cscale[2 * i]) : 0.0f;
        const __half h = __float2half_rn(x);
        hi[e] = h;
        lo[e] = __float2half_rn(x - __half2float(h));
    }
}

// largest |(theta - theta0)_i / cscale_i| over the call's points, as float bits into *dmax (reset by the caller)
__global__ void k_glm_delta_max(const float* __restrict__ pts, uint32_t pitch, uint32_t n, uint32_t dim,
                                const float* __restrict__ theta0, const float* __restrict__ cscale,
                                uint32_t* __restrict__ dmax) {
    float m = 0.f;
    const uint64_t total = (uint64_t)n * dim, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint32_t i = (uint32_t)(e / n), k = (uint32_t)(e % n);      // coalesced along the points
        m = fmaxf(m, fabsf(__fmul_rn(__fsub_rn(pts[(size_t)i * pitch + k], theta0[i]), cscale[2 * i + 1])));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(dmax, __float_as_uint(m));   // NaN / inf points: see below
}

// points -> fp16 hi/lo planes [n][kdp] of delta' = (theta - theta0) / cscale_i * scale, scale = the power of two that
// puts the largest |delta'| of the call at 2^13..2^14 (fp16 overflows at 65504); dscale = {scale, 1 / scale}.
// A non-finite point makes the maximum non-finite: scale falls back to 1 and that walker's row goes non-finite,
// which the accept test rejects — like the generic kernel.
__global__ void k_glm_split_points_f16(const float* __restrict__ pts, uint32_t pitch, uint32_t n, uint32_t dim,
                                       uint32_t kdp, const float* __restrict__ theta0, const float* __restrict__ cscale,
                                       const uint32_t* __restrict__ dmax, float* __restrict__ dscale,
                                       __half* __restrict__ hi, __half* __restrict__ lo) {
    const float m = __uint_as_float(*dmax);
    const uint32_t em = (__float_as_uint(m) >> 23) & 0xffu;
    float scale = 1.0f, inv = 1.0f;
    if (m > 0.f && em >= 20u && em <= 230u) {                // m in [2^(em-127), 2^(em-126)) -> m * scale in [2^13, 2^14)
        scale = __uint_as_float((267u - em) << 23);          // 2^(140-em)
        inv = __uint_as_float((em - 13u) << 23);             // 2^(em-140)
    }
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) { dscale[0] = scale; dscale[1] = inv; }
    if (e >= (uint64_t)n * kdp) return;
    const uint32_t k = (uint32_t)(e / kdp), i = (uint32_t)(e % kdp);
    const float x = i < dim ? __fmul_rn(__fmul_rn(__fsub_rn(pts[(size_t)i * pitch + k], theta0[i]), cscale[2 * i + 1]), scale) : 0.0f;
    const __half h = __float2half_rn(x);
    hi[e] = h;
    lo[e] = __float2half_rn(x - __half2float(h));
}

// dataset rows [y, x_1..x_dim] (stride dim + 1) -> bf16 hi/lo planes [rows][kdp] of x * log2(e), zero-padded past dim
// (the accumulators then hold eta * log2 e, what the exp2-based epilogue consumes)
__global__ void k_glm_split_rows(const float* __restrict__ data, uint64_t rows, uint32_t dim, uint32_t kdp,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const uint64_t total = rows * kdp;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t r = e / kdp;
        const uint32_t i = (uint32_t)(e % kdp);
        const float x = i < dim ? __fmul_rn(data[r * (dim + 1) + 1 + i], LOG2E) : 0.0f;
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        hi[e] = h;
        lo[e] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

}  // namespace tc
}  // namespace bay
