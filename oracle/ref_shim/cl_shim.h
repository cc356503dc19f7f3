/* OpenCL-C -> C11 shim so that the reference's OWN kernel text (the barrier-free kernels of
 * src/device/.../opencl/engines/amd-gcn-mcmc-stretch.cl and the distribution files) can be compiled by gcc where it
 * lies under /root/reference and executed work-item by work-item on the host.  TEST INFRASTRUCTURE ONLY (oracle/).
 * Nothing of the reference is copied: ref_driver.c #includes the files by path at build time (oracle/Makefile,
 * target `ref`), the outputs go to oracle/_ref/ (git-ignored). */
#ifndef BAY_REF_SHIM_CL_H
#define BAY_REF_SHIM_CL_H
#include <math.h>
#include <stdbool.h>
#include <stdint.h>

typedef uint32_t uint;
typedef uint64_t ulong;
typedef struct { float s0, s1, s2, s3; } float4;
typedef struct { uint32_t x, y, z, w; } uint4;
typedef struct { float s0, s1; } float2;
#define REAL2 float2

#define __kernel
#define __global
#define __local
#define inline static inline
#define CLK_LOCAL_MEM_FENCE 0

/* work-item identity: set by the driver loop before each call */
static _Thread_local uint bay_gid, bay_gsize, bay_lid, bay_lsize, bay_group, bay_ngroups;
static __inline__ uint get_global_id(uint d) { (void)d; return bay_gid; }
static __inline__ uint get_global_size(uint d) { (void)d; return bay_gsize; }
static __inline__ uint get_local_id(uint d) { (void)d; return bay_lid; }
static __inline__ uint get_local_size(uint d) { (void)d; return bay_lsize; }
static __inline__ uint get_group_id(uint d) { (void)d; return bay_group; }
static __inline__ uint get_num_groups(uint d) { (void)d; return bay_ngroups; }
/* kernels with barriers are compiled (they share the file) but never run on the host */
static __inline__ void work_group_barrier(int flags) { (void)flags; }
static __inline__ float work_group_reduction_sum(float* lacc, float v) { (void)lacc; return v; }   /* ClojureCL, absent */
static __inline__ float work_group_reduction_sum_2(float* lacc, float v) { (void)lacc; return v; }   /* ClojureCL, absent */

#define native_exp(x) expf(x)
#define native_log(x) logf(x)
#define native_powr(x, y) powf((x), (y))
#define native_sqrt(x) sqrtf(x)
#define lgamma(x) lgammaf(x)   /* OpenCL's lgamma(float) is single precision */
static __inline__ float pown(float x, int n) {   /* OpenCL pown: x^n, integer n */
    float r = 1.0f;
    for (int i = 0; i < (n < 0 ? -n : n); i++) r *= x;
    return n < 0 ? 1.0f / r : r;
}

/* rng/uniform-sampler.cl:11-22 uses OpenCL vector literals, so its two conversion helpers are restated here */
#define R123_0x1p_23f 1.1920928955078125E-7f
static __inline__ float4 u01fpt_oo_4x32_24(uint4 i) {
    const float4 r = {(0.5f + (i.x >> 9)) * R123_0x1p_23f, (0.5f + (i.y >> 9)) * R123_0x1p_23f,
                      (0.5f + (i.z >> 9)) * R123_0x1p_23f, (0.5f + (i.w >> 9)) * R123_0x1p_23f};
    return r;
}
#endif
