/* Host driver around the reference's own OpenCL kernel text (see ref_shim/cl_shim.h).  TEST INFRASTRUCTURE ONLY.
 * Built per (model, DIM) by `make -C oracle ref` when /root/reference is present:
 *   -DREF_MODEL_FILE="<distribution>.cl" -DLOGFN=<mcmc logpdf name> -DDIM=<d> -DWGS=<wgs>
 * The kernels are called exactly as the reference enqueues them (amd_gcn.clj: global size = H or W or DIM*W/4,
 * local size = WGS); work-items run one after another — the bare kernels have no barrier and, within a half-step,
 * no work-item reads what another writes. */
#include "ref_shim/cl_shim.h"
#define REAL float
#define ACCUMULATOR float
#include REF_MODEL_FILE
#include "uncomplicate/bayadera/internal/device/opencl/engines/amd-gcn-mcmc-stretch.cl"
#undef inline

static void set_item(uint gid, uint gsize) {
    bay_gid = gid; bay_gsize = gsize; bay_lsize = WGS; bay_lid = gid % WGS; bay_group = gid / WGS;
    bay_ngroups = (gsize + WGS - 1) / WGS;
}

int ref_dim(void) { return DIM; }

/* init_walkers: global size DIM*W/4 (amd_gcn.clj init-position!) */
void ref_init_walkers(uint seed, const float* limits /* DIM x (lo, hi) */, float* xs, uint walkers) {
    const uint n = DIM * walkers / 4;
    for (uint g = 0; g < n; g++) { set_item(g, n); init_walkers(seed, (const float2*)limits, xs); }
}

void ref_logfn(uint data_len, uint params_len, const float* params, const float* x, float* res, uint walkers) {
    for (uint g = 0; g < walkers; g++) { set_item(g, walkers); logfn(data_len, params_len, params, x, res); }
}

/* one half-step: global size K = walkers of the active half */
void ref_stretch_move_bare(uint seed, uint odd_or_even, uint data_len, uint params_len, const float* params,
                           const float* Scompl, float* X, float* logfn_X, float a, float beta, uint step_counter,
                           uint K) {
    /* within a half-step no work-item reads what another writes, so they may run on all host cores */
    #pragma omp parallel for schedule(static)
    for (uint g = 0; g < K; g++) {
        set_item(g, K);
        stretch_move_bare(seed, odd_or_even, data_len, params_len, params, Scompl, X, logfn_X, a, beta, step_counter);
    }
}
