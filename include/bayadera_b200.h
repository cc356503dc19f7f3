/*
 * bayadera_b200.h — C ABI of libbayadera_b200.so, the B200-native (sm_100a)
 * replacement for the device engines of uncomplicate/bayadera.
 *
 * Every entry point states the reference interface it replaces.  Aliases:
 *   P/  = /root/reference/src/clojure/uncomplicate/bayadera/internal/protocols.clj
 *   G/  = /root/reference/src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj
 *   K/  = /root/reference/src/device/uncomplicate/bayadera/internal/device/cuda/
 *
 * Conventions
 *  - Plain C: opaque handles, pointers and sizes only.  No torch / C++ types.
 *  - Every function returns BAY_OK (0) or a negative bay_status; the message
 *    of the last failure on the calling thread is bay_last_error().
 *  - Matrices crossing the boundary use the reference's layout: column-major
 *    DIM x n, column = one walker / one sample (ld = DIM), fp32.  `limits`
 *    is 2 x DIM column-major (lo_d, hi_d).  Internally walker state is SoA.
 *  - Pointers named *_host are host memory, *_dev are device pointers of the
 *    engine's device; `void* out` + `out_is_device` accepts either.
 *  - A handle is single-threaded; different handles may be used concurrently.
 *    All work is enqueued on the engine's stream; calls that return scalars
 *    or fill host buffers synchronise that stream, the others return after
 *    enqueue (like the reference: sync only at read-long / transfer).
 *  - There is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with BAY_ECUDA.
 */
#ifndef BAYADERA_B200_H
#define BAYADERA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bay_status {
    BAY_OK = 0,
    BAY_EINVAL = -1,          /* bad argument */
    BAY_EINVAL_WALKERS = -2,  /* G/:609-610 "Number of walkers (%d) must be a multiple of %d." */
    BAY_EACOR_TOO_SHORT = -3, /* G/:275-278 "The autocorrelation time is too long relative to the variance. ..." */
    BAY_ECOMPILE = -4,        /* NVRTC rejected the model source; log in bay_last_error() */
    BAY_ECUDA = -5,           /* CUDA driver/runtime error, or no device */
    BAY_ENCCL = -6,           /* NCCL error */
    BAY_ENOTSUP = -7
} bay_status;

typedef struct bay_engine bay_engine;   /* replaces GTXBayaderaFactory, G/:761-807 */
typedef struct bay_model bay_model;     /* replaces GTXStretchFactory + its NVRTC module, G/:543-610, 747-757 */
typedef struct bay_sampler bay_sampler; /* replaces GTXStretch, G/:282-541 */

/* model flags (extra keys of the model's args map; ignored by the reference) */
#define BAY_MODEL_FAST_MATH   0x1u  /* compile with -use_fast_math like G/:630-633 (default of the host mirror) */
#define BAY_MODEL_ROW_ADDITIVE 0x2u /* model also defines BAY_ROWLIK / BAY_PRIOR */
#define BAY_MODEL_GLM_LOGISTIC 0x4u /* data rows are [y, x_1..x_D]; likelihood is Bernoulli-logit of x·theta */
#define BAY_MODEL_GLM_POISSON 0x8u  /* data rows are [y, x_1..x_D]; likelihood is Poisson with log link: y*eta - exp(eta) */
#define BAY_MODEL_QUADFORM    0x10u /* Gaussian / quadratic form: params = [mu (D) | U (D x D row-major)], no data, and
                                       LOGFN(x) = -1/2 |U (x - mu)|^2.  D a multiple of 4, <= 128: moves of 128-walker
                                       tiles run as a dense contraction on the tensor cores; LOGFN
                                       itself still serves init-position!, run-sampler! and the density engine. */

const char *bay_last_error(void);
const char *bay_version(void);

/* ---- engine: gtx-bayadera-factory [ctx hstream compute-units WGS], G/:791-807 ----
 * device: CUDA ordinal.  stream: a CUstream/cudaStream_t handle to enqueue on,
 * or 0 to let the engine create its own non-blocking stream.  wgs: the
 * reference's work-group size = histogram bin count = accu block size
 * (power of two, 32..1024; the reference default is max-block-dim-x = 1024,
 * its tests use 256). */
int bay_engine_create(int device, uint64_t stream, int wgs, bay_engine **out);
/* The same engine inside the CUDA context that is CURRENT on the calling thread — the reference's situation:
 * ClojureCUDA's with-default creates a driver-API context (C/cuda.clj:27-31, C = the reference's
 * src/clojure/uncomplicate/bayadera/), gtx-bayadera-factory receives it (G/:791-807), every engine call runs inside
 * (in-context ctx ...) (G/:374, 402, ...) and parameter / result buffers are raw CUdeviceptr of that context
 * (Neanderthal cuda-float, G/:558, 798).  Nothing is bound to the primary context: the engine allocates, loads its
 * NVRTC modules and launches in the adopted context, so device pointers of that context can be passed to
 * bay_sampler_create_dev, bay_sample(out_is_device = 1) and bay_dataset_*(data_is_device = 1).  The caller keeps
 * ownership of the context and must release the engine's handles before destroying it.
 * Either way every entry point pushes the engine's context if it is not already current and pops it on return. */
int bay_engine_create_current(uint64_t stream, int wgs, bay_engine **out);
int bay_engine_release(bay_engine *e);
/* processing-elements, P/:138, G/:788-789  (= SM count * WGS) */
int bay_engine_processing_elements(bay_engine *e, int64_t *out);
int bay_engine_stream(bay_engine *e, uint64_t *stream_out);
int bay_engine_synchronize(bay_engine *e);

/* ---- multi-GPU: one process per GPU; no counterpart in the reference ----
 * The 128-byte ncclUniqueId is produced on rank 0 and distributed by the
 * caller (torch.distributed broadcast in the host mirror).  Call bay_engine_comm_init BEFORE compiling models.
 * Afterwards
 *   - GLM models (BAY_MODEL_GLM_*): walkers replicated, each rank passes ITS rows; per-walker sums are all-reduced;
 *   - every other model: ONE ensemble, rank r updates its slice of each half (results bit-identical to one GPU).
 *     Accepted walkers are stored by the kernel into every rank's ensemble over NVLink peer memory (CUDA IPC, up to
 *     8 ranks of one node; BAY_P2P=0 in the environment selects an NCCL all-gather exchange instead, BAY_PULL=1
 *     an experimental mode in which nothing is forwarded and partner rows are read from the owning rank).
 * Calls on such samplers — create, init-position!, burn-in!, run-sampler!, sample!, histogram!, mean, state
 * hand-off, release — are COLLECTIVE: every rank makes the same calls in the same order. */
int bay_nccl_unique_id(uint8_t id_out[128]);
int bay_engine_comm_init(bay_engine *e, const uint8_t id[128], int nranks, int rank);

/* ---- model: mcmc-factory [this model], P/:137; gtx-stretch-factory G/:747-757 ----
 * srcs: the model's (source) vector of C strings, concatenated in order;
 * logfn_name: (mcmc-logpdf model); dim: (dimension model);
 * params_size: (params-size model).  The source is compiled by NVRTC for
 * sm_100a together with the engine's stretch kernels. */
int bay_model_compile(bay_engine *e, const char *const *srcs, int nsrc, const char *logfn_name,
                      int dim, int params_size, uint32_t flags, bay_model **out);
int bay_model_release(bay_model *m);
/* device-less NVRTC build of the same program (the CPU "does it build" gate; log receives the
 * ptxas -v register/spill report, or the compiler errors). */
int bay_model_compile_check(const char *const *srcs, int nsrc, const char *logfn_name, int dim, int wgs,
                            uint32_t flags, int64_t *cubin_bytes, char *log_buf, int64_t log_cap);
/* 1 when moves of this model run on the tensor-core quadratic-form kernel (BAY_MODEL_QUADFORM accepted), else 0 */
int bay_model_uses_quadform(bay_model *m);
/* registers/spills/shared of the compiled stretch kernel (profiling aid) */
int bay_model_kernel_info(bay_model *m, const char *kernel, int *regs, int *local_bytes, int *smem_bytes);

/* ---- sampler: create-sampler [this seed walkers params], P/:120-121, G/:548-610 ----
 * params_host = [data (data_len) || hyperparams (params_size)], copied.
 * Fails with BAY_EINVAL_WALKERS unless walkers >= 2*wgs and walkers % (2*wgs) == 0.
 * With a communicator (bay_engine_comm_init) `walkers` is the GLOBAL count (a multiple of 2*wgs*nranks for
 * partitioned samplers).  An owned parameter vector of 64..16128 floats is also mirrored into __constant__ memory
 * and the sampler runs the program variant that reads it from there (BAY_CPARAMS=0 disables). */
int bay_sampler_create(bay_model *m, int32_t seed, int64_t walkers, const float *params_host,
                       int64_t params_count, bay_sampler **out);
/* same, with params already in device memory (the reference borrows a cuda-float vector) */
int bay_sampler_create_dev(bay_model *m, int32_t seed, int64_t walkers, uint64_t params_dev,
                           int64_t params_count, bay_sampler **out);
int bay_sampler_release(bay_sampler *s);

/* MCMC protocol, P/:97-103 */
int bay_init(bay_sampler *s, int32_t seed);                                          /* init!          G/:391-400 */
int bay_init_position_uniform(bay_sampler *s, int32_t seed, const float *limits_host); /* init-position! [seed limits] G/:409-418 */
int bay_init_position_from(bay_sampler *s, const bay_sampler *other);                /* init-position! [position]    G/:401-408 */
int bay_burn_in(bay_sampler *s, int64_t n, float a);                                 /* burn-in!       G/:419-429 */
/* anneal!: the Clojure schedule fn i -> T is pre-evaluated by the caller into temperature[n] */
int bay_anneal(bay_sampler *s, const float *temperature_host, int64_t n, float a);   /* anneal!        G/:430-440 */
int bay_acc_rate(bay_sampler *s, float a, double *out);                              /* acc-rate!      G/:466-478 */
/* run-sampler!: returns acceptance rate and Autocorrelation{tau mean sigma steps lag}
 * (tau/mean/sigma: DIM floats each, host). BAY_EACOR_TOO_SHORT when 5*lag > n. */
int bay_run_sampler(bay_sampler *s, int64_t n, float a, double *acc_rate, float *tau_host,
                    float *mean_host, float *sigma_host, int64_t *lag);              /* run-sampler!   G/:441-465 */
/* per-step ensemble means of the last run-sampler!/acc-rate! (DIM x n, column = step) */
int bay_last_means(bay_sampler *s, float *means_host, int64_t n);

/* MCMCStretch protocol, P/:105-109 */
int bay_init_move(bay_sampler *s, float a);          /* init-move!       G/:340-350 */
int bay_move(bay_sampler *s);                        /* move!            G/:351-357 */
int bay_move_bare(bay_sampler *s);                   /* move-bare!       G/:358-364 */
int bay_set_temperature(bay_sampler *s, float t);    /* set-temperature! G/:365-369 */
/* one raw launch of stretch_move_bare at the current counter, counter NOT advanced: the granularity of the
 * reference's kernel-level test (T/internal/nvidia_gtx_test.clj:217-235). half 0 = odd (s0), 1 = even (s1). */
int bay_move_bare_half(bay_sampler *s, int half);
int bay_set_a(bay_sampler *s, float a);              /* the `a` slot of the bare parameter packs, G/:423-424 */
/* raw results of the accu path (the reference's cu-accept / cu-means-acc buffers,
 * pinned at T/internal/nvidia_gtx_test.clj:251-254): per-block accept counts
 * (G = ceil(H/wgs) uint32) and the per-block position sums of the last move (DIM x G). */
int bay_accu_blocks(bay_sampler *s, uint32_t *accept_host, float *block_sums_host);

/* RandomSampler protocol, P/:94-95: sample! G/:371-389.  out is DIM x n. */
int bay_sample(bay_sampler *s, int64_t n, void *out, int out_is_device);

/* EstimateEngine / Location / Spread on the sampler, P/:20-28, 82-84; G/:482-541.
 * Histogram{limits 2xDIM, pdf WGSxDIM, bin-ranks WGSxDIM}, host fp32. */
int bay_histogram(bay_sampler *s, int64_t cycles, float *limits_host, float *pdf_host,
                  float *bin_ranks_host);
/* raw bin counts of the last bay_histogram (WGS x DIM uint32) — bit-exact parity hook */
int bay_histogram_counts(bay_sampler *s, uint32_t *counts_host);
int bay_mean(bay_sampler *s, float *out_host);
int bay_variance(bay_sampler *s, float *out_host);
int bay_sd(bay_sampler *s, float *out_host);

/* Info, G/:332-335 */
int bay_info(bay_sampler *s, int64_t *walkers, int64_t *iterations);

/* state hand-off (SURVEY §5 checkpoint/resume): positions DIM x W (host, AoS),
 * log-densities W, and the three integers that with the seed determine the
 * Philox stream: move-seed, move-bare-counter, move-counter. */
int bay_get_state(bay_sampler *s, float *xs_host, float *logfn_host, int32_t *bare_seed,
                  int32_t *move_seed, int64_t *bare_counter, int64_t *move_counter);
int bay_set_state(bay_sampler *s, const float *xs_host, const float *logfn_host /* NULL: recompute */,
                  int32_t bare_seed, int32_t move_seed, int64_t bare_counter, int64_t move_counter);
/* same hand-off with the log-densities in double: samplers of row-additive (GLM) models carry log-densities of
 * ~ -1e6..-1e7 whose O(1) differences fp32 cannot hold; with these nothing is recomputed on restore. */
int bay_get_state64(bay_sampler *s, float *xs_host, double *logfn64_host);
int bay_set_state64(bay_sampler *s, const float *xs_host, const double *logfn64_host);

/* ---- DatasetEngine / EstimateEngine on an arbitrary matrix, P/:71-73, 82-84; G/:145-228 ----
 * data: m x n, element (row, col) at data[offset + ld*col + row]. */
int bay_dataset_mean(bay_engine *e, const void *data, int data_is_device, int64_t m, int64_t n,
                     int64_t offset, int64_t ld, float *mean_host);
int bay_dataset_variance(bay_engine *e, const void *data, int data_is_device, int64_t m, int64_t n,
                         int64_t offset, int64_t ld, float *variance_host);
int bay_dataset_histogram(bay_engine *e, const void *data, int data_is_device, int64_t m, int64_t n,
                          int64_t offset, int64_t ld, float *limits_host, float *pdf_host,
                          float *bin_ranks_host, uint32_t *counts_host /* nullable */);

/* ---- AcorEngine, P/:86-87; G/:230-278.  series: dim x n (column = step), host. ---- */
int bay_acor(bay_engine *e, const float *series_host, int64_t dim, int64_t n, float *tau_host,
             float *mean_host, float *sigma_host, int64_t *lag);

/* ---- DensityEngine on the sampler's model: the logfn kernel over a point matrix
 * (K/engines/nvidia-gtx-mcmc-stretch.cu:197-206).  x: DIM x n. */
int bay_model_logfn(bay_model *m, const float *params_host, int64_t params_count,
                    const float *x_host, int64_t n, float *out_host);

/* ---- callers either side of the hot path ----
 * DensityEngine.log-density / density (P/:75-77; G/:67-131): the model is compiled with logfn_name = its logpdf
 * (distribution engine) or a loglik wrapper (likelihood engine, see bayadera_b200.models.likelihood_engine_model). */
int bay_model_density(bay_model *m, const float *params_host, int64_t params_count, const float *x_host,
                      int64_t n, int exponentiate, float *out_host);
/* LikelihoodEngine.evidence (P/:79-80; G/:132-141): mean over the n points of exp(logfn), double accumulation */
int bay_model_evidence(bay_model *m, const float *params_host, int64_t params_count, const float *x_host,
                       int64_t n, double *out);
/* the same with every block already in the engine's CUDA context (the reference's callers pass cuda-float blocks,
 * G/:83-104, 118-141): params_dev params_count floats, x_dev DIM x n column-major, out_dev n floats */
int bay_model_density_dev(bay_model *m, uint64_t params_dev, int64_t params_count, uint64_t x_dev, int64_t n,
                          int exponentiate, uint64_t out_dev);
int bay_model_evidence_dev(bay_model *m, uint64_t params_dev, int64_t params_count, uint64_t x_dev, int64_t n,
                           double *out);
/* RandomSamplerEngine.sample [seed params res] (P/:90-92; G/:48-63; K/rng/<family>-sampler.cu): family 0 uniform [a b],
 * 1 gaussian [mu sigma], 2 exponential [lambda], 3 erlang [lambda k]; n a multiple of 4; out: 1 x n. */
int bay_direct_sample(bay_engine *e, int family, int32_t seed, const float *params_host, int nparams, int64_t n,
                      void *out, int out_is_device);

/* ---- next row f-4: HDI on the device, mix! in one call ------------------------------------------------------
 * hdi (C/util.clj:102-110 = hdi-rank-count :52-65 + hdi-bins :67-83 + hdi-regions :85-100; C = the reference's
 * src/clojure/uncomplicate/bayadera/) for EVERY dimension of the sampler's latest histogram!, which is still on the
 * device.  counts[D]: number of ranked bins holding `mass`; nregions[D]; regions: D x (2 * max_regions) floats,
 * [lo0 hi0 lo1 hi1 ...] per dimension (regions beyond max_regions are counted, not stored). */
int bay_hdi(bay_sampler *s, double mass, int32_t *counts, int32_t *nregions, float *regions, int max_regions);
/* The same for caller-held Histogram columns (host arrays, column = dimension: limits 2 x dim, pdf and bin-ranks
 * bins x dim).  forced_counts != NULL with an entry >= 0 skips hdi-rank-count for that dimension (the
 * (hdi-regions limits bin-rank hdi-cnt) arity). */
int bay_hdi_histogram(bay_engine *e, int bins, int dim, const float *limits_host, const float *pdf_host,
                      const float *ranks_host, double mass, const int32_t *forced_counts, int32_t *counts,
                      int32_t *nregions, float *regions, int max_regions);
/* mix! (C/mcmc.clj:66-101): anneal step*D^dimension_power steps on `schedule` (0 minus-n, 1 sqrt-n, 2 pow-n with
 * schedule_power), tune a towards an acceptance rate in [min_acc, max_acc] with up to step+1 acc-rate! probes,
 * burn in.  One boundary crossing instead of the reference's ~70 protocol calls.  Outputs = the reference's map
 * {:a :acc-rate :acc-rate-2.0}. */
int bay_mix(bay_sampler *s, int64_t step, double dimension_power, int schedule, double schedule_power, double a,
            double min_acc, double max_acc, double *out_a, double *out_acc_rate, double *out_acc_rate_2);

/* ---- row-additive GLM path: precision probe (no counterpart in the reference, whose LOGFN loops over the dataset
 * serially in every thread, K/distributions/gaussian.cu:40-42) ----
 * sums_host[k] = sum over ALL ranks' rows of A(x_row . point_k) (A = softplus or exp) for n <= walkers caller points
 * (DIM x n, column = point).  method 0: the path the sampler itself runs (tcgen05 tensor cores when eligible),
 * 1: the fp32 SIMT kernel, 2: the same traversal in fp64 — the yardstick for the error of the Δlogp the accept test
 * consumes at full dataset size.  Collective on row-sharded samplers. */
int bay_glm_loglik_probe(bay_sampler *s, const float *points_host, int64_t n, int method, double *sums_host);

/* profiling counters: kernels launched by this library on the calling process */
int64_t bay_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BAYADERA_B200_H */
