// bayadera_b200 — C-ABI implementation (include/bayadera_b200.h).
//
// Host sequencing of the stretch-move sampler and its summary engines; the
// B200-native replacement of GTXStretch / GTXStretchFactory / GTXDatasetEngine /
// GTXAcorEngine in
//   G/ = /root/reference/src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj
// Model sources are compiled at run time by NVRTC for sm_100a together with
// stretch_program.inc; everything model-independent lives in kernels.cuh.
//
// libcuda / libnvrtc / libnccl are resolved lazily (dlopen) so that the
// library loads — and reports BAY_ECUDA loudly — on a machine without a GPU.
#include "../../include/bayadera_b200.h"
#include "kernels.cuh"
#include "kernels_glm_tc.cuh"
#include "kernels_quadform_tc.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvrtc.h>
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects itself

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

static const char* kStretchProgram =
#include "stretch_program.inc"
    ;
static const char* kGlmProgram =
#include "glm_program.inc"
    ;

// jitify-style stub so that `#include <stdint.h>` inside model sources resolves under NVRTC
// (the reference ships the same kind of stub, G/:647).
static const char* kStdintStub =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t;\n"
    "typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t;\n"
    "typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long long uintptr_t; typedef unsigned long size_t;\n";

// ------------------------------------------------------------------ errors --
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(BAY_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define CKLAUNCH()                                                                            \
    do {                                                                                      \
        g_launches++;                                                                         \
        cudaError_t e_ = cudaGetLastError();                                                  \
        if (e_ != cudaSuccess)                                                                \
            return fail(BAY_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),\
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define TRY(call)                    \
    do {                             \
        int r_ = (call);             \
        if (r_ != BAY_OK) return r_; \
    } while (0)

// Scratch device allocation that frees itself on every return path (the CK / TRY macros return early).
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// ------------------------------------------------------- lazy driver/nvrtc --
namespace {

struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*LaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned,
                                        unsigned, unsigned, CUstream, void**) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*CtxGetCurrent)(CUcontext*) = nullptr;
    CUresult (*CtxPushCurrent)(CUcontext) = nullptr;
    CUresult (*CtxPopCurrent)(CUcontext*) = nullptr;
    CUresult (*CtxGetDevice)(CUdevice*) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool ok = false;
};

struct NvrtcApi {
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                                 const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    bool ok = false;
    std::string why;
};

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

DriverApi g_cu;
NvrtcApi g_rtc;
NcclApi g_nccl;
std::mutex g_api_mutex;

template <typename F>
bool drv(const char* name, F* out) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(p);
    return true;
}

int load_driver() {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (g_cu.ok) return BAY_OK;
    bool ok = drv("cuModuleLoadData", &g_cu.ModuleLoadData) && drv("cuModuleUnload", &g_cu.ModuleUnload) &&
              drv("cuModuleGetFunction", &g_cu.ModuleGetFunction) &&
              drv("cuModuleGetGlobal", &g_cu.ModuleGetGlobal) && drv("cuLaunchKernel", &g_cu.LaunchKernel) &&
              drv("cuLaunchCooperativeKernel", &g_cu.LaunchCooperativeKernel) &&
              drv("cuFuncGetAttribute", &g_cu.FuncGetAttribute) &&
              drv("cuFuncSetAttribute", &g_cu.FuncSetAttribute) &&
              drv("cuOccupancyMaxActiveBlocksPerMultiprocessor",
                  &g_cu.OccupancyMaxActiveBlocksPerMultiprocessor) &&
              drv("cuGetErrorString", &g_cu.GetErrorString) && drv("cuCtxGetCurrent", &g_cu.CtxGetCurrent) &&
              drv("cuCtxPushCurrent", &g_cu.CtxPushCurrent) && drv("cuCtxPopCurrent", &g_cu.CtxPopCurrent) &&
              drv("cuCtxGetDevice", &g_cu.CtxGetDevice) &&
              drv("cuTensorMapEncodeTiled", &g_cu.TensorMapEncodeTiled);
    if (!ok) return fail(BAY_ECUDA, "CUDA driver entry points unavailable (no NVIDIA driver / GPU?)");
    g_cu.ok = true;
    return BAY_OK;
}

void* open_first(const char* const* names, std::string* why) {
    for (int i = 0; names[i]; i++) {
        void* h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
        *why += std::string(names[i]) + ": " + dlerror() + "; ";
    }
    return nullptr;
}

int load_nvrtc() {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (g_rtc.ok) return BAY_OK;
    static const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", nullptr};
    g_rtc.why.clear();
    void* h = open_first(names, &g_rtc.why);
    if (!h) return fail(BAY_ECOMPILE, "cannot load NVRTC: %s", g_rtc.why.c_str());
#define RTC(sym)                                                                     \
    g_rtc.sym = reinterpret_cast<decltype(g_rtc.sym)>(dlsym(h, "nvrtc" #sym));       \
    if (!g_rtc.sym) return fail(BAY_ECOMPILE, "NVRTC symbol nvrtc" #sym " missing");
    RTC(CreateProgram) RTC(CompileProgram) RTC(GetProgramLogSize) RTC(GetProgramLog) RTC(GetCUBINSize)
    RTC(GetCUBIN) RTC(DestroyProgram) RTC(GetErrorString)
#undef RTC
    g_rtc.ok = true;
    return BAY_OK;
}

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_api_mutex);
    if (g_nccl.ok) return BAY_OK;
    static const char* names[] = {"libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", nullptr};
    g_nccl.why.clear();
    void* h = open_first(names, &g_nccl.why);
    if (!h) return fail(BAY_ENCCL, "cannot load NCCL: %s", g_nccl.why.c_str());
#define NC(sym)                                                                      \
    g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(h, "nccl" #sym));      \
    if (!g_nccl.sym) return fail(BAY_ENCCL, "NCCL symbol nccl" #sym " missing");
    NC(GetUniqueId) NC(CommInitRank) NC(CommDestroy) NC(AllGather) NC(AllReduce) NC(GroupStart) NC(GroupEnd)
    NC(GetErrorString)
#undef NC
    g_nccl.ok = true;
    return BAY_OK;
}

int cu_fail(CUresult r, const char* what) {
    const char* s = nullptr;
    if (g_cu.GetErrorString) g_cu.GetErrorString(r, &s);
    return fail(BAY_ECUDA, "%s failed: %s", what, s ? s : "unknown CUDA driver error");
}

#define CKNCCL(call)                                                                           \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) return fail(BAY_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

#define BAY_GLM_ANY (BAY_MODEL_GLM_LOGISTIC | BAY_MODEL_GLM_POISSON)

// ----------------------------------------------------------------- handles --
struct bay_engine {
    int device = 0;
    CUcontext ctx = nullptr;      // the context every call of this engine runs in (the caller's, or the device's primary)
    bool borrowed_ctx = false;    // created by bay_engine_create_current: the caller owns the context
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int wgs = 256;
    int sm_count = 0;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    uint64_t tc_configured = 0;   // kernel variants whose dynamic-shared-memory opt-in is set on this device
};

struct bay_sampler;

struct bay_model {
    bay_engine* e = nullptr;
    CUmodule mod = nullptr;
    CUfunction f_bare = nullptr, f_accu = nullptr, f_logfn = nullptr, f_loop = nullptr, f_accu_loop = nullptr;
    int accu_loop_capacity = 0;   // co-resident CTAs (WGS threads) of the persistent run-sampler! kernel
    int loop_capacity = 0;   // co-resident CTAs of the persistent step-loop kernel on this device
    int loop_block = 128;    // its CTA size
    // GLM (row-additive Bernoulli-logit) path, present iff `glm`
    CUfunction f_glm_propose = nullptr, f_glm_loglik = nullptr, f_glm_lp_init = nullptr, f_glm_accept = nullptr;
    CUfunction f_glm_loglik64 = nullptr;   // fp64 yardstick of the likelihood (bay_glm_loglik_probe)
    bool glm = false;
    int glm_link = 0;      // 0 Bernoulli-logit (softplus), 1 Poisson-log (exp)
    bool rowadd = false;   // generic row-additive model (BAY_MODEL_ROW_ADDITIVE without a GLM flag): rides the GLM host path
    uint32_t row_stride = 1;                  // its BAY_ROW_STRIDE (read back from the compiled module)
    CUfunction f_rowadd_loglik = nullptr;
    bool quadform = false; // BAY_MODEL_QUADFORM: logp = -1/2 |U (x - mu)|^2, moves run on k_quadform_move_tc
    bool mirror = false;   // AoS mirror of the ensemble for the partner gather (DIM >= 4, non-GLM)
    bool peers = false;    // kernels store accepted walkers into every rank's ensemble block (multi-GPU mode A)
    bool pull = false;     // ... or (experimental, BAY_PULL=1) nothing is stored remotely and partner rows are pulled
    // Constant-parameter variant (stretch_program.inc, BAY_CPARAMS): the same program compiled with the parameter
    // vector in __constant__ memory; built on demand for samplers whose parameters fit (cparams_try)
    CUmodule cmod = nullptr;
    CUfunction c_bare = nullptr, c_accu = nullptr, c_logfn = nullptr, c_loop = nullptr, c_accu_loop = nullptr;
    int c_loop_capacity = 0, c_accu_loop_capacity = 0;
    int cvar_state = 0;                       // 0 not tried, 1 built, -1 unavailable
    CUdeviceptr cparams = 0;
    const bay_sampler* cparams_owner = nullptr;   // whose parameters the constant block holds right now
    std::vector<std::string> srcs;            // what bay_model_compile was given (to build the variant later)
    std::string logfn_name;
    int dima = 1;          // mirror row length: DIM rounded up to a multiple of 4
    int dim = 1, params_size = 0;
    uint32_t flags = 0;
    int block = 128;  // bare/logfn block size
};

struct bay_sampler {
    bay_model* m = nullptr;
    int64_t W = 0, H = 0;
    int D = 1;
    uint32_t G = 0;  // accu blocks per half
    float* params = nullptr;
    bool own_params = false;
    uint32_t data_len = 0, params_len = 0;
    float* xs = nullptr;  // D x W SoA, pitch W
    float* lp = nullptr;  // W
    float* xa = nullptr;  // W x dima AoS mirror (only if m->mirror)
    unsigned int* loop_bar = nullptr;   // grid-barrier words of the persistent step loop
    // multi-GPU mode A over peer memory: xs | lp | xa | flags live in ONE block that every rank maps (CUDA IPC)
    float* peer_block = nullptr;
    bay::PeerTable peer_tab;            // block base on every rank + offsets; passed by value to the kernels
    uint64_t peer_flags_off = 0;        // offset (in 4-byte words) of the barrier flags inside the block
    uint32_t peer_epoch = 0;
    bool cp = false;                    // runs the constant-parameter kernel variant
    bool remote_stale = false;          // pull mode: other ranks' slices of the local copy are out of date
    bool soa_stale = false;             // peers forwarded mirror rows only: rebuild xs from xa before a read-out
    bool soa_own_stale = false;         // the tensor-core move updates ONLY the mirror: xs is stale everywhere, also
                                        // for this rank's own walkers (generic kernels must wait for a rebuild)
    bool qf_configured = false;         // dynamic shared-memory opt-in of k_quadform_move_tc done on this device
    void* peer_mapped[8] = {nullptr};   // what cudaIpcOpenMemHandle returned (to close on release)
    float* loop_betas = nullptr;        // per-step inverse temperatures (anneal!)
    int64_t loop_betas_cap = 0;
    uint32_t* accept = nullptr;               // G
    uint32_t* accept_all = nullptr;           // G: all-reduced copy (walker-partitioned multi-GPU)
    float* blk_sums = nullptr;                // D x G
    unsigned long long* accept_total = nullptr;  // 1
    float* means = nullptr;                   // D x means_cap
    int64_t means_cap = 0, means_n = 0;
    uint32_t* hist_counts = nullptr;          // wgs x D
    uint32_t* mm = nullptr;                   // 2 x D ordered-uint min/max
    float* limits = nullptr;                  // 2 x D
    float* pdf = nullptr;                     // wgs x D
    float* ranks = nullptr;                   // wgs x D
    double* macc = nullptr;                   // D x 2
    float* vec_d = nullptr;                   // 4 x D scratch
    float* stage = nullptr;                   // D x W AoS staging for state hand-off (lazy)
    float* sample_stage = nullptr;            // staging of sample! results bound for host memory (lazy, grown on demand)
    size_t sample_stage_cap = 0;
    int64_t params_count = 0;                 // floats in `params` (may be fewer than data_len + params_len)
    uint64_t version = 1;                     // bumped by everything that changes the ensemble
    uint64_t macc_version = 0;                // ensemble version the moment sums in `macc` were taken from (0: none)
    bool macc_aos = false;                    // ... and whether they were taken from the AoS mirror
    // GLM path: repacked dataset, proposals and double-precision log-densities
    uint64_t glm_rows = 0;                    // local rows (this rank's shard)
    uint64_t glm_rows_total = 0;              // rows of all shards
    bool own_glm_x = true;                    // glm_x is an allocation of the sampler (generic row-additive: it is `params`)
    float* glm_x = nullptr;                   // rows x D row-major
    double* glm_sy = nullptr;                 // D: X^T y (all-reduced over row shards)
    double* glm_sx = nullptr;                 // D: column sums of the LOCAL rows
    float* glm_yt = nullptr;                  // D x W proposals / points, SoA
    float* glm_z = nullptr;                   // H
    float* glm_u = nullptr;                   // H
    double* glm_partial = nullptr;            // chunks x W
    double* glm_sp = nullptr;                 // W: sum_rows softplus
    double* lp64 = nullptr;                   // W
    uint32_t glm_chunks = 0, glm_rows_per_chunk = 0;
    // tensor-core variant (DIM <= 64): bf16 hi/lo planes of the dataset and of the walker block + TMA maps
    bool glm_tc = false;
    uint32_t glm_nkc = 1;                     // 64-wide K chunks per row (DIM <= 64: 1, <= 128: 2)
    __nv_bfloat16 *glm_xh = nullptr, *glm_xl = nullptr, *glm_ah = nullptr, *glm_am = nullptr, *glm_al = nullptr;
    int glm_terms = 3;                        // MMAs per K step: 3 = fp16 pieces (dh+dl).Xh + dh.Xl; 4 = bf16 pieces
                                              // (dh+dm+dl).Xh + dh.Xl, 5 adds dm.Xl  (BAY_GLM_TERMS)
    float* glm_cscale = nullptr;              // fp16 path: per column (scale, 1 / scale) of the dataset planes
    float* glm_dscale = nullptr;              // fp16 path: {scale, 1 / scale} of the current call's delta planes
    float* glm_theta0 = nullptr;              // D: reference point of the tensor-core contraction (kernels_glm_tc.cuh)
    float* glm_eta0 = nullptr;                // rows padded to a tile: log2(e) * x_row . theta0
    float* glm_mean = nullptr;                // D: mean of the points of the current likelihood call
    int* glm_ref_changed = nullptr;           // device flag: the reference point moved, eta0 must be recomputed
    bool glm_ref_init = false;                // theta0 / eta0 have been computed at least once
    CUtensorMap glm_map_xh, glm_map_xl;
    struct GlmWalkerMaps { CUtensorMap hi, mid, lo; };
    std::map<uint64_t, GlmWalkerMaps> glm_amaps;   // (first walker, count) -> maps of the three theta planes
    // host-side counters: G/:282-287, 340-400
    int32_t bare_seed = 0, move_seed = 0;
    uint32_t bare_counter = 0, move_counter = 0;
    int64_t iterations = 0;
    float a_bare = 2.0f, a_move = 2.0f, beta = 1.0f;
};

// ---------------------------------------------------------------- plumbing --
// Every entry point runs inside the engine's context, the way the reference wraps each engine call in
// (in-context ctx ...) (G/:374, 402, ...): the context is pushed unless it is already current and popped on return,
// so the caller's context stack is left as it was found.  No entry point calls cudaSetDevice (which would swap the
// caller's context for the primary one); runtime-API calls below operate on whatever context is current.
struct CtxScope {
    int rc = BAY_OK;
    bool pushed = false;
    explicit CtxScope(const bay_engine* e) {
        if (!e || !e->ctx || !g_cu.ok) { rc = fail(BAY_EINVAL, "engine has no CUDA context"); return; }
        CUcontext cur = nullptr;
        if (g_cu.CtxGetCurrent(&cur) == CUDA_SUCCESS && cur == e->ctx) return;
        const CUresult r = g_cu.CtxPushCurrent(e->ctx);
        if (r != CUDA_SUCCESS) rc = cu_fail(r, "cuCtxPushCurrent");
        else pushed = true;
    }
    ~CtxScope() {
        CUcontext c = nullptr;
        if (pushed) g_cu.CtxPopCurrent(&c);
    }
    CtxScope(const CtxScope&) = delete;
    CtxScope& operator=(const CtxScope&) = delete;
};
// NVTX range named after the C-ABI entry point (the reference has no tracing at all): shows up as
// bay_burn_in / bay_histogram / ... on the nsys / ncu timeline.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define USE_ENGINE(e)                  \
    NvtxRange nvtx_range_(__func__);   \
    CtxScope ctx_scope_(e);            \
    if (ctx_scope_.rc != BAY_OK) return ctx_scope_.rc

extern "C" const char* bay_last_error(void) { return g_err.c_str(); }
extern "C" const char* bay_version(void) { return "bayadera_b200 0.1 (sm_100a)"; }
extern "C" int64_t bay_launch_count(void) { return g_launches.load(); }

static int engine_finish_create(CUcontext ctx, bool borrowed, int device, uint64_t stream, int wgs, bay_engine** out) {
    bay_engine* e = new bay_engine();
    e->device = device;
    e->ctx = ctx;
    e->borrowed_ctx = borrowed;
    e->wgs = wgs;
    CtxScope scope(e);
    cudaError_t ce = scope.rc == BAY_OK ? cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device)
                                        : cudaErrorInvalidValue;
    if (ce == cudaSuccess) {
        if (stream) {
            e->stream = reinterpret_cast<cudaStream_t>(stream);
        } else {
            ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
            e->own_stream = ce == cudaSuccess;
        }
    }
    if (ce != cudaSuccess) {
        delete e;
        return scope.rc != BAY_OK ? scope.rc : fail(BAY_ECUDA, "engine creation failed: %s", cudaGetErrorString(ce));
    }
    *out = e;
    return BAY_OK;
}

extern "C" int bay_engine_create(int device, uint64_t stream, int wgs, bay_engine** out) {
    if (!out) return fail(BAY_EINVAL, "out is NULL");
    if (wgs < 32 || wgs > 1024 || (wgs & (wgs - 1))) return fail(BAY_EINVAL, "wgs must be a power of two in [32, 1024], got %d", wgs);
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(BAY_ECUDA, "no CUDA device available (%s); bayadera_b200 has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    }
    if (device < 0 || device >= count) return fail(BAY_EINVAL, "device %d out of range [0,%d)", device, count);
    // this entry point asks for the device's PRIMARY context (what torch and the runtime API use); it is bound here,
    // once, and pushed by every later call if the caller has something else current
    CK(cudaSetDevice(device));
    CK(cudaFree(0));
    TRY(load_driver());
    CUcontext ctx = nullptr;
    CUresult cr = g_cu.CtxGetCurrent(&ctx);
    if (cr != CUDA_SUCCESS || !ctx) return fail(BAY_ECUDA, "no primary context on device %d", device);
    return engine_finish_create(ctx, false, device, stream, wgs, out);
}

// The reference never touches the primary context: ClojureCUDA creates a driver-API context (with-default,
// C/cuda.clj:27-31), every engine call runs (in-context ctx ...) and parameters / results are raw CUdeviceptr of THAT
// context (Neanderthal cuda-float, G/:558, 798).  This constructor adopts the context that is current on the calling
// thread, whatever kind it is; the engine allocates, loads its modules and launches in it, so borrowed device
// pointers (bay_sampler_create_dev, bay_sample(out_is_device = 1), bay_dataset_*(data_is_device = 1)) are valid.
extern "C" int bay_engine_create_current(uint64_t stream, int wgs, bay_engine** out) {
    if (!out) return fail(BAY_EINVAL, "out is NULL");
    if (wgs < 32 || wgs > 1024 || (wgs & (wgs - 1))) return fail(BAY_EINVAL, "wgs must be a power of two in [32, 1024], got %d", wgs);
    if (load_driver() != BAY_OK)
        return fail(BAY_ECUDA, "no CUDA device available (driver entry points missing); bayadera_b200 has no CPU fallback");
    CUcontext ctx = nullptr;
    CUresult cr = g_cu.CtxGetCurrent(&ctx);
    if (cr != CUDA_SUCCESS) return cu_fail(cr, "cuCtxGetCurrent");
    if (!ctx) return fail(BAY_ECUDA, "bay_engine_create_current: no CUDA context is current on the calling thread");
    CUdevice dev = 0;
    cr = g_cu.CtxGetDevice(&dev);
    if (cr != CUDA_SUCCESS) return cu_fail(cr, "cuCtxGetDevice");
    return engine_finish_create(ctx, true, (int)dev, stream, wgs, out);
}

extern "C" int bay_engine_release(bay_engine* e) {
    if (!e) return BAY_OK;
    {
        CtxScope scope(e);
        if (e->comm && g_nccl.ok) g_nccl.CommDestroy(e->comm);
        if (e->own_stream) cudaStreamDestroy(e->stream);
    }
    delete e;
    return BAY_OK;
}

extern "C" int bay_engine_processing_elements(bay_engine* e, int64_t* out) {
    if (!e || !out) return fail(BAY_EINVAL, "NULL argument");
    *out = (int64_t)e->sm_count * e->wgs;
    return BAY_OK;
}

extern "C" int bay_engine_stream(bay_engine* e, uint64_t* s) {
    if (!e || !s) return fail(BAY_EINVAL, "NULL argument");
    *s = reinterpret_cast<uint64_t>(e->stream);
    return BAY_OK;
}

extern "C" int bay_engine_synchronize(bay_engine* e) {
    if (!e) return fail(BAY_EINVAL, "NULL engine");
    USE_ENGINE(e);
    CK(cudaStreamSynchronize(e->stream));
    return BAY_OK;
}

extern "C" int bay_nccl_unique_id(uint8_t id_out[128]) {
    TRY(load_nccl());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    CKNCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return BAY_OK;
}

extern "C" int bay_engine_comm_init(bay_engine* e, const uint8_t id[128], int nranks, int rank) {
    if (!e || !id) return fail(BAY_EINVAL, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(BAY_EINVAL, "bad rank %d / nranks %d", rank, nranks);
    TRY(load_nccl());
    USE_ENGINE(e);
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    CKNCCL(g_nccl.CommInitRank(&e->comm, nranks, uid, rank));
    e->nranks = nranks;
    e->rank = rank;
    return BAY_OK;
}

// ------------------------------------------------------------------- model --
// Walker-partitioned samplers exchange through peer memory (NVLink stores from the stretch kernel itself) unless
// BAY_P2P=0 asks for the NCCL all-gather exchange instead (the baseline the peer path is measured against).
static bool engine_wants_peers(const bay_engine* e) {
    if (!e->comm || e->nranks < 2 || e->nranks > 8) return false;
    const char* env = getenv("BAY_P2P");
    return !(env && env[0] == '0');
}

// models whose half-step is propose -> dataset likelihood -> accept (glm_program.inc): the two GLM families on
// tensor cores / SIMT tiles, and any model that states its row decomposition (BAY_MODEL_ROW_ADDITIVE)
static bool is_rowadd_generic(uint32_t flags) { return (flags & BAY_MODEL_ROW_ADDITIVE) && !(flags & BAY_GLM_ANY); }
static bool is_dataset_model(int dim, uint32_t flags) {
    return ((flags & BAY_GLM_ANY) && dim % 4 == 0) || is_rowadd_generic(flags);
}

// The AoS mirror pays off once a walker spans several 32-byte sectors; GLM models gather only H rows per half-step.
// BAY_MIRROR=0 in the environment disables it (A/B measurements).
static bool model_wants_mirror(int dim, uint32_t flags) {
    if (dim < 4) return false;
    if (is_dataset_model(dim, flags)) return false;
    const char* env = getenv("BAY_MIRROR");
    return !(env && env[0] == '0');
}

// CTA size of the persistent step-loop kernel: as large as the register budget of a DIM-float proposal allows,
// so that few CTAs arrive at each grid barrier.
static int loop_block_for(int dim, int block) { return dim <= 2 ? 1024 : (dim <= 8 ? 512 : block); }

static bool engine_wants_pull(const bay_engine* e) {
    const char* env = getenv("BAY_PULL");
    return engine_wants_peers(e) && env && env[0] == '1';
}

// NVRTC step shared by bay_model_compile and bay_model_compile_check.
// gtx-stretch-factory (G/:747-757): model sources first, engine kernels after;
// stretch-options (G/:630-633) retargeted to sm_100a.
static int nvrtc_build(const char* const* srcs, int nsrc, const char* logfn_name, int dim, int wgs, int block,
                       uint32_t flags, bool peers, bool verbose, std::vector<char>* cubin, std::string* log_out,
                       int cparams = 0, bool pull = false) {
    if (!srcs || !logfn_name) return fail(BAY_EINVAL, "NULL argument");
    if (dim < 1 || dim > 4096) return fail(BAY_EINVAL, "dimension %d out of range", dim);
    if (wgs < 32 || wgs > 1024 || (wgs & (wgs - 1))) return fail(BAY_EINVAL, "wgs must be a power of two in [32, 1024], got %d", wgs);
    TRY(load_nvrtc());
    std::string src = "#include <stdint.h>\n";
    for (int i = 0; i < nsrc; i++) {
        if (!srcs[i]) return fail(BAY_EINVAL, "srcs[%d] is NULL", i);
        src += srcs[i];
        src += "\n";
    }
    src += kStretchProgram;
    if (is_dataset_model(dim, flags)) src += kGlmProgram;
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-default-device", "-lineinfo", "--std=c++17",
                                     "-DREAL=float", "-DREAL2=float2", "-DACCUMULATOR=float",
                                     "-DLOGFN=" + std::string(logfn_name), "-DDIM=" + std::to_string(dim),
                                     "-DWGS=" + std::to_string(wgs), "-DBAY_BLOCK=" + std::to_string(block)};
    if (flags & BAY_MODEL_FAST_MATH) opts.push_back("-use_fast_math");
    if (model_wants_mirror(dim, flags)) {
        opts.push_back("-DBAY_MIRROR=1");
        opts.push_back("-DBAY_DIMA=" + std::to_string((dim + 3) / 4 * 4));
    }
    if (peers) opts.push_back("-DBAY_PEERS=1");
    if (peers && pull) opts.push_back("-DBAY_PULL=1");
    if (flags & BAY_MODEL_GLM_POISSON) opts.push_back("-DBAY_GLM_LINK=1");
    if (is_rowadd_generic(flags)) opts.push_back("-DBAY_ROWADD=1");
    if (cparams > 0) opts.push_back("-DBAY_CPARAMS=" + std::to_string(cparams));
    // ensemble traffic policy (stretch_program.inc, BAY_STREAM): large-DIM models keep a per-thread local array and
    // their parameter block in L1, which the once-touched ensemble data would otherwise evict
    {
        int stream = dim >= 16 ? 1 : 0;
        if (const char* env = getenv("BAY_STREAM")) stream = atoi(env);
        if (stream) opts.push_back("-DBAY_STREAM=" + std::to_string(stream));
    }
    // Occupancy floor: a thread holds the whole proposal (DIM floats), so for large DIM ptxas takes 255 registers
    // and only 8 warps fit an SM — the kernel then stalls on load latency (ncu: long_scoreboard).  Asking for more
    // resident CTAs caps the registers and trades L1-resident spills for more warps.  BAY_MINB overrides.
    {
        int minb = 1;   // measured on DIM = 100: 1 and 8 tie (0.86 ms), 4 is worse (L1 thrash by local arrays)
        // with the parameters out of the LSU's way (BAY_CPARAMS) 3 CTAs = 12 warps at 168 registers is the optimum for
        // DIM = 100: 1.24 G walker-steps/s against 0.85 G at 8 warps and 0.38 G at 16 (128 registers: spills)
        if (cparams > 0 && dim >= 64) minb = 3;
        if (const char* env = getenv("BAY_MINB")) minb = atoi(env);
        if (minb > 1) opts.push_back("-DBAY_MINB=" + std::to_string(minb));
    }
    opts.push_back("-DBAY_LOOP_BLOCK=" + std::to_string(loop_block_for(dim, block)));
    if (verbose) opts.push_back("--ptxas-options=-v");
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());

    nvrtcProgram prog;
    const char* hdr_src[] = {kStdintStub};
    const char* hdr_name[] = {"stdint.h"};
    nvrtcResult r = g_rtc.CreateProgram(&prog, src.c_str(), "bayadera_stretch.cu", 1, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) return fail(BAY_ECOMPILE, "nvrtcCreateProgram: %s", g_rtc.GetErrorString(r));
    r = g_rtc.CompileProgram(prog, (int)copts.size(), copts.data());
    size_t n = 0;
    g_rtc.GetProgramLogSize(prog, &n);
    std::string log(n + 1, '\0');
    if (n) g_rtc.GetProgramLog(prog, &log[0]);
    if (log_out) *log_out = log.c_str();
    if (r != NVRTC_SUCCESS) {
        g_rtc.DestroyProgram(&prog);
        return fail(BAY_ECOMPILE, "NVRTC compilation of model '%s' failed: %s\n%s", logfn_name,
                    g_rtc.GetErrorString(r), log.c_str());
    }
    size_t cs = 0;
    g_rtc.GetCUBINSize(prog, &cs);
    cubin->resize(cs);
    g_rtc.GetCUBIN(prog, cubin->data());
    g_rtc.DestroyProgram(&prog);
    return BAY_OK;
}

static int bare_block_for(int dim) { return dim <= 8 ? 256 : 128; }

extern "C" int bay_model_compile_check(const char* const* srcs, int nsrc, const char* logfn_name, int dim, int wgs,
                                       uint32_t flags, int64_t* cubin_bytes, char* log_buf, int64_t log_cap) {
    std::vector<char> cubin;
    std::string log;
    int r = nvrtc_build(srcs, nsrc, logfn_name, dim, wgs, bare_block_for(dim), flags, false, true, &cubin, &log);
    if (log_buf && log_cap > 0) {
        const std::string& text = r == BAY_OK ? log : g_err;
        snprintf(log_buf, (size_t)log_cap, "%s", text.c_str());
    }
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return r;
}

extern "C" int bay_model_compile(bay_engine* e, const char* const* srcs, int nsrc, const char* logfn_name,
                                 int dim, int params_size, uint32_t flags, bay_model** out) {
    if (!e || !out) return fail(BAY_EINVAL, "NULL argument");
    USE_ENGINE(e);
    TRY(load_driver());
    std::vector<char> cubin;
    const bool glm_model = is_dataset_model(dim, flags);
    const bool peers = engine_wants_peers(e) && !glm_model;
    const bool pull = peers && engine_wants_pull(e);
    TRY(nvrtc_build(srcs, nsrc, logfn_name, dim, e->wgs, bare_block_for(dim), flags, peers, false, &cubin, nullptr, 0, pull));

    bay_model* m = new bay_model();
    m->e = e;
    m->dim = dim;
    m->params_size = params_size;
    m->flags = flags;
    m->block = bare_block_for(dim);
    m->mirror = model_wants_mirror(dim, flags);
    {
        const char* env = getenv("BAY_QUADFORM_TC");
        m->quadform = (flags & BAY_MODEL_QUADFORM) && m->mirror && dim % 4 == 0 && dim <= bay::qf::MAX_D &&
                      params_size == dim + dim * dim && !(env && env[0] == '0');
    }
    m->peers = peers;
    m->pull = pull;
    m->dima = (dim + 3) / 4 * 4;
    for (int i = 0; i < nsrc; i++) m->srcs.emplace_back(srcs[i]);
    m->logfn_name = logfn_name;

    CUresult cr = g_cu.ModuleLoadData(&m->mod, cubin.data());
    if (cr != CUDA_SUCCESS) {
        delete m;
        return cu_fail(cr, "cuModuleLoadData");
    }
    struct { const char* name; CUfunction* f; } fns[] = {
        {"bay_stretch_bare", &m->f_bare}, {"bay_stretch_accu", &m->f_accu}, {"bay_logfn", &m->f_logfn},
        {"bay_stretch_loop", &m->f_loop}, {"bay_stretch_accu_loop", &m->f_accu_loop}};
    for (auto& fn : fns) {
        cr = g_cu.ModuleGetFunction(fn.f, m->mod, fn.name);
        if (cr != CUDA_SUCCESS) {
            g_cu.ModuleUnload(m->mod);
            delete m;
            return cu_fail(cr, fn.name);
        }
    }
    {
        int per_sm = 0;
        m->loop_block = loop_block_for(dim, m->block);
        if (g_cu.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->f_loop, m->loop_block, 0) == CUDA_SUCCESS)
            m->loop_capacity = per_sm * e->sm_count;
        if (g_cu.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->f_accu_loop, e->wgs, 0) == CUDA_SUCCESS)
            m->accu_loop_capacity = per_sm * e->sm_count;
    }
    if (is_dataset_model(dim, flags)) {
        m->rowadd = is_rowadd_generic(flags);
        std::vector<std::pair<const char*, CUfunction*>> gfns = {
            {"bay_glm_propose", &m->f_glm_propose}, {"bay_glm_lp_init", &m->f_glm_lp_init},
            {"bay_glm_accept", &m->f_glm_accept}};
        if (m->rowadd) {
            gfns.push_back({"bay_rowadd_loglik", &m->f_rowadd_loglik});
        } else {
            gfns.push_back({"bay_glm_loglik", &m->f_glm_loglik});
            gfns.push_back({"bay_glm_loglik_f64", &m->f_glm_loglik64});
        }
        for (auto& fn : gfns) {
            cr = g_cu.ModuleGetFunction(fn.second, m->mod, fn.first);
            if (cr != CUDA_SUCCESS) {
                g_cu.ModuleUnload(m->mod);
                delete m;
                return cu_fail(cr, fn.first);
            }
        }
        if (m->rowadd) {   // the row stride the model declared (BAY_ROW_STRIDE), exported by the program as a global
            CUdeviceptr dptr = 0;
            size_t bytes = 0;
            cr = g_cu.ModuleGetGlobal(&dptr, &bytes, m->mod, "bay_row_stride");
            cudaError_t ce = cr == CUDA_SUCCESS ? cudaMemcpy(&m->row_stride, reinterpret_cast<void*>(dptr), sizeof(uint32_t), cudaMemcpyDeviceToHost)
                                                : cudaErrorUnknown;
            if (ce != cudaSuccess || m->row_stride < 1) {
                g_cu.ModuleUnload(m->mod);
                delete m;
                return fail(BAY_ECOMPILE, "row-additive model: cannot read BAY_ROW_STRIDE from the compiled program");
            }
        }
        m->glm = true;
        m->glm_link = (flags & BAY_MODEL_GLM_POISSON) ? 1 : 0;
    }
    *out = m;
    return BAY_OK;
}

extern "C" int bay_model_release(bay_model* m) {
    if (!m) return BAY_OK;
    CtxScope scope(m->e);
    if (m->mod && g_cu.ok) g_cu.ModuleUnload(m->mod);
    if (m->cmod && g_cu.ok) g_cu.ModuleUnload(m->cmod);
    delete m;
    return BAY_OK;
}

extern "C" int bay_model_uses_quadform(bay_model* m) { return m && m->quadform ? 1 : 0; }

extern "C" int bay_model_kernel_info(bay_model* m, const char* kernel, int* regs, int* local_bytes, int* smem_bytes) {
    if (!m || !kernel) return fail(BAY_EINVAL, "NULL argument");
    USE_ENGINE(m->e);
    CUfunction f = nullptr;
    CUresult cr = g_cu.ModuleGetFunction(&f, m->cvar_state == 1 ? m->cmod : m->mod, kernel);   // the variant samplers run
    if (cr != CUDA_SUCCESS) return cu_fail(cr, kernel);
    if (regs) g_cu.FuncGetAttribute(regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f);
    if (local_bytes) g_cu.FuncGetAttribute(local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, f);
    if (smem_bytes) g_cu.FuncGetAttribute(smem_bytes, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f);
    return BAY_OK;
}

// ---- constant-parameter kernel variant ------------------------------------------------------------------------
// A warp evaluates LOGFN for 32 walkers against ONE parameter vector: every lane reads the same address.  Through
// the LSU such loads cost a wavefront per 4 bytes (ncu on the 100-D Gaussian: 93 wavefronts per 64 FFMAs, L1 data
// pipe 59 % busy, 0.75 G walker-steps/s); from __constant__ memory they become uniform-datapath loads / FFMA
// operands (LDCU + FFMA R, R, UR, R) and the same model runs at 1.24 G.  The variant is compiled on first use for
// samplers that OWN a host-supplied parameter vector of 64..16128 floats (a borrowed device vector may be rewritten
// by its owner between calls, which a constant copy would not see).  BAY_CPARAMS=0 disables it.
static const int kCparamsCap = 16128;   // floats; the 64 KB constant bank minus what the program itself needs

static void cparams_try(bay_sampler* s, int64_t params_count) {
    bay_model* m = s->m;
    if (m->glm || params_count < 64 || params_count > kCparamsCap) return;
    if (const char* env = getenv("BAY_CPARAMS")) if (env[0] == '0') return;
    if (m->cvar_state == 0) {
        m->cvar_state = -1;
        std::vector<const char*> srcs;
        for (auto& t : m->srcs) srcs.push_back(t.c_str());
        std::vector<char> cubin;
        if (nvrtc_build(srcs.data(), (int)srcs.size(), m->logfn_name.c_str(), m->dim, m->e->wgs, m->block, m->flags, m->peers,
                        false, &cubin, nullptr, kCparamsCap, m->pull) != BAY_OK) return;
        if (g_cu.ModuleLoadData(&m->cmod, cubin.data()) != CUDA_SUCCESS) { m->cmod = nullptr; return; }
        size_t bytes = 0;
        bool ok = g_cu.ModuleGetFunction(&m->c_bare, m->cmod, "bay_stretch_bare") == CUDA_SUCCESS &&
                  g_cu.ModuleGetFunction(&m->c_accu, m->cmod, "bay_stretch_accu") == CUDA_SUCCESS &&
                  g_cu.ModuleGetFunction(&m->c_logfn, m->cmod, "bay_logfn") == CUDA_SUCCESS &&
                  g_cu.ModuleGetFunction(&m->c_loop, m->cmod, "bay_stretch_loop") == CUDA_SUCCESS &&
                  g_cu.ModuleGetFunction(&m->c_accu_loop, m->cmod, "bay_stretch_accu_loop") == CUDA_SUCCESS &&
                  g_cu.ModuleGetGlobal(&m->cparams, &bytes, m->cmod, "bay_cparams") == CUDA_SUCCESS &&
                  bytes >= sizeof(float) * kCparamsCap;
        if (!ok) { g_cu.ModuleUnload(m->cmod); m->cmod = nullptr; return; }
        int per_sm = 0;
        if (g_cu.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->c_loop, m->loop_block, 0) == CUDA_SUCCESS)
            m->c_loop_capacity = per_sm * m->e->sm_count;
        if (g_cu.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m->c_accu_loop, m->e->wgs, 0) == CUDA_SUCCESS)
            m->c_accu_loop_capacity = per_sm * m->e->sm_count;
        m->cvar_state = 1;
    }
    s->cp = m->cvar_state == 1;
}

static CUfunction fn_bare(const bay_sampler* s) { return s->cp ? s->m->c_bare : s->m->f_bare; }
static CUfunction fn_accu(const bay_sampler* s) { return s->cp ? s->m->c_accu : s->m->f_accu; }
static CUfunction fn_logfn(const bay_sampler* s) { return s->cp ? s->m->c_logfn : s->m->f_logfn; }
static CUfunction fn_loop(const bay_sampler* s) { return s->cp ? s->m->c_loop : s->m->f_loop; }
static int loop_capacity(const bay_sampler* s) { return s->cp ? s->m->c_loop_capacity : s->m->loop_capacity; }
static CUfunction fn_accu_loop(const bay_sampler* s) { return s->cp ? s->m->c_accu_loop : s->m->f_accu_loop; }
static int accu_loop_capacity(const bay_sampler* s) {
    return s->cp ? s->m->c_accu_loop_capacity : s->m->accu_loop_capacity;
}

// The constant block belongs to the module, i.e. to all samplers of the model: (re)load it when another sampler's
// parameters are in it.  Stream-ordered, so kernels already queued keep the values they were launched with.
static int bind_params(bay_sampler* s) {
    bay_model* m = s->m;
    if (!s->cp || m->cparams_owner == s) return BAY_OK;
    CK(cudaMemcpyAsync(reinterpret_cast<void*>(m->cparams), s->params, sizeof(float) * (size_t)s->params_count,
                       cudaMemcpyDeviceToDevice, m->e->stream));
    m->cparams_owner = s;
    return BAY_OK;
}

// ----------------------------------------------------------------- sampler --
static int launch(bay_engine* e, CUfunction f, unsigned grid, unsigned block, void** args) {
    CUresult cr = g_cu.LaunchKernel(f, grid, 1, 1, block, 1, 1, 0, reinterpret_cast<CUstream>(e->stream), args, nullptr);
    if (cr != CUDA_SUCCESS) return cu_fail(cr, "cuLaunchKernel");
    g_launches++;
    return BAY_OK;
}

static void stretch_coeffs(float a, float* cA, float* cB, float* cC) {
    // z = A u^2 + B u + C (K/engines/nvidia-gtx-mcmc-stretch.cu:69-71), IEEE fp32, source order
    volatile float inv = 1.0f / a;
    volatile float t = a - 2.0f;
    *cA = t + inv;
    volatile float o = 1.0f - inv;
    *cB = 2.0f * o;
    *cC = inv;
}

static int aos_to_soa(bay_engine* e, const float* in, uint64_t offset, uint64_t ld, uint32_t dim, uint64_t n, float* out,
                      uint64_t pitch);

#include "engine_glm.inc"

// ---- multi-GPU mode A over peer memory ----------------------------------------------------------------------
// One block per sampler holds [xs D x W | lp W | mirror W x dima | 16 barrier flags]; its IPC handle is all-gathered
// over the engine's NCCL communicator (the only bootstrap channel the C ABI has) and every peer's block is mapped
// here.  Collective: all ranks create (and release) their partitioned samplers in the same order.
static int peer_block_alloc(bay_sampler* s) {
    bay_engine* e = s->m->e;
    const size_t W = (size_t)s->W, D = (size_t)s->D, dima = s->m->mirror ? (size_t)s->m->dima : 0;
    const size_t lp_off = D * W, xa_off = lp_off + W, flags_off = xa_off + W * dima;
    const size_t words = flags_off + 16;
    CK(cudaMalloc(&s->peer_block, sizeof(float) * words));
    CK(cudaMemsetAsync(s->peer_block, 0, sizeof(float) * words, e->stream));
    s->xs = s->peer_block;
    s->lp = s->peer_block + lp_off;
    s->xa = dima ? s->peer_block + xa_off : nullptr;
    s->peer_flags_off = flags_off;

    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, s->peer_block));
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    DevBuf dev_buf;
    CK(dev_buf.alloc(hb * e->nranks));
    char* dev = dev_buf.as<char>();
    std::vector<cudaIpcMemHandle_t> all((size_t)e->nranks);
    cudaError_t ce = cudaMemcpyAsync(dev + hb * e->rank, &mine, hb, cudaMemcpyHostToDevice, e->stream);
    ncclResult_t nr = ncclSuccess;
    if (ce == cudaSuccess) nr = g_nccl.AllGather(dev + hb * e->rank, dev, hb, ncclChar, e->comm, e->stream);
    if (ce == cudaSuccess && nr == ncclSuccess)
        ce = cudaMemcpyAsync(all.data(), dev, hb * e->nranks, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);   // also: every rank's block is zeroed by now
    if (nr != ncclSuccess) return fail(BAY_ENCCL, "handle exchange failed: %s", g_nccl.GetErrorString(nr));
    if (ce != cudaSuccess) return fail(BAY_ECUDA, "handle exchange failed: %s", cudaGetErrorString(ce));

    bay::PeerTable& t = s->peer_tab;
    memset(&t, 0, sizeof(t));
    t.n = (uint32_t)e->nranks;
    t.self = (uint32_t)e->rank;
    t.lp_off = lp_off;
    t.xa_off = xa_off;
    for (int r = 0; r < e->nranks; r++) {
        if (r == e->rank) { t.base[r] = (unsigned long long)(uintptr_t)s->peer_block; continue; }
        void* p = nullptr;
        ce = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess)
            return fail(BAY_ECUDA, "cannot map the ensemble of rank %d (no NVLink/P2P path?): %s — set BAY_P2P=0 for the "
                        "NCCL exchange", r, cudaGetErrorString(ce));
        s->peer_mapped[r] = p;
        t.base[r] = (unsigned long long)(uintptr_t)p;
    }
    return BAY_OK;
}

// tiny all-reduce used as a host-visible rendezvous of all ranks
static int comm_rendezvous(bay_engine* e) {
    DevBuf flag_buf;
    CK(flag_buf.alloc(sizeof(int)));
    int* flag = flag_buf.as<int>();
    cudaError_t ce = cudaMemsetAsync(flag, 0, sizeof(int), e->stream);
    ncclResult_t nr = ncclSuccess;
    if (ce == cudaSuccess) nr = g_nccl.AllReduce(flag, flag, 1, ncclInt, ncclSum, e->comm, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (nr != ncclSuccess) return fail(BAY_ENCCL, "rendezvous failed: %s", g_nccl.GetErrorString(nr));
    if (ce != cudaSuccess) return fail(BAY_ECUDA, "rendezvous failed: %s", cudaGetErrorString(ce));
    return BAY_OK;
}

static void peer_block_release(bay_sampler* s) {
    if (!s->peer_block) return;
    bay_engine* e = s->m->e;
    // every half-step ended with the flag barrier, so no peer store is in flight once the streams are drained
    for (int r = 0; r < 8; r++)
        if (s->peer_mapped[r]) cudaIpcCloseMemHandle(s->peer_mapped[r]);
    comm_rendezvous(e);                       // every mapping of this block is closed before it is freed
    cudaFree(s->peer_block);
    s->peer_block = nullptr;
    s->xs = s->lp = s->xa = nullptr;
}

static int sampler_alloc(bay_sampler* s) {
    const bay_engine* e = s->m->e;
    const size_t W = (size_t)s->W, D = (size_t)s->D, wgs = (size_t)e->wgs;
    if (s->m->peers) {
        TRY(peer_block_alloc(s));
    } else {
        CK(cudaMalloc(&s->xs, sizeof(float) * D * W));
        CK(cudaMalloc(&s->lp, sizeof(float) * W));
        if (s->m->mirror) CK(cudaMalloc(&s->xa, sizeof(float) * W * (size_t)s->m->dima));
    }
    if (s->xa) CK(cudaMemsetAsync(s->xa, 0, sizeof(float) * W * (size_t)s->m->dima, e->stream));
    CK(cudaMalloc(&s->loop_bar, sizeof(unsigned int) * 2));
    CK(cudaMemsetAsync(s->loop_bar, 0, sizeof(unsigned int) * 2, e->stream));
    CK(cudaMalloc(&s->accept, sizeof(uint32_t) * s->G));
    CK(cudaMalloc(&s->blk_sums, sizeof(float) * 2 * D * s->G));   // second half: scratch of the run-sampler! loop
    CK(cudaMalloc(&s->accept_total, sizeof(unsigned long long)));
    CK(cudaMalloc(&s->hist_counts, sizeof(uint32_t) * wgs * D));
    CK(cudaMalloc(&s->mm, sizeof(uint32_t) * 2 * D));
    CK(cudaMalloc(&s->limits, sizeof(float) * 2 * D));
    CK(cudaMalloc(&s->pdf, sizeof(float) * wgs * D));
    CK(cudaMalloc(&s->ranks, sizeof(float) * wgs * D));
    CK(cudaMalloc(&s->macc, sizeof(double) * 2 * D));
    CK(cudaMalloc(&s->vec_d, sizeof(float) * 4 * D));
    CK(cudaMemsetAsync(s->xs, 0, sizeof(float) * D * W, e->stream));
    CK(cudaMemsetAsync(s->lp, 0, sizeof(float) * W, e->stream));
    CK(cudaMemsetAsync(s->accept, 0, sizeof(uint32_t) * s->G, e->stream));
    CK(cudaMemsetAsync(s->blk_sums, 0, sizeof(float) * 2 * D * s->G, e->stream));
    CK(cudaMemsetAsync(s->hist_counts, 0, sizeof(uint32_t) * wgs * D, e->stream));
    return BAY_OK;
}

static int sampler_create_common(bay_model* m, int32_t seed, int64_t walkers, int64_t params_count,
                                 bay_sampler** out) {
    if (!m || !out) return fail(BAY_EINVAL, "NULL argument");
    const int wgs = m->e->wgs;
    // G/:552, 609-610
    if (walkers < 2 * wgs || walkers % (2 * wgs) != 0 || walkers > (int64_t)1 << 31)
        return fail(BAY_EINVAL_WALKERS, "Number of walkers (%lld) must be a multiple of %d.", (long long)walkers, 2 * wgs);
    if (params_count < 0) return fail(BAY_EINVAL, "negative params_count");
    // element indices of the ensemble are 32-bit in the kernels (walker index, DIM x W / 4 init counters)
    if ((uint64_t)walkers * (uint64_t)m->dim >= ((uint64_t)1 << 32))
        return fail(BAY_EINVAL, "ensemble of %lld walkers x %d dimensions has 2^32 or more elements", (long long)walkers, m->dim);
    if (m->e->comm && m->e->nranks > 1 && !m->glm && walkers % ((int64_t)2 * wgs * m->e->nranks) != 0)
        return fail(BAY_EINVAL_WALKERS, "Number of walkers (%lld) must be a multiple of %d.", (long long)walkers,
                    2 * wgs * m->e->nranks);
    USE_ENGINE(m->e);
    bay_sampler* s = new bay_sampler();
    s->m = m;
    s->W = walkers;
    s->H = walkers / 2;
    s->D = m->dim;
    s->G = cdiv(s->H, wgs);
    s->params_len = (uint32_t)m->params_size;
    // G/:559-560: data-len = max(0, entries(params) - params-size)
    s->data_len = (uint32_t)(params_count > m->params_size ? params_count - m->params_size : 0);
    s->params_count = params_count;
    int r = sampler_alloc(s);
    if (r != BAY_OK) { bay_sampler_release(s); return r; }
    bay_init(s, seed);
    *out = s;
    return BAY_OK;
}

extern "C" int bay_sampler_create(bay_model* m, int32_t seed, int64_t walkers, const float* params_host,
                                  int64_t params_count, bay_sampler** out) {
    if (params_count > 0 && !params_host) return fail(BAY_EINVAL, "params_host is NULL");
    if (!m) return fail(BAY_EINVAL, "NULL model");
    USE_ENGINE(m->e);
    bay_sampler* s = nullptr;
    TRY(sampler_create_common(m, seed, walkers, params_count, &s));
    const size_t bytes = sizeof(float) * (size_t)(params_count > 0 ? params_count : 1);
    cudaError_t ce = cudaMalloc(&s->params, bytes);
    if (ce == cudaSuccess && params_count > 0)
        ce = cudaMemcpyAsync(s->params, params_host, sizeof(float) * params_count, cudaMemcpyHostToDevice, m->e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(m->e->stream);
    if (ce != cudaSuccess) {
        bay_sampler_release(s);
        return fail(BAY_ECUDA, "params upload failed: %s", cudaGetErrorString(ce));
    }
    s->own_params = true;
    cparams_try(s, params_count);
    if (m->glm) {
        int r = glm_setup(s);
        if (r != BAY_OK) { bay_sampler_release(s); return r; }
    }
    *out = s;
    return BAY_OK;
}

extern "C" int bay_sampler_create_dev(bay_model* m, int32_t seed, int64_t walkers, uint64_t params_dev,
                                      int64_t params_count, bay_sampler** out) {
    if (!m) return fail(BAY_EINVAL, "NULL model");
    USE_ENGINE(m->e);
    bay_sampler* s = nullptr;
    TRY(sampler_create_common(m, seed, walkers, params_count, &s));
    s->params = reinterpret_cast<float*>(params_dev);
    s->own_params = false;
    if (m->glm) {
        int r = glm_setup(s);
        if (r != BAY_OK) { bay_sampler_release(s); return r; }
    }
    *out = s;
    return BAY_OK;
}

extern "C" int bay_sampler_release(bay_sampler* s) {
    if (!s) return BAY_OK;
    CtxScope scope(s->m->e);
    cudaStreamSynchronize(s->m->e->stream);
    if (s->own_params) cudaFree(s->params);
    if (s->m->cparams_owner == s) s->m->cparams_owner = nullptr;
    glm_release(s);
    peer_block_release(s);   // clears xs / lp / xa when they live in the shared block
    void* bufs[] = {s->xs, s->lp, s->accept, s->blk_sums, s->accept_total, s->means, s->hist_counts, s->mm,
                    s->limits, s->pdf, s->ranks, s->macc, s->vec_d, s->stage, s->xa, s->loop_bar, s->loop_betas, s->accept_all,
                    s->sample_stage};
    for (void* b : bufs) if (b) cudaFree(b);
    delete s;
    return BAY_OK;
}

// init! G/:391-400
extern "C" int bay_init(bay_sampler* s, int32_t seed) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    s->bare_seed = seed;
    s->a_bare = 2.0f;
    s->beta = 1.0f;
    s->bare_counter = 0;
    s->move_seed = seed;
    return BAY_OK;
}

// rebuild the AoS mirror from the SoA state (after init / state hand-off)
static int mirror_sync(bay_sampler* s) {
    if (!s->m->mirror) return BAY_OK;
    bay_engine* e = s->m->e;
    dim3 grid(cdiv((uint64_t)s->W, 32), cdiv((uint64_t)s->D, 32)), block(32, 8);
    bay::k_soa_to_aos<<<grid, block, 0, e->stream>>>(s->xs, (uint32_t)s->W, (uint32_t)s->D, (uint64_t)s->W, s->xa,
                                                     (uint32_t)s->m->dima);
    CKLAUNCH();
    return BAY_OK;
}

static int peer_settle(bay_sampler* s);
static int soa_fresh(bay_sampler* s);
static int aos_to_soa(bay_engine* e, const float* in, uint64_t offset, uint64_t ld, uint32_t dim, uint64_t n, float* out,
                      uint64_t pitch);

static int launch_logfn_all(bay_sampler* s) {
    bay_model* m = s->m;
    if (m->glm) return glm_logfn_all(s);
    uint32_t n = (uint32_t)s->W, pitch = (uint32_t)s->W;
    void* args[] = {&n, &s->data_len, &s->params_len, &s->params, &s->xs, &pitch, &s->lp};
    TRY(bind_params(s));
    return launch(m->e, fn_logfn(s), cdiv(n, m->block), m->block, args);
}

// init-position! [seed limits] G/:409-418
extern "C" int bay_init_position_uniform(bay_sampler* s, int32_t seed, const float* limits_host) {
    if (!s || !limits_host) return fail(BAY_EINVAL, "NULL argument");
    bay_engine* e = s->m->e;
    USE_ENGINE(e);
    CK(cudaMemcpyAsync(s->limits, limits_host, sizeof(float) * 2 * s->D, cudaMemcpyHostToDevice, e->stream));
    const uint32_t n4 = (uint32_t)((uint64_t)s->D * s->W / 4);
    bay::k_init_walkers<<<cdiv(n4, 256), 256, 0, e->stream>>>(n4, (uint32_t)s->D, (uint32_t)seed, s->limits, s->xs,
                                                           (uint32_t)s->W);
    CKLAUNCH();
    s->soa_stale = s->soa_own_stale = s->remote_stale = false;   // a full rewrite: every local copy is current
    s->version++;
    TRY(mirror_sync(s));
    TRY(launch_logfn_all(s));
    TRY(peer_settle(s));
    // limits_host may be pageable: the async copy above is staged before return, but be safe
    CK(cudaStreamSynchronize(e->stream));
    s->iterations = 0;
    return BAY_OK;
}

// init-position! [position] G/:401-408
extern "C" int bay_init_position_from(bay_sampler* s, const bay_sampler* other) {
    if (!s || !other) return fail(BAY_EINVAL, "NULL argument");
    if (s->W != other->W || s->D != other->D) return fail(BAY_EINVAL, "samplers differ in shape");
    bay_engine* e = s->m->e;
    USE_ENGINE(e);
    TRY(soa_fresh(const_cast<bay_sampler*>(other)));
    CK(cudaMemcpyAsync(s->xs, other->xs, sizeof(float) * (size_t)s->D * s->W, cudaMemcpyDeviceToDevice, e->stream));
    TRY(peer_settle(const_cast<bay_sampler*>(other)));
    s->soa_stale = s->soa_own_stale = s->remote_stale = false;
    s->version++;
    TRY(mirror_sync(s));
    TRY(launch_logfn_all(s));
    TRY(peer_settle(s));
    s->iterations = 0;
    return BAY_OK;
}

// ---- multi-GPU mode A: walker partition ------------------------------------------------------
// With a communicator, a non-GLM sampler's ensemble is replicated on every rank but rank r only UPDATES the walkers
// [r*H/R, (r+1)*H/R) of the active half; the updated slice (SoA rows, log-densities, mirror rows) is then
// all-gathered so that every rank holds the full half for the next half-step's partner gather.  Philox counters
// are global walker indices and every walker is computed by the same code from the same inputs, so an R-GPU chain
// is bit-identical to the 1-GPU chain.
static bool partitioned(const bay_sampler* s) { return s->m->e->comm != nullptr && s->m->e->nranks > 1 && !s->m->glm; }

static void my_slice(const bay_sampler* s, uint32_t* k_begin, uint32_t* k_end) {
    if (!partitioned(s)) { *k_begin = 0; *k_end = (uint32_t)s->H; return; }
    const uint32_t hs = (uint32_t)(s->H / s->m->e->nranks);
    *k_begin = hs * (uint32_t)s->m->e->rank;
    *k_end = *k_begin + hs;
}

static int exchange_half(bay_sampler* s, int half) {
    if (!partitioned(s)) return BAY_OK;
    bay_engine* e = s->m->e;
    if (s->m->peers) {
        // the kernel already stored the accepted walkers into every rank's block: only a barrier is left, so that
        // no rank starts the next half-step before all stores into its block have landed
        s->peer_epoch++;
        bay::k_peer_barrier<<<1, 32, 0, e->stream>>>(s->peer_tab, s->peer_flags_off, (uint32_t)e->rank, s->peer_epoch);
        g_launches++;
        CK(cudaGetLastError());
        return BAY_OK;
    }
    const size_t hs = (size_t)(s->H / e->nranks), H = (size_t)s->H, W = (size_t)s->W, r = (size_t)e->rank;
    const size_t h0 = half ? H : 0;
    CKNCCL(g_nccl.GroupStart());
    for (int d = 0; d < s->D; d++) {
        float* row = s->xs + (size_t)d * W + h0;
        CKNCCL(g_nccl.AllGather(row + r * hs, row, hs, ncclFloat, e->comm, e->stream));
    }
    CKNCCL(g_nccl.AllGather(s->lp + h0 + r * hs, s->lp + h0, hs, ncclFloat, e->comm, e->stream));
    if (s->xa) {
        float* base = s->xa + h0 * s->m->dima;
        CKNCCL(g_nccl.AllGather(base + r * hs * s->m->dima, base, hs * s->m->dima, ncclFloat, e->comm, e->stream));
    }
    CKNCCL(g_nccl.GroupEnd());
    return BAY_OK;
}

// The half-step barrier orders "my stores have landed" before "you go on"; it does not stop a fast rank from storing
// step k+1 walkers into a slow rank that is still READING the whole ensemble (sample!, histogram!, mean, state
// hand-off) or REWRITING it (init-position!).  Every such whole-ensemble access therefore ends with one more flag
// barrier: peers pass it only after this rank's access has completed.  (Collective: all ranks make the same calls.)
static int peer_settle(bay_sampler* s) {
    if (!partitioned(s) || !s->m->peers) return BAY_OK;
    return exchange_half(s, 0);
}

// Peers forward only mirror rows (stretch_program.inc, bay_forward_peers): before anything reads the SoA matrix
// outside a rank's own slice (sample!, histogram!, mean, state hand-off) it is rebuilt from the mirror.
// Pull mode: a rank only maintains its own slices; before a read-out it copies every other rank's slices (SoA rows,
// log-densities, mirror rows) out of that rank's block.  The half-step barrier has already made them final.
static int ensemble_gather(bay_sampler* s) {
    if (!s->remote_stale) return BAY_OK;
    bay_engine* e = s->m->e;
    const size_t W = (size_t)s->W, H = (size_t)s->H, D = (size_t)s->D, hs = H / e->nranks;
    const size_t dima = s->xa ? (size_t)s->m->dima : 0;
    for (int r = 0; r < e->nranks; r++) {
        if (r == e->rank) continue;
        const float* blk = reinterpret_cast<const float*>((uintptr_t)s->peer_tab.base[r]);
        for (size_t half = 0; half < 2; half++) {
            const size_t c0 = half * H + (size_t)r * hs;
            CK(cudaMemcpy2DAsync(s->xs + c0, W * sizeof(float), blk + c0, W * sizeof(float), hs * sizeof(float), D,
                                 cudaMemcpyDeviceToDevice, e->stream));
            CK(cudaMemcpyAsync(s->lp + c0, blk + s->peer_tab.lp_off + c0, hs * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
            if (dima)
                CK(cudaMemcpyAsync(s->xa + c0 * dima, blk + s->peer_tab.xa_off + c0 * dima, hs * dima * sizeof(float),
                                   cudaMemcpyDeviceToDevice, e->stream));
        }
    }
    s->remote_stale = false;
    return BAY_OK;
}

static int soa_fresh(bay_sampler* s) {
    TRY(ensemble_gather(s));
    if (!s->soa_stale && !s->soa_own_stale) return BAY_OK;
    TRY(aos_to_soa(s->m->e, s->xa, 0, (uint64_t)s->m->dima, (uint32_t)s->D, (uint64_t)s->W, s->xs, (uint64_t)s->W));
    s->soa_stale = false;
    s->soa_own_stale = false;
    return BAY_OK;
}

// Before a GENERIC kernel (which reads its own walkers from the SoA matrix and, between GPUs, expects every replica
// to be current) runs after tensor-core moves: bring in the other ranks' slices, rebuild the SoA matrix, and — peers
// may go on to overwrite their slices — pass a barrier once everybody has read what it needs.  Collective.
static int generic_ready(bay_sampler* s) {
    if (!s->soa_own_stale && !s->remote_stale) return BAY_OK;
    const bool gathered = s->remote_stale;
    TRY(soa_fresh(s));
    if (gathered) TRY(peer_settle(s));
    return BAY_OK;
}

// ---- BAY_MODEL_QUADFORM: half-ensemble move on the tensor cores (kernels_quadform_tc.cuh) ----------------------
// Eligible: the model's parameters are [mu | U] in device memory and, between GPUs, the ensemble lives in the
// peer-mapped block (partner rows are pulled from the owning rank over NVLink; nothing is forwarded).
static bool quadform_usable(const bay_sampler* s) {
    const bay_model* m = s->m;
    if (!m->quadform || !s->xa || s->params_count < (int64_t)s->D + (int64_t)s->D * s->D) return false;
    if (partitioned(s) && !m->peers) return false;    // NCCL all-gather exchange: generic kernels only
    return true;
}

static int quadform_half(bay_sampler* s, int half, uint32_t seed, uint32_t tag, float cA, float cB, float cC,
                         float beta, uint32_t step) {
    bay_model* m = s->m;
    bay_engine* e = m->e;
    const bool bulk = !partitioned(s);
    const size_t smem = bay::qf::smem_bytes(bulk);
    if (!s->qf_configured) {
        CK(cudaFuncSetAttribute(bay::qf::k_quadform_move_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)bay::qf::smem_bytes(true)));
        CK(cudaFuncSetAttribute(bay::qf::k_quadform_move_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)bay::qf::smem_bytes(false)));
        s->qf_configured = true;
    }
    bay::qf::Args a;
    memset(&a, 0, sizeof a);
    const size_t dima = (size_t)m->dima, H = (size_t)s->H;
    a.mu = s->params;
    a.U = s->params + s->D;
    a.xa_active = s->xa + (half ? H : 0) * dima;
    a.lp_active = s->lp + (half ? H : 0);
    const size_t compl_off = (half ? 0 : H) * dima;
    if (partitioned(s)) {
        for (int r = 0; r < e->nranks; r++)
            a.xa_compl[r] = s->peer_tab.base[r] + sizeof(float) * (s->peer_tab.xa_off + compl_off);
        a.hs = (uint32_t)(s->H / e->nranks);
        s->remote_stale = true;
    } else {
        a.xa_compl[0] = (unsigned long long)(uintptr_t)(s->xa + compl_off);
        a.hs = (uint32_t)s->H;
    }
    a.K = (uint32_t)s->H;
    my_slice(s, &a.k_begin, &a.k_end);
    a.D = (uint32_t)s->D;
    a.DA = (uint32_t)m->dima;
    a.seed = seed; a.tag = tag; a.step = step;
    a.cA = cA; a.cB = cB; a.cC = cC; a.beta = beta;
    a.accepted = nullptr;
    const uint32_t tiles = cdiv(a.k_end - a.k_begin, bay::qf::TILE);
    uint32_t grid = (uint32_t)e->sm_count < tiles ? (uint32_t)e->sm_count : tiles;
    if (!bulk) bay::qf::k_quadform_move_tc<false><<<grid, bay::qf::THREADS, smem, e->stream>>>(a);
    else bay::qf::k_quadform_move_tc<true><<<grid, bay::qf::THREADS, smem, e->stream>>>(a);
    CKLAUNCH();
    s->soa_own_stale = true;
    return exchange_half(s, half);
}

static int half_bare(bay_sampler* s, int half, uint32_t seed, uint32_t tag, float cA, float cB, float cC,
                     float beta, uint32_t step) {
    bay_model* m = s->m;
    s->version++;
    if (m->glm) return glm_half(s, half, seed, tag, cA, cB, cC, beta, step, 0u);
    if (quadform_usable(s)) return quadform_half(s, half, seed, tag, cA, cB, cC, beta, step);
    TRY(generic_ready(s));
    uint32_t K = (uint32_t)s->H, pitch = (uint32_t)s->W;
    float* act = s->xs + (half ? s->H : 0);
    float* cmp = s->xs + (half ? 0 : s->H);
    float* lp = s->lp + (half ? s->H : 0);
    float* cmp_a = s->xa ? s->xa + (size_t)(half ? 0 : s->H) * m->dima : nullptr;
    float* act_a = s->xa ? s->xa + (size_t)(half ? s->H : 0) * m->dima : nullptr;
    uint32_t kb, ke;
    my_slice(s, &kb, &ke);
    uint32_t col0 = half ? K : 0u;
    std::vector<void*> args = {&K, &seed, &tag, &s->data_len, &s->params_len, &s->params, &cmp, &act, &pitch, &lp,
                               &cA, &cB, &cC, &beta, &step, &kb, &ke};
    if (m->mirror) { args.push_back(&cmp_a); args.push_back(&act_a); }
    uint32_t pull_hs = (uint32_t)(s->H / m->e->nranks), ccol0 = half ? 0u : K;
    if (m->peers) {
        args.push_back(&s->peer_tab);
        args.push_back(&col0);
        if (m->pull) { args.push_back(&pull_hs); args.push_back(&ccol0); s->remote_stale = true; }
        else s->soa_stale = m->mirror;
    }
    TRY(bind_params(s));
    TRY(launch(m->e, fn_bare(s), cdiv(ke - kb, m->block), m->block, args.data()));
    return exchange_half(s, half);
}

static int half_accu(bay_sampler* s, int half, uint32_t seed, uint32_t tag, float cA, float cB, float cC,
                     uint32_t step) {
    bay_model* m = s->m;
    s->version++;
    if (m->glm) return glm_half(s, half, seed, tag, cA, cB, cC, 1.0f, step, 1u);
    TRY(generic_ready(s));
    uint32_t K = (uint32_t)s->H, pitch = (uint32_t)s->W, accumulate = half ? 1u : 0u;
    float* act = s->xs + (half ? s->H : 0);
    float* cmp = s->xs + (half ? 0 : s->H);
    float* lp = s->lp + (half ? s->H : 0);
    float* cmp_a = s->xa ? s->xa + (size_t)(half ? 0 : s->H) * m->dima : nullptr;
    float* act_a = s->xa ? s->xa + (size_t)(half ? s->H : 0) * m->dima : nullptr;
    uint32_t kb, ke;
    my_slice(s, &kb, &ke);
    uint32_t col0 = half ? K : 0u;
    std::vector<void*> args = {&K, &seed, &tag, &s->data_len, &s->params_len, &s->params, &cmp, &act, &pitch, &lp,
                               &s->accept, &s->blk_sums, &cA, &cB, &cC, &step, &accumulate, &kb, &ke};
    if (m->mirror) { args.push_back(&cmp_a); args.push_back(&act_a); }
    uint32_t pull_hs = (uint32_t)(s->H / m->e->nranks), ccol0 = half ? 0u : K;
    if (m->peers) {
        args.push_back(&s->peer_tab);
        args.push_back(&col0);
        if (m->pull) { args.push_back(&pull_hs); args.push_back(&ccol0); s->remote_stale = true; }
        else s->soa_stale = m->mirror;
    }
    TRY(bind_params(s));
    TRY(launch(m->e, fn_accu(s), cdiv(ke - kb, m->e->wgs), m->e->wgs, args.data()));
    return exchange_half(s, half);
}

// Persistent step loop (bay_stretch_loop): n moves in one cooperative launch when the whole half-ensemble is
// co-resident — the launch-latency-bound regime of small ensembles.  BAY_LOOP=0 disables it.
static bool loop_usable(const bay_sampler* s, int64_t n) {
    const bay_model* m = s->m;
    if (m->glm || !m->f_loop || n < 2) return false;
    if (quadform_usable(s)) return false;             // the tensor-core move runs one launch per half-step
    if (partitioned(s) && (!m->peers || m->pull)) return false;   // NCCL exchange / pull mode: one kernel per half-step
    uint32_t kb, ke;
    my_slice(s, &kb, &ke);
    if ((int64_t)cdiv(ke - kb, m->loop_block) > loop_capacity(s)) return false;
    const char* env = getenv("BAY_LOOP");
    return !(env && env[0] == '0');
}

static int move_bare_loop(bay_sampler* s, int64_t n, const float* betas, float cA, float cB, float cC) {
    bay_model* m = s->m;
    bay_engine* e = m->e;
    s->version++;
    TRY(generic_ready(s));
    float* betas_dev = nullptr;
    if (betas) {
        if (s->loop_betas_cap < n) {
            if (s->loop_betas) { CK(cudaStreamSynchronize(e->stream)); CK(cudaFree(s->loop_betas)); s->loop_betas = nullptr; }
            CK(cudaMalloc(&s->loop_betas, sizeof(float) * n));
            s->loop_betas_cap = n;
        }
        CK(cudaMemcpyAsync(s->loop_betas, betas, sizeof(float) * n, cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));   // betas is a caller-owned host vector
        betas_dev = s->loop_betas;
    }
    int64_t done = 0;
    while (done < n) {
        const int64_t chunk = (n - done) < (1 << 20) ? (n - done) : (1 << 20);
        uint32_t K = (uint32_t)s->H, pitch = (uint32_t)s->W, seed = (uint32_t)s->bare_seed, step0 = s->bare_counter;
        uint32_t n_steps = (uint32_t)chunk;
        float beta_const = s->beta;
        const float* bptr = betas_dev ? betas_dev + done : nullptr;
        uint32_t kb, ke, epoch0 = s->peer_epoch;
        my_slice(s, &kb, &ke);
        std::vector<void*> args = {&K, &seed, &s->data_len, &s->params_len, &s->params, &s->xs, &pitch, &s->lp, &cA, &cB,
                                   &cC, &bptr, &beta_const, &step0, &n_steps, &s->loop_bar, &kb, &ke};
        if (m->mirror) args.push_back(&s->xa);
        if (m->peers) {
            // the loop's grid barrier also spans the GPUs: one flag epoch per half-step
            args.push_back(&s->peer_tab);
            args.push_back(&s->peer_flags_off);
            args.push_back(&epoch0);
            s->peer_epoch += 2u * n_steps;
            s->soa_stale = m->mirror;
        }
        TRY(bind_params(s));
        CUresult cr = g_cu.LaunchCooperativeKernel(fn_loop(s), cdiv(ke - kb, m->loop_block), 1, 1, m->loop_block, 1, 1, 0,
                                                   reinterpret_cast<CUstream>(e->stream), args.data());
        if (cr != CUDA_SUCCESS) return cu_fail(cr, "cuLaunchCooperativeKernel(bay_stretch_loop)");
        g_launches++;
        s->bare_counter += (uint32_t)chunk;
        done += chunk;
    }
    return BAY_OK;
}

// move-bare! G/:358-364: odd = (X=s0, S=s1, seed, tag 3333), even = (X=s1, S=s0, seed+1, tag 4444)
static int move_bare_n(bay_sampler* s, int64_t n, const float* betas /* nullable: use s->beta */) {
    float cA, cB, cC;
    stretch_coeffs(s->a_bare, &cA, &cB, &cC);
    if (loop_usable(s, n)) return move_bare_loop(s, n, betas, cA, cB, cC);
    for (int64_t i = 0; i < n; i++) {
        const float beta = betas ? betas[i] : s->beta;
        TRY(half_bare(s, 0, (uint32_t)s->bare_seed, 3333u, cA, cB, cC, beta, s->bare_counter));
        TRY(half_bare(s, 1, (uint32_t)(s->bare_seed + 1), 4444u, cA, cB, cC, beta, s->bare_counter));
        s->bare_counter++;
    }
    return BAY_OK;
}

extern "C" int bay_move_bare(bay_sampler* s) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    USE_ENGINE(s->m->e);
    return move_bare_n(s, 1, nullptr);
}

// One launch of stretch_move_bare at the current counter (no counter increment): the
// raw-launch granularity of the reference's kernel-level test
// (T/internal/nvidia_gtx_test.clj:217-235).  half 0 = odd launch, 1 = even launch.
extern "C" int bay_move_bare_half(bay_sampler* s, int half) {
    if (!s || (half != 0 && half != 1)) return fail(BAY_EINVAL, "bad argument");
    USE_ENGINE(s->m->e);
    float cA, cB, cC;
    stretch_coeffs(s->a_bare, &cA, &cB, &cC);
    return half_bare(s, half, (uint32_t)(s->bare_seed + half), half ? 4444u : 3333u, cA, cB, cC, s->beta,
                     s->bare_counter);
}

extern "C" int bay_set_a(bay_sampler* s, float a) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    s->a_bare = a;
    return BAY_OK;
}

extern "C" int bay_set_temperature(bay_sampler* s, float t) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    s->beta = (float)(1.0 / (double)t);  // G/:366: (cast-prim (/ 1.0 t))
    return BAY_OK;
}

// burn-in! G/:419-429
extern "C" int bay_burn_in(bay_sampler* s, int64_t n, float a) {
    if (!s || n < 0) return fail(BAY_EINVAL, "bad argument");
    USE_ENGINE(s->m->e);
    s->a_bare = a;
    s->beta = 1.0f;
    TRY(move_bare_n(s, n, nullptr));
    s->iterations += n;
    return BAY_OK;
}

// anneal! G/:430-440
extern "C" int bay_anneal(bay_sampler* s, const float* temperature_host, int64_t n, float a) {
    if (!s || n < 0 || (n > 0 && !temperature_host)) return fail(BAY_EINVAL, "bad argument");
    USE_ENGINE(s->m->e);
    s->a_bare = a;
    std::vector<float> betas((size_t)n);
    for (int64_t i = 0; i < n; i++) betas[i] = (float)(1.0 / (double)temperature_host[i]);
    TRY(move_bare_n(s, n, betas.data()));
    if (n > 0) s->beta = betas[n - 1];
    s->iterations += n;
    return BAY_OK;
}

static int ensure_means(bay_sampler* s, int64_t n) {
    if (s->means_cap >= n) return BAY_OK;
    int64_t cap = s->means_cap * 2 > n ? s->means_cap * 2 : n;
    float* fresh = nullptr;
    CK(cudaMalloc(&fresh, sizeof(float) * (size_t)s->D * cap));
    if (s->means && s->means_n > 0)
        CK(cudaMemcpyAsync(fresh, s->means, sizeof(float) * (size_t)s->D * s->means_n, cudaMemcpyDeviceToDevice,
                           s->m->e->stream));
    if (s->means) {
        CK(cudaStreamSynchronize(s->m->e->stream));
        CK(cudaFree(s->means));
    }
    s->means = fresh;
    s->means_cap = cap;
    return BAY_OK;
}

// init-move! G/:340-350 (accept is zero-filled: SURVEY Appendix B-3)
extern "C" int bay_init_move(bay_sampler* s, float a) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    bay_engine* e = s->m->e;
    USE_ENGINE(e);
    s->move_seed += 2;
    s->move_counter = 0;
    s->a_move = a;
    s->means_n = 0;
    CK(cudaMemsetAsync(s->accept, 0, sizeof(uint32_t) * s->G, e->stream));
    CK(cudaMemsetAsync(s->blk_sums, 0, sizeof(float) * (size_t)s->D * s->G, e->stream));
    return BAY_OK;
}

// move! G/:351-357: odd = (X=s0, S=s1, move-seed, 1111), even = (X=s1, S=s0, move-seed+1, 2222)
extern "C" int bay_move(bay_sampler* s) {
    if (!s) return fail(BAY_EINVAL, "NULL sampler");
    bay_engine* e = s->m->e;
    USE_ENGINE(e);
    float cA, cB, cC;
    stretch_coeffs(s->a_move, &cA, &cB, &cC);
    TRY(ensure_means(s, s->means_n + 1));
    if (partitioned(s))   // blocks of other ranks must contribute exact zeros to the all-reduce below
        CK(cudaMemsetAsync(s->blk_sums, 0, sizeof(float) * (size_t)s->D * s->G, e->stream));
    TRY(half_accu(s, 0, (uint32_t)s->move_seed, 1111u, cA, cB, cC, s->move_counter));
    TRY(half_accu(s, 1, (uint32_t)(s->move_seed + 1), 2222u, cA, cB, cC, s->move_counter));
    if (partitioned(s))   // every block sum is produced by exactly one rank: x + 0 + ... + 0 is exact
        CKNCCL(g_nccl.AllReduce(s->blk_sums, s->blk_sums, (size_t)s->D * s->G, ncclFloat, ncclSum, e->comm, e->stream));
    const float factor = 0.5f / ((float)e->wgs * (float)s->G);
    bay::k_step_means<<<cdiv((uint64_t)s->D * 32, 128), 128, 0, e->stream>>>(
        (uint32_t)s->D, s->G, s->blk_sums, factor, s->means + (size_t)s->means_n * s->D);
    CKLAUNCH();
    s->means_n++;
    s->move_counter++;
    return BAY_OK;
}

// n x move! in one cooperative launch (bay_stretch_accu_loop) when the half-ensemble's blocks are co-resident.
static bool accu_loop_usable(const bay_sampler* s, int64_t n) {
    const bay_model* m = s->m;
    if (m->glm || !fn_accu_loop(s) || n < 2 || partitioned(s)) return false;
    if ((int64_t)s->G > accu_loop_capacity(s)) return false;
    const char* env = getenv("BAY_LOOP");
    return !(env && env[0] == '0');
}

static int move_accu_loop(bay_sampler* s, int64_t n) {
    bay_model* m = s->m;
    bay_engine* e = m->e;
    float cA, cB, cC;
    stretch_coeffs(s->a_move, &cA, &cB, &cC);
    TRY(ensure_means(s, s->means_n + n));
    s->version++;
    TRY(generic_ready(s));
    TRY(bind_params(s));
    uint32_t K = (uint32_t)s->H, pitch = (uint32_t)s->W, seed = (uint32_t)s->move_seed, step0 = s->move_counter;
    uint32_t n_steps = (uint32_t)n;
    float factor = 0.5f / ((float)e->wgs * (float)s->G);
    float* means = s->means + (size_t)s->means_n * s->D;
    std::vector<void*> args = {&K, &seed, &s->data_len, &s->params_len, &s->params, &s->xs, &pitch, &s->lp, &s->accept,
                               &s->blk_sums, &cA, &cB, &cC, &step0, &n_steps, &s->loop_bar, &means, &factor};
    if (m->mirror) args.push_back(&s->xa);
    CUresult cr = g_cu.LaunchCooperativeKernel(fn_accu_loop(s), s->G, 1, 1, (unsigned)e->wgs, 1, 1, 0,
                                               reinterpret_cast<CUstream>(e->stream), args.data());
    if (cr != CUDA_SUCCESS) return cu_fail(cr, "cuLaunchCooperativeKernel(bay_stretch_accu_loop)");
    g_launches++;
    s->means_n += n;
    s->move_counter += (uint32_t)n;
    return BAY_OK;
}

#include "engine_estimate.inc"

#include "kernels_rng.cuh"
#include "engine_next.inc"
