"""The drop-in boundary inside a NON-primary CUDA context (SURVEY §8b "CUDA context / buffer interop").

The reference never uses the primary context: ClojureCUDA's ``with-default`` creates a driver-API context
(/root/reference/src/clojure/uncomplicate/bayadera/cuda.clj:27-31), every engine call is wrapped in
``(in-context ctx ...)`` (internal/device/nvidia_gtx.clj:374, 402, ...) and parameters / results are raw CUdeviceptr
of that context (``cuda-float`` vectors, nvidia_gtx.clj:558, 798).  Here the test plays ClojureCUDA through ctypes on
libcuda: cuCtxCreate, cuMemAlloc'd params and result buffers, and the engine adopting the current context.
"""
import ctypes as C

import numpy as np
import pytest

import goldens as G
import bayadera_b200 as bb
from bayadera_b200 import models
from bayadera_b200.engine import DeviceParams
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def f32(x):
    return np.asarray(x, dtype=np.float32)


class Driver:
    """The handful of driver-API calls ClojureCUDA makes around the engine."""

    def __init__(self):
        self.cu = C.CDLL("libcuda.so.1")
        self.ok(self.cu.cuInit(0))

    @staticmethod
    def ok(rc):
        assert rc == 0, f"CUDA driver error {rc}"

    def ctx_create(self, ordinal=0):
        dev, ctx = C.c_int(), C.c_void_p()
        self.ok(self.cu.cuDeviceGet(C.byref(dev), ordinal))
        self.ok(self.cu.cuCtxCreate_v2(C.byref(ctx), 0, dev))     # created AND made current
        return ctx

    def current(self):
        ctx = C.c_void_p()
        self.ok(self.cu.cuCtxGetCurrent(C.byref(ctx)))
        return ctx.value

    def pop(self):
        ctx = C.c_void_p()
        self.ok(self.cu.cuCtxPopCurrent_v2(C.byref(ctx)))
        return ctx.value

    def push(self, ctx):
        self.ok(self.cu.cuCtxPushCurrent_v2(ctx))

    def destroy(self, ctx):
        self.ok(self.cu.cuCtxDestroy_v2(ctx))

    def alloc(self, nbytes):
        p = C.c_uint64()
        self.ok(self.cu.cuMemAlloc_v2(C.byref(p), C.c_size_t(nbytes)))
        return p.value

    def free(self, p):
        self.ok(self.cu.cuMemFree_v2(C.c_uint64(p)))

    def h2d(self, p, a):
        self.ok(self.cu.cuMemcpyHtoD_v2(C.c_uint64(p), a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)))

    def d2h(self, a, p):
        self.ok(self.cu.cuMemcpyDtoH_v2(a.ctypes.data_as(C.c_void_p), C.c_uint64(p), C.c_size_t(a.nbytes)))


def test_sampler_in_a_non_primary_context_with_borrowed_buffers():
    drv = Driver()
    before = drv.current()
    ctx = drv.ctx_create(0)
    try:
        assert drv.current() == ctx.value and ctx.value != before
        W, seed = G.W, G.SEED
        params = f32([-1, 2])
        p_dev = drv.alloc(params.nbytes)                       # the cuda-float params vector the sampler borrows
        drv.h2d(p_dev, params)
        out_dev = drv.alloc(4 * W)                             # the cuda-float result matrix of sample!
        factory = bb.B200BayaderaFactory(wgs=G.WGS, current_context=True)
        sf = factory.mcmc_factory(models.UNIFORM)
        gpu = sf.create_sampler(seed, W, DeviceParams(p_dev, params.size))
        gpu.init(seed).init_position(seed, f32([-1, 2]))
        assert drv.current() == ctx.value                      # the caller's context stack is left as found
        # the reference's golden chain (nvidia_gtx_test.clj:202-212), results written into the caller's buffer
        got = np.zeros(W, dtype=np.float32)
        for want in G.UNIFORM_SAMPLES:
            gpu.sample_into_device(W, out_dev)
            factory.synchronize()
            drv.d2h(got, out_dev)
            assert np.array_equal(got[:4], f32(want)), (got[:4], want)
        # the caller switches to another context: the engine must push its own around every call, and restore
        assert drv.pop() == ctx.value
        outside = drv.current()
        gpu.burn_in(8, 2.0)
        rate = gpu.acc_rate(2.0)
        hist = gpu.histogram(2)
        xs = gpu.get_state()["xs"].reshape(-1)
        assert drv.current() == outside
        cpu = orc.OracleStretch(models.UNIFORM, seed, W, params, wgs=G.WGS)
        cpu.init(seed).init_position(seed, f32([-1, 2]))
        for _ in G.UNIFORM_SAMPLES:
            cpu.sample()
        cpu.burn_in(8, 2.0)
        rate_c = cpu.acc_rate(2.0)
        hc = cpu.histogram(2)
        assert rate == rate_c
        assert np.array_equal(gpu.histogram_counts(), hc["counts"])
        assert np.array_equal(hist.bin_ranks, hc["bin-ranks"])
        assert np.array_equal(xs, cpu.xs)
        # a dataset-engine call on a device matrix of that context
        drv.push(ctx)
        data = np.random.default_rng(1).random((4096, 3), dtype=np.float32)
        d_dev = drv.alloc(data.nbytes)
        drv.h2d(d_dev, data)
        mean = np.zeros(3, dtype=np.float32)
        bb._lib.check(factory._L.bay_dataset_mean(factory._h, C.c_void_p(d_dev), 1, 3, 4096, 0, 3, mean))
        assert np.allclose(mean, data.mean(axis=0), atol=1e-5)
        gpu.release()
        sf.release()
        factory.release()
        for p in (p_dev, out_dev, d_dev):
            drv.free(p)
        assert drv.current() == ctx.value
        drv.pop()
    finally:
        drv.destroy(ctx)
    assert drv.current() == before


def test_create_current_without_a_context_fails_loudly():
    import threading
    result = {}

    def run():          # a fresh thread has no current context
        try:
            bb.B200BayaderaFactory(wgs=256, current_context=True)
            result["err"] = None
        except bb.BayaderaError as e:
            result["err"] = str(e)

    t = threading.Thread(target=run)
    t.start()
    t.join()
    assert result["err"] and "no CUDA context is current" in result["err"]
