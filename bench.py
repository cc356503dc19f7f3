#!/usr/bin/env python
"""bench.py — walker-steps/s of the stretch-move hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c5] [--impl b200|reference]

One bench "step" = `moves_per_step` calls of move-bare! (odd + even half-ensemble launch each) over the whole
ensemble, i.e. W * moves_per_step walker-steps (= logpdf evaluations).  Timed on the device with CUDA events on
the engine's stream; L2 is flushed (256 MiB write) between timed steps; SM clocks and throttle reasons are
sampled through NVML during the timed region.  `value` has the ensemble resident in HBM; `e2e` goes through the
C-ABI with HOST buffers (ensemble uploaded from pinned memory, advanced, downloaded) inside the timed region.

--impl reference times the CPU oracle (oracle/, the C restatement of the reference's kernels — the reference's
own OpenCL/CUDA-JIT kernels cannot run here: no JVM, no POCL) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "walker-steps/sec"
UNIT = "walker-steps/s"


def f32(v):
    return np.asarray(v, dtype=np.float32)


# ------------------------------------------------------------------------------------------ workloads --
def workload(name: str):
    """BASELINE.json configs (SURVEY §8d synthetic inputs).  Returns a dict describing the job."""
    from bayadera_b200 import models
    if name == "c1":
        return dict(key="c1", desc="1-D Gaussian(0,1), 2^14 walkers (BASELINE configs[0])", model=models.GAUSSIAN,
                    params=f32([0, 1]), limits=f32([-7, 7]), walkers=2 ** 14, moves=256, a=2.0,
                    cpu_walkers=2 ** 14, ref_stem="gaussian_d1")
    if name == "c2":
        params = np.concatenate([models.binomial_lik_params(50, 15), models.beta_params(3, 2)])
        return dict(key="c2", desc="beta-binomial coin posterior, 2^18 walkers (BASELINE configs[1])",
                    model=models.beta_binomial_posterior(), params=params, limits=f32([0, 1]), walkers=2 ** 18,
                    moves=128, a=2.0, cpu_walkers=2 ** 16)
    if name == "c3":
        m = models.therapeutic_touch_model()
        return dict(key="c3", desc="hierarchical therapeutic-touch, D=30, 280 trials, 2^16 walkers (BASELINE configs[2])",
                    model=m, params=models.therapeutic_touch_data(), limits=m.limits_array(), walkers=2 ** 16,
                    moves=32, a=2.0, cpu_walkers=2 ** 13)
    if name == "c5":
        m = models.mvn_model(100)
        return dict(key="c5", desc="100-D correlated Gaussian, 2^20 walkers (BASELINE configs[4])", model=m,
                    params=models.mvn_params(100)[0], limits=m.limits_array(), walkers=2 ** 20, moves=4, a=1.25,
                    cpu_walkers=2 ** 11)
    if name == "rowadd":
        # generic row-additive path: posterior of (mu, sigma) given 10^6 i.i.d. Gaussian data
        m = models.gaussian_mean_sd_posterior()
        rng = np.random.default_rng(3)
        n = 10 ** 6
        params = np.concatenate([(2.5 + 1.7 * rng.standard_normal(n)).astype(np.float32), f32([0.0, 5.0, 0.5])])
        return dict(key="rowadd", desc="Gaussian (mu, sigma) posterior of 10^6 data, generic row-additive path", model=m,
                    params=params, limits=m.limits_array(), walkers=4096, moves=4, a=2.0, cpu_walkers=512,
                    data_len=n)
    if name == "c4":
        d, rows = 64, 10 ** 7
        m = models.logistic_regression_model(d)
        return dict(key="c4", desc="Bayesian logistic regression, D=64, 10^7 synthetic rows (BASELINE configs[3])",
                    model=m, params=None, limits=m.limits_array(), walkers=4096, moves=8, a=1.2,
                    cpu_walkers=512, rows=rows, cpu_rows=4000, glm=True)
    raise SystemExit(f"unknown workload {name}")


def logreg_rows_host(rows: int, d: int, seed: int = 2024) -> np.ndarray:
    """Synthetic c4 data: X ~ N(0,1), theta* ~ N(0,1/8), y ~ Bernoulli(sigmoid(X theta*));
    packed as the model expects: rows [y, x_1..x_d] followed by the hyper-parameter 1/(2*10^2)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((rows, d), dtype=np.float32)
    theta = (rng.standard_normal(d) / np.sqrt(8.0)).astype(np.float32)
    y = (rng.random(rows) < 1.0 / (1.0 + np.exp(-(x @ theta)))).astype(np.float32)
    return np.concatenate([np.concatenate([y[:, None], x], axis=1).reshape(-1), f32([1.0 / 200.0])])


def logreg_rows_device(torch, rows: int, d: int, seed: int, device):
    """Same distribution generated on the device (nothing shipped over PCIe for the 2.6 GB matrix).  `seed` draws
    the rows (a different stream per row shard); the generating coefficients theta* are the same on every rank, so
    the union of the shards is ONE logistic-regression dataset.  Returns (params vector, theta*)."""
    gt = torch.Generator(device=device)
    gt.manual_seed(2024)
    theta = torch.randn(d, generator=gt, device=device) / (8.0 ** 0.5)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(rows * (d + 1) + 1, dtype=torch.float32, device=device)
    mat = out[:-1].view(rows, d + 1)
    mat[:, 1:].normal_(generator=g)
    p = torch.sigmoid(mat[:, 1:] @ theta)
    mat[:, 0] = (torch.rand(rows, generator=g, device=device) < p).float()
    out[-1] = 1.0 / 200.0
    return out, theta.double().cpu().numpy()


def algorithmic_bytes_per_walker_step(dim: int, p_acc: float) -> float:
    """4(2D+1) read + 4(D+1) p_acc written."""
    return 4.0 * (2 * dim + 1) + 4.0 * (dim + 1) * p_acc


def ncu_traffic(key: str, kernel: str) -> dict:
    """roofline.traffic: DRAM bytes per launch of the dominant kernel, measured once under `ncu --set full`
    (profiles/ncu_traffic.json names the capture); None when no capture of this workload's kernel is committed."""
    try:
        entry = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text()).get(key)
    except (OSError, ValueError):
        entry = None
    if not entry or entry.get("kernel") != kernel:
        return {"traffic": None}
    return {"traffic": entry["bytes"], "traffic_source": entry["source"]}


# --------------------------------------------------------------------------------------------- clocks --
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons through NVML every 20 ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.02)

    def finish(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference --
def job_config(wl: dict, args) -> dict:
    """The workload both arms are measured on (identical in the b200 and the reference line)."""
    cfg = {"workload": wl["desc"], "walkers": wl["walkers"], "dim": wl["model"].dimension,
           "moves_per_step": wl["moves"], "a": wl["a"], "wgs": args.wgs}
    if wl.get("rows"):
        cfg["rows"] = wl["rows"]
    return cfg


def cpu_rate(wl: dict, budget_s: float, walkers: int):
    """walker-steps/s of the CPU oracle (all host threads) on a bounded sample of the workload.  The cost of a
    walker-step does not depend on the ensemble size (every walker evaluates the model once per step), so the rate
    measured on `walkers` walkers is the rate of the config's ensemble."""
    from oracle import oracle as orc
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for all host cores explicitly
    orc.lib().orc_set_num_threads(os.cpu_count() or 1)
    params = wl["params"]
    if params is None:                                # c4: bounded row sample, same generator family
        params = logreg_rows_host(wl["cpu_rows"], wl["model"].dimension)
    from oracle import ref_text
    if wl.get("ref_stem") and ref_text.available(wl["ref_stem"]):
        # the reference's OWN kernel text compiled for the host (oracle/_ref, see oracle/ref_shim/cl_shim.h)
        s = ref_text.ReferenceTextStretch(wl["ref_stem"], wl["model"], 123, walkers, params, wgs=256)
        wl["cpu_kind"] = "reference"
    else:
        s = orc.OracleStretch(wl["model"], 123, walkers, params, wgs=256)
        wl["cpu_kind"] = "port"
    s.init_position(123, wl["limits"])
    s.a_bare = wl["a"]
    s.move_bare()                                   # warm-up + calibration
    t0 = time.perf_counter()
    s.move_bare()
    per_move = max(time.perf_counter() - t0, 1e-6)
    n = int(max(2, min(10000, budget_s / per_move)))
    t0 = time.perf_counter()
    for _ in range(n):
        s.move_bare()
    dt = time.perf_counter() - t0
    rate = walkers * n / dt
    if wl.get("rows"):                                # cost is linear in the rows: extrapolate to the full dataset
        rate *= wl["cpu_rows"] / wl["rows"]
    return rate, n, dt, orc.lib().orc_num_threads()


def run_reference(args, wl: dict, rank: int, world: int):
    if rank != 0:
        return
    walkers = wl["cpu_walkers"]
    budget = 4.0
    for _ in range(args.warmup):
        cpu_rate(wl, 0.2, walkers)
    rates, total_t = [], 0.0
    for _ in range(args.steps):
        r, n, dt, threads = cpu_rate(wl, budget, walkers)
        rates.append(r)
        total_t += dt
    value = float(np.mean(rates))
    sample = (f"{walkers} of the config's {wl['walkers']} walkers (cost per walker-step is independent of the ensemble "
              f"size) x ~{n} moves per step ({budget:.0f} s budget) of workload {wl['key']}")
    if wl.get("rows"):
        sample += f", on {wl['cpu_rows']} of {wl['rows']} rows, rate scaled linearly by rows (cost is linear in rows)"
    kind = wl.get("cpu_kind", "port")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
            "higher_is_better": True, "scaling": "strong" if wl.get("glm") else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": job_config(wl, args),
            "engine": ("the reference's own OpenCL kernel text compiled by gcc for the host (oracle/_ref), "
                       "OpenMP over work-items") if kind == "reference" else
                      "CPU oracle (C restatement of the reference kernels, OpenMP over walkers)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------- summary --
def run_summary(args):
    """Not a bench line of record: device time of the summary engines on the config-5
    ensemble (100 x 2^20 fp32 = 419 MB, larger than L2): histogram = min/max pass + binning pass + finish,
    mean and variance = one fused pass each call.  Reported as GB/s of algorithmic bytes (4*D*n per pass)."""
    import torch

    import bayadera_b200 as bb
    from bayadera_b200 import models
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    m = models.mvn_model(100)
    factory = bb.B200BayaderaFactory(device=0, stream=stream.cuda_stream, wgs=args.wgs)
    s = factory.mcmc_factory(m).create_sampler(1, 2 ** 20, models.mvn_params(100)[0]).init_position(2, m.limits_array())
    nbytes = 4.0 * 100 * 2 ** 20
    out = {}
    def fresh_mean():
        s.move_bare_half(0)          # a new ensemble state: the moment sums are cached per state
        return s.mean()

    for name, fn, passes in (("AoS mirror: histogram (min/max + bin)", lambda: s.histogram(1), 2),
                             ("mean (incl. one half-ensemble move per call)", fresh_mean, 1),
                             ("variance after mean (same state: no pass)", s.variance, 0)):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        out[name] = {"ms_per_call": ms, "algorithmic_GBps": passes * nbytes / (ms * 1e-3) / 1e9, "passes": passes}
    # the same ensemble held by a sampler that keeps its SoA matrix current (generic kernels): the SoA passes
    os.environ["BAY_QUADFORM_TC"] = "0"
    g = factory.mcmc_factory(m).create_sampler(1, 2 ** 20, models.mvn_params(100)[0]).init_position(2, m.limits_array())
    del os.environ["BAY_QUADFORM_TC"]
    for name, fn, passes in (("SoA kernels: histogram (min/max + bin)", lambda: g.histogram(1), 2),):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        out[name] = {"ms_per_call": ms, "algorithmic_GBps": passes * nbytes / (ms * 1e-3) / 1e9, "passes": passes}
    g.release()
    # BASELINE configs[4]'s summary workload: histogram! with 256 cycles = 2^28 samples of the 100-D ensemble
    # (nvidia_gtx.clj:482-512: min/max of the first snapshot, then 255 x (move-bare! + binning) with those limits)
    s.burn_in(32, 1.25)
    cycles = 256
    s.histogram(2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    s.histogram(cycles)
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    a.record(stream)
    s.burn_in(cycles - 1, 1.25)
    b.record(stream)
    torch.cuda.synchronize()
    ms_moves = a.elapsed_time(b)
    out["histogram! 256 cycles (2^28 samples x 100 dims)"] = {
        "ms": ms, "ms_per_cycle": ms / cycles, "samples_per_s": cycles * 2 ** 20 / (ms * 1e-3),
        "of_which_moves_ms": ms_moves,
        "binning_GBps": (cycles * nbytes) / ((ms - ms_moves) * 1e-3) / 1e9 if ms > ms_moves else None,
        "note": "binning reads the AoS mirror the tensor-core move maintains; limits from the first snapshot"}
    ds = factory.dataset_engine()
    data = np.random.default_rng(0).random((31 * 2 ** 16, 22), dtype=np.float32)       # T/core_test.clj:19
    ds.histogram(data[:4096])
    t0 = time.perf_counter()
    ds.histogram(data)
    out["dataset histogram 22 x 2031616 from host (e2e)"] = {"ms_per_call": 1e3 * (time.perf_counter() - t0)}
    ddev = torch.from_numpy(data).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lim = np.zeros((22, 2), dtype=np.float32)
    from bayadera_b200._lib import check, ptr
    import ctypes as C
    for _ in range(5):
        check(factory._L.bay_dataset_histogram(factory._h, C.c_void_p(ddev.data_ptr()), 1, 22, data.shape[0], 0, 22, ptr(lim),
                                               None, None, None))
    dt = (time.perf_counter() - t0) / 5
    out["dataset histogram 22 x 2031616 on the device"] = {"ms_per_call": 1e3 * dt,
                                                           "algorithmic_GBps": 2 * data.nbytes / dt / 1e9, "passes": 2}
    print(json.dumps({"workload": "summary engines on the c5 ensemble (100 x 2^20)", "wgs": args.wgs, "results": out}))


# ----------------------------------------------------------------------------------------------- main --
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BAY_WORKLOAD", "c4"))
    ap.add_argument("--moves-per-step", type=int, default=0)
    ap.add_argument("--walkers", type=int, default=0)
    ap.add_argument("--wgs", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-mode-a", action="store_true", help="skip the c5 walker-partition companion measurement")
    ap.add_argument("--strong", action="store_true",
                    help="walker-partitioned workloads: keep the config's TOTAL ensemble and split it over the GPUs "
                         "(default: the config's ensemble per GPU, weak scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "summary":
        run_summary(args)
        return
    wl = workload(args.workload)
    if args.walkers:
        wl["walkers"] = args.walkers
    if args.moves_per_step:
        wl["moves"] = args.moves_per_step

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist

    import bayadera_b200 as bb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: bayadera_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, D, M, a, model = wl["walkers"], wl["model"].dimension, wl["moves"], wl["a"], wl["model"]
    # an explicit (non-default) stream: the engine enqueues on it and the timing events are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    factory = bb.B200BayaderaFactory(device=local_rank, stream=stream.cuda_stream, wgs=args.wgs)
    assert factory.stream() == stream.cuda_stream
    sharded = bool(wl.get("glm")) and world > 1          # SURVEY §8e mode B
    partition = (not wl.get("glm")) and world > 1        # SURVEY §8e mode A
    if world > 1:
        # mode B (GLM): walkers replicated, dataset rows sharded, per-walker partial sums all-reduced by the engine
        #   over NCCL; total work fixed (10^7 rows) -> strong scaling.
        # mode A (generic models): ONE global ensemble of world*W walkers, rank r updates its slice of each half and
        #   the slices are all-gathered every half-step; per-GPU walkers fixed -> weak scaling.
        from bayadera_b200.distributed import init_engine_comm, shard_rows
        init_engine_comm(factory, rank, world, torch.device("cuda", local_rank))
    sfactory = factory.mcmc_factory(model)
    params = wl["params"]
    theta_star = None
    if params is None:
        if sharded:
            b, e = shard_rows(wl["rows"], world, rank)
            local_rows, seed = e - b, 2025 + 1000 * rank
        else:
            local_rows, seed = wl["rows"], 2025
        rows_dev, theta_star = logreg_rows_device(torch, local_rows, D, seed, torch.device("cuda", local_rank))
        params = bb.DeviceParams.from_torch(rows_dev)
    # every rank drives the same (replicated or partitioned) ensemble: identical seeds everywhere
    if partition and args.strong:
        W = max(W // world // (2 * args.wgs), 1) * 2 * args.wgs     # this GPU's share of the config's ensemble
    W_global = W * world if partition else W
    # one-off cost of create-sampler (G/:548-610): for the GLM path it includes splitting the dataset into the bf16
    # hi/lo planes the tensor cores read and the X^T y / column-sum passes — reported, not part of a step
    barrier()
    t0 = time.perf_counter()
    sampler = sfactory.create_sampler(123, W_global, params)
    factory.synchronize()
    setup_s = time.perf_counter() - t0
    sampler.init_position(1000, wl["limits"])
    sampler.burn_in(max(64, M), a)                      # leave the initial box before timing
    p_acc = sampler.acc_rate(a)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    for _ in range(args.warmup):
        sampler.burn_in(M, a)
    barrier()

    clocks = ClockSampler(local_rank)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = bb.launch_count()
    barrier()
    for s_ev, e_ev in ev:
        flush.fill_(1)                                  # evict the ensemble from L2 between timed steps
        s_ev.record(stream)
        sampler.burn_in(M, a)
        e_ev.record(stream)
    barrier()
    launches = bb.launch_count() - launches0
    clk = clocks.finish()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- e2e: host ensemble -> device, advance, device -> host, through the C-ABI with pinned HOST buffers ----
    xs0, lp0 = sampler.get_state64()
    xs_host = torch.from_numpy(xs0).pin_memory().numpy()
    lp_host = torch.from_numpy(lp0).pin_memory().numpy()
    sampler.set_state64(xs_host, lp_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sampler.set_state64(xs_host, lp_host)          # H2D: positions (fp32) + log-densities (fp64)
        sampler.burn_in(M, a)
        sampler.get_state64(xs_host, lp_host)          # D2H into the same pinned buffers
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- parity self-checks, outside every timed region (bayadera_b200/selfcheck.py) --------------------------------
    parity = {}
    if not args.no_parity_check:
        from bayadera_b200 import selfcheck
        if wl.get("glm"):
            # the timed sampler itself, at the full dataset: Δlogp of stretch proposals between points scattered at the
            # posterior's own scale (sd ~ 1/sqrt(0.2 rows)) around the generating coefficients — where the accept
            # test consumes O(1)..O(10) differences of ~7e6 sums — tensor-core path against the fp64 traversal of
            # the same rows (all-reduced over the row shards)
            parity["glm_dlogp_vs_fp64"] = selfcheck.glm_delta_logp(sampler, theta_star, (0.2 * wl["rows"]) ** -0.5,
                                                                   pairs=128, a=a)
        if world > 1:
            single = bb.B200BayaderaFactory(device=local_rank, stream=stream.cuda_stream, wgs=args.wgs)
            parity["mode_a_partition_vs_1gpu"] = selfcheck.mode_a_bit_identity(factory, single, world)
            parity["mode_b_row_shards_vs_1gpu"] = selfcheck.mode_b_replicas(factory, single, rank, world)
            parity["mode_b_rowadd_shards_vs_1gpu"] = selfcheck.mode_b_rowadd(factory, single, rank, world)
        parity["ok"] = all(v.get("ok", False) for v in parity.values())
        ok_t = torch.tensor([1 if parity["ok"] else 0], dtype=torch.int32, device="cuda")
        if world > 1:
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(int(ok_t[0]))
        parity["verdict"] = "ok" if parity["ok"] else "MISMATCH"

    # ---- mode A companion measurement (BASELINE configs[4]): the 100-D Gaussian with the config's TOTAL ensemble of
    # 2^20 walkers partitioned over the GPUs (strong scaling), so that the driver's 1/2/4/8-GPU runs of this file also
    # carry the walker-partition path.  Same timing rules as the main line.
    mode_a = None
    if wl["key"] == "c4" and not args.no_mode_a:
        mode_a = run_mode_a(args, torch, dist, bb, factory, stream, flush, barrier, rank, world)

    total_ws = (1.0 if sharded else float(world)) * W * M * args.steps
    value = total_ws / (dev_ms * 1e-3)
    e2e_value = total_ws / (e2e_ms * 1e-3)
    per_launch_ms = dev_ms / (2.0 * M * args.steps)
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(peaks_file.read_text()) if peaks_file.exists() else None
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    if wl.get("glm"):
        # dominant kernel: the dataset likelihood (one launch per half-step; propose/finish/accept are ~1 % of it).
        # Dense contraction X(rows x D) . Theta(D x H): 2*rows*D flops per walker-step, bf16-dense peak
        # as the denominator (an fp32-accurate 3-term split can reach at most 1/3 of it).
        # one launch handles up to 1024 walkers (4 groups of 256) against all local rows; a half-step of H = W/2
        # walkers is ceil(H/1024) back-to-back launches, so the per-launch duration is the half-step time / that count
        kernel_name = "k_glm_loglik_tc"
        per = 1024
        n_groups = -(-(W // 2) // per)
        per_launch_ms = per_launch_ms / n_groups
        flops_per_launch = 2.0 * (wl["rows"] / (world if sharded else 1)) * D * min(W // 2, per)   # per GPU
        terms = {"4": 4, "5": 5}.get(os.environ.get("BAY_GLM_TERMS", ""), 3)
        peak = float(peaks["bf16_tflops_sustained"]) if peaks else 1400.0
        achieved = flops_per_launch / (per_launch_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                **ncu_traffic(wl["key"], kernel_name), "kernel": kernel_name, "peak_source": peak_src,
                "flops_per_launch": flops_per_launch,
                "launch_us": per_launch_ms * 1e3,
                "executed_mma_tflops": terms * achieved, "mma_terms": terms, "executed_frac": terms * achieved / peak,
                "note": "achieved counts ALGORITHMIC flops (2*rows*D per walker).  fp32-level accuracy on 16-bit tensor-core "
                        f"inputs takes {terms} MMAs per product (theta - theta0 and the dataset in two fp16 pieces each; three "
                        "bf16 pieces of theta with BAY_GLM_TERMS=4), so the tensor pipe executes that multiple of the "
                        f"algorithmic flops and frac cannot exceed 1/{terms}; the kernel runs at the board's power cap (see clocks.reasons) "
                        "with the XU (exp2) pipe 87 % busy; DESIGN.md 4.2 has the pipe utilisations ncu reports",
                "hbm_view": {"bytes_per_launch": flops_per_launch / min(W // 2, per) * 2.0,
                             "achieved_GBps": flops_per_launch / min(W // 2, per) * 2.0 / (per_launch_ms * 1e-3) / 1e9,
                             "note": "dataset bytes (bf16 hi+lo planes = 4 B/element) streamed from HBM once per launch"}}
    else:
        kernel_name = sampler_kernel_name(sampler)
        bytes_per_launch = (W / 2) * algorithmic_bytes_per_walker_step(D, p_acc)
        peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                **ncu_traffic(wl["key"], kernel_name), "kernel": kernel_name, "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch,
                "launch_us": per_launch_ms * 1e3}
        if wl.get("data_len"):
            # row-additive posterior: the dataset is read once per 128 walkers per half-step (not once per walker);
            # datum-evals = walker-steps x data_len; dataset bytes per launch = 4 * data_len * ceil(H / 128) from L2/HBM
            roof["datum_evals_per_s"] = value * wl["data_len"]
            roof["kernel"] = "bay_rowadd_loglik"
            blocks = -(-(W // 2) // 128)
            roof["bytes_per_launch"] = 4.0 * wl["data_len"] * blocks
            roof["achieved"] = roof["bytes_per_launch"] / (per_launch_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["note"] = ("the 4 MB dataset is L2-resident and re-read by each of the 16 walker blocks: the kernel is "
                            "bound by the FMA pipe (datum_evals_per_s), not by HBM; frac is the dataset stream against "
                            "the HBM peak, for completeness")
        elif W * D * 4 < 64 << 20:
            # the state of configs 1-3 is L2-resident; what bounds them is the latency of a half-step
            # (launch or grid barrier + a handful of dependent L2 round trips), not HBM — read `launch_us`, not `frac`
            roof["note"] = ("ensemble is L2-resident: latency-bound; launch_us (one half-step) against the ~2 us "
                            "grid-barrier floor is the figure of merit, frac is reported for completeness only")

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            r, n, dt, threads = cpu_rate(wl, 12.0, wl["cpu_walkers"])
            cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": wl.get("cpu_kind", "port"),
                   "sample": f"{wl['cpu_walkers']} walkers x {n} moves ({dt:.1f} s) of workload {wl['key']}, CPU oracle (OpenMP); "
                             "cost per walker-step is independent of the ensemble size"
                             + (f"; on {wl['cpu_rows']} of {wl['rows']} rows, rate scaled linearly by rows" if wl.get("rows") else "")}
        if wl.get("glm"):   # statically compiled kernel: numbers from the nvcc -Xptxas -v log (profiles/)
            info = {"block": 576, "grid": "1 CTA per SM (persistent), 4 walker groups x 37 row-tile slices",
                    "launches_per_half_step": -(-(W // 2) // 1024)}
        elif sampler.uses_quadform():
            info = {"registers": 128, "block": 512, "grid": "1 CTA per SM (persistent), 128-walker tiles",
                    "shared_bytes": 1024 + 131072 + 8192}
        else:
            info = sfactory.kernel_info("bay_stretch_bare")
        wl["walkers"] = W_global if partition else W
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if (wl.get("glm") or args.strong) else "weak", "vs_baseline": None,
                "dtype": "f32 (tensor cores on fp16 pieces: theta - theta0 and the dataset in 2 each, 3 MMAs per product, fp32 accumulate)"
                         if wl.get("glm") else ("f32 (tensor cores on fp16 hi/lo pieces, fp32 accumulate)"
                                                if sampler.uses_quadform() else "f32"),
                "data": "synthetic",
                "config": job_config(wl, args),
                "details": {"walkers_per_gpu": W, "acceptance": round(p_acc, 4),
                            "parallelism": "single GPU" if world == 1 else (
                                f"rows sharded over {world} GPUs, walkers replicated, NCCL all-reduce of per-walker sums"
                                if sharded else f"one ensemble of {W_global} walkers partitioned over {world} GPUs"),
                            "l2": "flushed between timed steps (256 MiB write)",
                            "kernel": {"name": kernel_name, **info}},
                "e2e": {"value": e2e_value, "unit": UNIT,
                        "h2d_bytes_per_step": int(xs_host.nbytes + lp_host.nbytes),
                        "d2h_bytes_per_step": int(xs_host.nbytes + lp_host.nbytes),
                        "what": "bay_set_state64(pinned host ensemble) + burn-in! + bay_get_state64(pinned host) per step"},
                "setup": {"sampler_create_s": setup_s,
                          "what": "one-off bay_sampler_create_dev: dataset -> bf16 hi/lo planes, X^T y, column sums"
                                  if wl.get("glm") else "one-off bay_sampler_create"},
                "gpu_launches": int(launches),
                "roofline": roof,
                "parity_check": parity.get("verdict"), "parity": parity,
                "mode_a": mode_a,
                "cpu_baseline": cpu, "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sampler_kernel_name(sampler) -> str:
    return "k_quadform_move_tc" if getattr(sampler, "uses_quadform", lambda: False)() else "bay_stretch_bare"


def run_mode_a(args, torch, dist, bb, factory, stream, flush, barrier, rank, world):
    """c5 (100-D correlated Gaussian), 2^20 walkers in TOTAL, partitioned by walker over the GPUs of this run."""
    from bayadera_b200 import models
    m = models.mvn_model(100)
    W_total, moves, a, steps = 2 ** 20, 4, 1.25, max(5, min(args.steps, 10))
    s = factory.mcmc_factory(m).create_sampler(321, W_total, models.mvn_params(100)[0]).init_position(322, m.limits_array())
    s.burn_in(64, a)
    p_acc = s.acc_rate(a)
    for _ in range(3):
        s.burn_in(moves, a)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s_ev, e_ev in ev:
        flush.fill_(1)
        s_ev.record(stream)
        s.burn_in(moves, a)
        e_ev.record(stream)
    barrier()
    ms = torch.tensor([sum(x.elapsed_time(y) for x, y in ev)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    value = W_total * moves * steps / (ms * 1e-3)
    H_local = W_total // 2 // world
    out = {"workload": "100-D correlated Gaussian, 2^20 walkers in total partitioned over the GPUs (BASELINE configs[4])",
           "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "scaling": "strong", "steps": steps,
           "moves_per_step": moves, "ms_per_step": ms / steps, "acceptance": round(p_acc, 4),
           "kernel": sampler_kernel_name(s),
           "half_step_us": ms * 1e3 / (2.0 * moves * steps),
           "hbm_GBps_per_gpu": H_local * algorithmic_bytes_per_walker_step(100, p_acc) / (ms * 1e-3 / (2.0 * moves * steps)) / 1e9}
    if world > 1:
        # partner rows that live on another rank: (R-1)/R of this rank's walkers fetch one 400-byte row each
        nv = (world - 1) / world * H_local * 400.0
        out["nvlink_bytes_per_half_step_per_gpu"] = nv
        out["nvlink_roofline_us"] = nv / 770e9 * 1e6
        out["nvlink_note"] = "bytes that must cross NVLink / the measured 770 GB/s per direction per GPU (B200_PROFILING.md)"
    s.release()
    return out


if __name__ == "__main__":
    main()
