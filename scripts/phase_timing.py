"""Wall-clock of the MCMC phases per workload (burn-in!, run-sampler!, sample!, histogram!) — a quick profile of
what the reference's `mix!` + `sample!` + `histogram!` workflow spends where.  python scripts/phase_timing.py c2 c3 c5"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bayadera_b200 as bb  # noqa: E402
import bench  # noqa: E402


def main():
    factory = bb.B200BayaderaFactory(device=0, wgs=256)
    for name in sys.argv[1:] or ["c1", "c2", "c3", "c5"]:
        wl = bench.workload(name)
        s = factory.mcmc_factory(wl["model"]).create_sampler(1, wl["walkers"], wl["params"])
        s.init_position(2, wl["limits"])
        n = wl["moves"] * 4
        s.burn_in(n, wl["a"])
        factory.synchronize()
        out = {}
        for phase, fn in (("burn-in!", lambda: s.burn_in(n, wl["a"])), ("run-sampler!", lambda: s.run_sampler(max(n, 64), wl["a"])),
                          ("acc-rate!", lambda: s.acc_rate(wl["a"])), ("histogram!", lambda: s.histogram(1)),
                          ("mean", lambda: s.mean())):
            fn()
            factory.synchronize()
            t0 = time.perf_counter()
            fn()
            factory.synchronize()
            out[phase] = time.perf_counter() - t0
        steps = {"burn-in!": n, "run-sampler!": max(n, 64), "acc-rate!": 1}
        line = [f"{name}:"]
        for k, v in out.items():
            if k in steps:
                line.append(f"{k} {v * 1e3:.2f} ms ({wl['walkers'] * steps[k] / v / 1e9:.2f} G ws/s)")
            else:
                line.append(f"{k} {v * 1e3:.2f} ms")
        print("  ".join(line), flush=True)


if __name__ == "__main__":
    main()
