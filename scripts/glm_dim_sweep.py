"""Throughput of the GLM (logistic regression) path against the model dimension: D <= 64 runs one 64-wide K chunk of
the tcgen05 kernel, 64 < D <= 128 two, D > 128 the fp32 SIMT kernel.  python scripts/glm_dim_sweep.py [rows]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bayadera_b200 as bb  # noqa: E402
from bayadera_b200 import models  # noqa: E402
from bench import logreg_rows_device  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    f = bb.B200BayaderaFactory(device=0, wgs=256)
    for d in (16, 64, 96, 128):
        data = logreg_rows_device(torch, rows, d, 1, torch.device("cuda", 0))
        m = models.logistic_regression_model(d)
        s = f.mcmc_factory(m).create_sampler(1, 1024, bb.DeviceParams(data.data_ptr(), data.numel(), owner=data))
        s.init_position(2, m.limits_array())
        s.burn_in(3, 1.2)
        f.synchronize()
        t0 = time.perf_counter()
        s.burn_in(10, 1.2)
        f.synchronize()
        dt = time.perf_counter() - t0
        print(f"D={d}: {1024 * 10 / dt / 1e3:.1f} K walker-steps/s on {rows} rows "
              f"({1024 * 10 * rows / dt / 1e12:.2f} T datum-evals/s)", flush=True)
        del s, data


if __name__ == "__main__":
    main()
