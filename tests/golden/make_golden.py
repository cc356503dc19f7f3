"""Regenerate tests/golden/*.npz from the reference's own test fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

acor_fixtures.npz  <- /root/reference/test/clojure/uncomplicate/bayadera/internal/acor-data-{67,367,112640}
                      (one float per line; consumed at T/internal/nvidia_gtx_test.clj:320-358)
The numeric goldens of the reference's Midje tests (positions, block sums, accept counts …) are
short enough to be written literally, with their file:line, in tests/goldens.py.
"""
from pathlib import Path

import numpy as np

REF = Path("/root/reference/test/clojure/uncomplicate/bayadera/internal")
OUT = Path(__file__).resolve().parent


def main():
    data = {}
    for n in (67, 367, 112640):
        vals = np.loadtxt(REF / f"acor-data-{n}", dtype=np.float64).astype(np.float32)
        assert vals.size == n, (n, vals.size)
        data[f"acor_{n}"] = vals
    np.savez_compressed(OUT / "acor_fixtures.npz", **data)
    print("wrote", OUT / "acor_fixtures.npz", {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
