"""Where does the tensor-core path's error against fp64 come from?  (diagnostic, run on the GPU box)"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bayadera_b200 as bb
from bayadera_b200 import models

def bf16(x):
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)

d, rows, pairs = 64, 1_500_000, 512
rng = np.random.default_rng(1)
theta = (rng.standard_normal(d) / np.sqrt(8)).astype(np.float32)
f = bb.B200BayaderaFactory(device=0, wgs=256)
model = models.logistic_regression_model(d)
sf = f.mcmc_factory(model)
LOG2E = np.float32(1.4426950408889634)
for xmode in ("normal", "bf16-exact planes"):
    x = rng.standard_normal((rows, d)).astype(np.float32)
    if xmode != "normal":
        x = (bf16(x * LOG2E) / LOG2E).astype(np.float32)
    y = (rng.random(rows) < 1 / (1 + np.exp(-(x @ theta)))).astype(np.float32)
    params = np.concatenate([np.concatenate([y[:, None], x], axis=1).reshape(-1), np.float32([1 / 200.0])])
    s = sf.create_sampler(1, 2 * pairs, params)
    for tmode in ("normal", "bf16-exact theta", "tiny theta"):
        scale = (0.2 * rows) ** -0.5
        cur = (theta[None, :] + scale * rng.standard_normal((pairs, d))).astype(np.float32)
        oth = (theta[None, :] + scale * rng.standard_normal((pairs, d))).astype(np.float32)
        if tmode == "tiny theta":
            cur, oth = (cur - theta).astype(np.float32), (oth - theta).astype(np.float32)
        z = (((0.2) * rng.random(pairs) + 1.0) ** 2 / 1.2).astype(np.float32)
        prop = (oth + z[:, None] * (cur - oth)).astype(np.float32)
        if tmode == "bf16-exact theta":
            cur, prop = bf16(cur), bf16(prop)
        pts = np.concatenate([cur, prop])
        tc, simt, ref = s.glm_loglik_probe(pts, 0), s.glm_loglik_probe(pts, 1), s.glm_loglik_probe(pts, 2)
        dd = lambda v: v[pairs:] - v[:pairs]
        e_tc, e_si = np.abs(dd(tc) - dd(ref)), np.abs(dd(simt) - dd(ref))
        print(f"X {xmode:18s} theta {tmode:16s} |dlogp| {np.abs(dd(ref)).mean():9.3f}  TC err mean {e_tc.mean():.2e} max {e_tc.max():.2e}"
              f"   SIMT err mean {e_si.mean():.2e} max {e_si.max():.2e}   level TC {np.abs(tc-ref).mean():.2e} SIMT {np.abs(simt-ref).mean():.2e}", flush=True)
    s.release()
