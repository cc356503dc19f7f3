"""Mirror of ``uncomplicate.bayadera.mcmc`` (/root/reference/src/clojure/uncomplicate/bayadera/mcmc.clj:15-101):
arity defaults, cooling schedules and the ``mix!`` warm-up tuner.  Engine-agnostic host logic — it works on any
object with the MCMC protocol methods (the CUDA ``B200Stretch`` in production; the CPU oracle in tests).
"""
from __future__ import annotations

import math
import os
import struct
from typing import Callable, Optional


def generate_seed() -> int:
    """uncomplicate.commons.utils/generate-seed: a random int32."""
    return struct.unpack("i", os.urandom(4))[0]


def _dimension(samp) -> int:
    return int(samp.model.dimension)


def init_position(samp, seed_or_position, limits=None):
    return samp.init_position(seed_or_position, limits)


def acc_rate(samp, a: float = 2.0) -> float:
    return float(samp.acc_rate(a))


def burn_in(samp, steps: Optional[int] = None, a: float = 2.0):
    """mcmc.clj:27-33 — default 256 * dimension steps at a = 2.0."""
    steps = 256 * _dimension(samp) if steps is None else steps
    return samp.burn_in(int(steps), a)


def run_sampler(samp, steps: Optional[int] = None, a: float = 2.0):
    """mcmc.clj:35-41 — default 64 * dimension steps at a = 2.0."""
    steps = 64 * _dimension(samp) if steps is None else steps
    return samp.run_sampler(int(steps), a)


# cooling schedules, mcmc.clj:43-54: (schedule steps) -> fn i -> temperature
def sqrt_n(temp: float) -> Callable[[int], float]:
    return lambda i: math.sqrt(temp - i)


def pow_n(power: float) -> Callable[[float], Callable[[int], float]]:
    def schedule(temp: float) -> Callable[[int], float]:
        return lambda i: math.pow(temp - i, power)
    schedule.pow_n_power = power          # lets the engine run this schedule inside bay_mix
    return schedule


def minus_n(temp: float) -> Callable[[int], float]:
    return lambda i: temp - i


def anneal(samp, steps: Optional[int] = None, a: float = 2.0, schedule=minus_n):
    """mcmc.clj:56-64 — ``(p/anneal! samp (schedule steps) steps a)``; default 256 * dimension steps."""
    steps = 256 * _dimension(samp) if steps is None else steps
    steps = int(steps)
    return samp.anneal(schedule(float(steps)), steps, a)


def mix(samp, options: Optional[dict] = None) -> dict:
    """``mix!`` (mcmc.clj:66-101): anneal, tune ``a`` towards an acceptance rate in [min, max], burn in."""
    o = dict(options or {})
    step = int(o.get("step", 64))
    dimension_power = float(o.get("dimension-power", 0.8))
    schedule = o.get("cooling-schedule", minus_n)
    a = float(o.get("a", 2.0))
    min_acc = float(o.get("min-acc-rate", 0.2))
    max_acc = float(o.get("max-acc-rate", 0.5))
    target = (max_acc + min_acc) / 2.0
    dim = _dimension(samp)
    n = int(step * math.pow(dim, dimension_power))
    anneal(samp, n, a, schedule)
    i = 0
    while True:
        rate = acc_rate(samp, a)
        if step < i:
            break
        if rate < min_acc:
            a = 1.0 + (a - 1.0) * (rate / target)
        elif max_acc < rate:
            a = a * (rate / target)
        else:
            break
        i += 1
    burn_in(samp, n, a)
    burn_in(samp, step, 2.0)
    return {"a": a, "acc-rate": acc_rate(samp, a), "acc-rate-2.0": acc_rate(samp)}
