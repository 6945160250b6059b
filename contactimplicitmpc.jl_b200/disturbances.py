"""Disturbance generators of the simulator (`src/simulator/disturbances.jl:4-84`): `disturbances(d, x, t)` returns the
`w` of simulator step t (1-based) — the `w1` entry of θ that `A_func(q)ᵀ w` injects into the dynamics
(src/dynamics/model.jl:33).  Host-side and tiny; `MonteCarloRollouts.run(..., dist=...)` uploads one (R, nw) row
block per step.  A per-rollout variant is obtained by giving `w` a leading rollout axis."""
from __future__ import annotations

import numpy as np


class Disturbances:
    def __call__(self, t: int) -> np.ndarray:
        raise NotImplementedError


class EmptyDisturbances(Disturbances):
    """RoboDojo `empty_disturbances(model)`: zeros."""

    def __init__(self, nw: int):
        self.w = np.zeros(nw)

    def __call__(self, t):
        return self.w


class OpenLoopDisturbance(Disturbances):
    """`open_loop_disturbances(w, N_sample)` (disturbances.jl:4-39): w[k] is held for N_sample simulator steps and
    divided by N_sample (an impulse per MPC step spread over the simulator steps)."""

    def __init__(self, w, N_sample: int):
        self.w = np.asarray(w, dtype=np.float64)
        self.N = int(N_sample)
        self.idx, self.cnt = 0, self.N

    def __call__(self, t):
        if t == 1:
            self.idx, self.cnt = 0, self.N
        if self.cnt == self.N:
            self.idx += 1
            self.cnt = 0
        self.cnt += 1
        return self.w[self.idx - 1] / self.N


class ImpulseDisturbance(Disturbances):
    """`impulse_disturbances(w, idx)` (disturbances.jl:41-61): w[i] at simulator step idx[i] (1-based), zero otherwise."""

    def __init__(self, w, idx):
        self.w = np.asarray(w, dtype=np.float64)
        self.idx = [int(i) for i in idx]

    def __call__(self, t):
        for i, ti in enumerate(self.idx):
            if ti == t:
                return self.w[i]
        return np.zeros_like(self.w[0])


class RandomDisturbance(Disturbances):
    """`random_disturbances(model, w_amp, H, h)` (disturbances.jl:63-84): H draws `rand(nw) .* w_amp[1]`, looked up with
    `searchsortedlast(d.t, t)` on the time grid (t − 1)·h — note that the reference compares the step INDEX t with times.
    Philox stream (Julia's MersenneTwister stream is not reproducible elsewhere)."""

    def __init__(self, nw: int, w_amp, H: int, h: float, seed: int = 0):
        w_amp = np.atleast_1d(np.asarray(w_amp, dtype=np.float64))
        rng = np.random.Generator(np.random.Philox(seed))
        self.w = rng.random((H, nw)) * w_amp[0]
        self.t = np.arange(H) * h

    def __call__(self, t):
        k = int(np.searchsorted(self.t, t, side="right"))  # searchsortedlast, 1-based
        return self.w[max(k, 1) - 1]
