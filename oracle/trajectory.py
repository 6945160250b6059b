"""`ContactTraj`, gait loading and `ImplicitTrajectory` / `implicit_dynamics!` (CPU oracle).

TEST INFRASTRUCTURE ONLY.  Restates
  * `ContactTraj`, `update_z!`, `update_θ!`, `get_trajectory(:split_traj_alt)`   src/controller/trajectory.jl:1-82, 168-179
  * `pack_z` (LinearizedCone)                                                    src/simulation/index.jl:437-441
  * `ImplicitTrajectory` ctor and `implicit_dynamics!`                            src/controller/implicit_dynamics.jl:21-90, 156-192
  * `z_initialize!`                                                               src/simulation/simulation.jl:59-63
"""
from __future__ import annotations

import copy

import numpy as np

from .ip import IPOptions, interior_point_solve
from .linearized import LinProblem, lin_blocks, linearized_step
from .models import Model
from .residual import Index, Residual


class ContactTraj:
    """q: (H+2, nq); u, w, γ, b, z, θ: (H, ·).  trajectory.jl:1-19"""

    def __init__(self, m: Model, H: int, h: float, kappa: float = 0.0):
        self.H, self.h, self.kappa = H, h, kappa
        self.q = np.zeros((H + 2, m.nq))
        self.u = np.zeros((H, m.nu))
        self.w = np.zeros((H, m.nw))
        self.gamma = np.zeros((H, m.nc))
        self.b = np.zeros((H, m.nb))
        self.z = np.zeros((H, m.nz))
        self.theta = np.zeros((H, m.ntheta))
        self.theta[:, -1] = h                      # trajectory.jl:38
        self.idx = Index(m)

    def copy(self):
        return copy.deepcopy(self)

    def update_z(self, t=None):                    # trajectory.jl:51-65
        ts = range(self.H) if t is None else [t]
        i = self.idx
        for t in ts:
            if 0 <= t < self.H:
                self.z[t, i.q2] = self.q[t + 2]
                self.z[t, i.g1] = self.gamma[t]
                self.z[t, i.b1] = self.b[t]

    def update_theta(self, t=None):                # trajectory.jl:67-82
        ts = range(self.H) if t is None else [t]
        i = self.idx
        for t in ts:
            if 0 <= t < self.H:
                self.theta[t, i.q0] = self.q[t]
                self.theta[t, i.q1] = self.q[t + 1]
                self.theta[t, i.u1] = self.u[t]
                self.theta[t, i.w1] = self.w[t]


def phi_numeric(m: Model, q):
    return np.array([float(e) for e in m.phi_func([float(v) for v in q])])


def pack_z(m: Model, q2, g1, b1, psi1, eta1):
    """index.jl:437-441 — note `model.μ_world`, not the gait's μ."""
    s1 = phi_numeric(m, q2)
    s2 = m.mu_world * g1 - b1.reshape(m.nc, m.nf).sum(axis=1)
    return np.concatenate([q2, g1, b1, psi1, s1, eta1, s2])


def trajectory_from_gait(m: Model, gait: dict) -> ContactTraj:
    """get_trajectory(..., load_type = :split_traj_alt)  trajectory.jl:168-179."""
    H = gait["u"].shape[0]
    tr = ContactTraj(m, H, float(gait["h"]))
    tr.q[:] = gait["q"]
    tr.u[:] = gait["u"]
    tr.gamma[:] = gait["gamma"]
    tr.b[:] = gait["b"]
    for t in range(H):
        tr.z[t] = pack_z(m, gait["q"][t + 2], gait["gamma"][t], gait["b"][t], gait["psi"][t], gait["eta"][t])
        tr.theta[t] = np.concatenate([gait["q"][t], gait["q"][t + 1], gait["u"][t], np.zeros(m.nw),
                                      [gait["mu"]], [gait["h"]]])
    return tr


class ImplicitTrajectory:
    """implicit_dynamics.jl:6-90: one linearized IP problem per reference knot."""

    def __init__(self, ref: ContactTraj, res: Residual, kappa: float, opts: IPOptions, mode="configuration"):
        m = res.model
        self.model, self.res, self.opts, self.mode = m, res, opts, mode
        self.H = ref.H
        self.kappa = kappa
        self.nd = m.nq if mode == "configuration" else m.nq + m.nc + m.nb
        self.ncol = 2 * m.nq + m.nu
        self.lin = []
        self.ip = []
        for t in range(ref.H):
            r0, rz0, rth0 = linearized_step(res, ref.z[t], ref.theta[t], kappa)
            blk = lin_blocks(res.idx, m.nc, ref.z[t], ref.theta[t], r0, rz0, rth0)
            self.lin.append((r0, rz0, rth0))
            self.ip.append(LinProblem(blk, res.idx))
        self.z = np.zeros((ref.H, m.nz))
        self.dz = np.zeros((ref.H, m.nz, m.ntheta))
        self.d = np.zeros((ref.H, self.nd))
        self.status = np.zeros(ref.H, dtype=bool)
        self.iters = np.zeros(ref.H, dtype=int)

    def set_altitude(self, alt):                   # implicit_dynamics.jl:141-154
        for p in self.ip:
            p.b.alt = np.asarray(alt, dtype=np.float64).copy()

    # views δq0, δq1, δu1 (implicit_dynamics.jl:82-86)
    def dq0(self, t):
        return self.dz[t, :self.nd, 0:self.model.nq]

    def dq1(self, t):
        return self.dz[t, :self.nd, self.model.nq:2 * self.model.nq]

    def du1(self, t):
        m = self.model
        return self.dz[t, :self.nd, 2 * m.nq:2 * m.nq + m.nu]


def implicit_dynamics(im: ImplicitTrajectory, traj: ContactTraj, window=None):
    """implicit_dynamics!  implicit_dynamics.jl:156-192.  window: 0-based knots, length H+2."""
    m = im.model
    if window is None:
        window = list(range(traj.H + 2))
    stages = window[:-2]
    for i, t in enumerate(stages):
        z = np.ones(m.nz)                          # z_initialize!  simulation.jl:59-63
        z[im.res.idx.q2] = traj.q[i + 2]
        status, zs, dz, iters = interior_point_solve(im.ip[t], z, traj.theta[i], im.opts)
        im.z[t] = zs
        if dz is not None:
            im.dz[t] = dz
        im.status[t], im.iters[t] = status, iters
        d = zs[:im.nd].copy()
        d[:m.nq] -= traj.q[i + 2]
        if im.mode == "configurationforce":
            d[m.nq:m.nq + m.nc] -= traj.gamma[i]
            d[m.nq + m.nc:] -= traj.b[i]
        im.d[t] = d
    return im
