"""`newton_solve!` and the CIMPC `policy` glue (CPU oracle, numpy fp64, dense KKT + LAPACK solve).

TEST INFRASTRUCTURE ONLY.  Restates
  * `NewtonOptions`, `reset!`, `newton_solve!`                         src/controller/newton.jl:2-11, 130-167, 169-288
  * `NewtonResidualConfiguration[Force]`, `residual!`, `update_traj!`,
    `gradient!` (TrackingObjective)                                    src/controller/newton_residual.jl:3-176, 178-220
  * `NewtonJacobianConfiguration[Force]`, `initialize_jacobian!`,
    `update_jacobian!`, `hessian!` (TrackingObjective)                 src/controller/newton_jacobian.jl:2-198, 200-219
  * `lu_solver` path: `linear_solve!(solver, Δ, R, r)` = dense LU       src/solver/lu.jl:4-12
  * `TrackingObjective`                                                src/controller/objective.jl:3-16
  * `rot_n_stride!`, `rotate!`, `mpc_stride!`, `get_stride`            src/controller/mpc_utils.jl:1-107
  * `reset_window!`, `update_window!`, the `policy` step               src/controller/policy.jl:98-171
The wall-clock `max_time` early exits (newton.jl:181-277) are not restated.
Layout of the Newton vector (newton_residual.jl:69-98): per stage [u1 (nu); (γ1; b1); q2 (nq)],
then the duals ν (nd per stage).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .trajectory import ContactTraj


@dataclass
class NewtonOptions:  # newton.jl:2-11
    r_tol: float = 1.0e-5
    max_iter: int = 10
    beta_init: float = 1.0e-5


@dataclass
class TrackingObjective:
    """`TrackingObjective` (objective.jl:3-16) and, when `v` is given, `TrackingVelocityObjective`
    (objective.jl:18-47).  `q` holds per-stage diagonals (H, nq) or full matrices (H, nq, nq) — e.g.
    `relative_state_cost` (src/dynamics/centroidal_quadruped/model.jl:168-183); `v_target` (H, nq) as in
    objective.jl:34-46, from which `q_target` is accumulated."""
    q: np.ndarray  # (H, nq) or (H, nq, nq)
    u: np.ndarray  # (H, nu)
    gamma: np.ndarray  # (H, nc)
    b: np.ndarray  # (H, nb)
    v: np.ndarray | None = None  # (H, nq) velocity weights
    v_target: np.ndarray | None = None  # (H, nq)

    def __post_init__(self):
        self.q_target = None
        if self.v_target is not None and np.any(self.v_target != 0.0):  # objective.jl:36-44
            qt = np.zeros_like(self.v_target)
            for t in range(1, len(qt)):
                qt[t] = qt[t - 1] + self.v_target[t - 1]
            self.q_target = qt

    def q_mul(self, t, x):
        return self.q[t] @ x if self.q[t].ndim == 2 else self.q[t] * x

    def q_mat(self, t):
        return self.q[t] if self.q[t].ndim == 2 else np.diag(self.q[t])


class NewtonLayout:
    def __init__(self, m, H, mode):
        self.H, self.mode = H, mode
        nq, nu, nc, nb = m.nq, m.nu, m.nc, m.nb
        self.nq, self.nu, self.nc, self.nb = nq, nu, nc, nb
        force = mode == "configurationforce"
        self.nd = nq + nc + nb if force else nq
        self.nr = nq + nu + (nc + nb if force else 0)
        o = 0
        self.iu = np.arange(o, o + nu); o += nu
        if force:
            self.ig = np.arange(o, o + nc); o += nc
            self.ib = np.arange(o, o + nb); o += nb
        else:
            self.ig = self.ib = np.arange(0)
        self.iq = np.arange(o, o + nq); o += nq
        self.iz = np.concatenate([self.iq, self.ig, self.ib])  # IP solution [q2, γ1, b1]
        self.n = H * (self.nr + self.nd)

    def pr(self, t, idx):  # primal indices of stage t (0-based)
        return t * self.nr + idx

    def du(self, t):
        return self.H * self.nr + t * self.nd + np.arange(self.nd)


class Newton:
    """`Newton` (newton.jl:17-91) for one rollout.  `dyn(window, traj)` is the implicit-dynamics backend:
    it returns d (H, nd), dq0 (H, nd, nq), dq1 (H, nd, nq), du1 (H, nd, nu) for stages i = 0..H-1
    (the `d`, `δq0`, `δq1`, `δu1` views of `ImplicitTrajectory`, indexed here by stage, not by knot)."""

    def __init__(self, m, H, h, obj: TrackingObjective, kappa, opts: NewtonOptions, mode="configuration"):
        self.m, self.H, self.h, self.obj, self.kappa, self.opts, self.mode = m, H, h, obj, kappa, opts, mode
        self.lay = NewtonLayout(m, H, mode)
        self.traj = ContactTraj(m, H, h, kappa)
        self.traj_cand = ContactTraj(m, H, h, kappa)
        self.nu_ = np.zeros((H, self.lay.nd))
        self.nu_cand = np.zeros((H, self.lay.nd))
        self.beta = opts.beta_init
        self.res = np.zeros(self.lay.n)
        self.stats = {"ip_sweeps": 0, "iters": 0}

    # reset!  newton.jl:130-167
    def reset(self, ref: ContactTraj, q0, q1, warm_start):
        H = self.H
        self.beta = self.opts.beta_init
        if not warm_start:
            self.nu_[:] = 0.0
            self.nu_cand[:] = 0.0
            for name in ("u", "w", "gamma", "b", "z", "theta"):
                getattr(self.traj, name)[:] = getattr(ref, name)[:H]
            self.traj.q[:] = ref.q[:H + 2]
        self.traj.q[0] = q0
        self.traj.q[1] = q1
        self.traj.update_theta(0)
        self.traj.update_theta(1)
        self._copy(self.traj_cand, self.traj)

    @staticmethod
    def _copy(dst: ContactTraj, src: ContactTraj):
        for name in ("q", "u", "w", "gamma", "b", "z", "theta"):
            getattr(dst, name)[:] = getattr(src, name)

    # residual!  newton_residual.jl:113-138  (+ gradient! :178-220)
    def residual(self, nu, dyn_out, traj: ContactTraj, ref: ContactTraj):
        L, H = self.lay, self.H
        d, dq0, dq1, du1 = dyn_out
        r = np.zeros(L.n)
        for t in range(H):
            qt = 0.0 if self.obj.q_target is None else self.obj.q_target[t]  # newton_residual.jl:229, 262
            r[L.pr(t, L.iq)] += self.obj.q_mul(t, traj.q[t + 2] - (ref.q[t + 2] + qt))
            r[L.pr(t, L.iu)] += self.obj.u[t] * (traj.u[t] - ref.u[t])
            if self.mode == "configurationforce":
                r[L.pr(t, L.ig)] += self.obj.gamma[t] * (traj.gamma[t] - ref.gamma[t])
                r[L.pr(t, L.ib)] += self.obj.b[t] * (traj.b[t] - ref.b[t])
            if self.obj.v is not None:  # gradient!(…, ::TrackingVelocityObjective)  newton_residual.jl:221-281
                dv = self.obj.v[t] * (traj.q[t + 2] - traj.q[t + 1])
                if self.obj.v_target is not None and self.mode == "configuration":
                    dv = dv - self.obj.v[t] * self.obj.v_target[t]  # only the :configuration method has it (:270, :276)
                r[L.pr(t, L.iq)] += dv
                if t >= 1:
                    r[L.pr(t - 1, L.iq)] -= dv
        for i in range(H):
            if i >= 2:
                r[L.pr(i - 2, L.iq)] += dq0[i].T @ nu[i]
            if i >= 1:
                r[L.pr(i - 1, L.iq)] += dq1[i].T @ nu[i]
            r[L.pr(i, L.iu)] += du1[i].T @ nu[i]
            r[L.du(i)] += d[i]
            r[L.pr(i, L.iz)] -= nu[i]
        return r

    # jacobian!  newton_jacobian.jl:148-198  (+ hessian! :200-219)
    def jacobian(self, dyn_out, beta):
        L, H = self.lay, self.H
        _, dq0, dq1, du1 = dyn_out
        R = np.zeros((L.n, L.n))
        for t in range(H):
            R[np.ix_(L.pr(t, L.iq), L.pr(t, L.iq))] += self.obj.q_mat(t)
            R[L.pr(t, L.iu), L.pr(t, L.iu)] += self.obj.u[t]
            if self.mode == "configurationforce":
                R[L.pr(t, L.ig), L.pr(t, L.ig)] += self.obj.gamma[t]
                R[L.pr(t, L.ib), L.pr(t, L.ib)] += self.obj.b[t]
            if self.obj.v is not None:  # hessian!(…, ::TrackingVelocityObjective)  newton_jacobian.jl:221-248
                R[L.pr(t, L.iq), L.pr(t, L.iq)] += self.obj.v[t]
                if t >= 1:
                    R[L.pr(t - 1, L.iq), L.pr(t - 1, L.iq)] += self.obj.v[t]
                    R[L.pr(t - 1, L.iq), L.pr(t, L.iq)] -= self.obj.v[t]
                    R[L.pr(t, L.iq), L.pr(t - 1, L.iq)] -= self.obj.v[t]
            R[L.pr(t, L.iz), L.du(t)] -= 1.0
            R[L.du(t), L.pr(t, L.iz)] -= 1.0
        for i in range(H):
            if i >= 2:
                R[np.ix_(L.du(i), L.pr(i - 2, L.iq))] += dq0[i]
                R[np.ix_(L.pr(i - 2, L.iq), L.du(i))] += dq0[i].T
            if i >= 1:
                R[np.ix_(L.du(i), L.pr(i - 1, L.iq))] += dq1[i]
                R[np.ix_(L.pr(i - 1, L.iq), L.du(i))] += dq1[i].T
            R[np.ix_(L.du(i), L.pr(i, L.iu))] += du1[i]
            R[np.ix_(L.pr(i, L.iu), L.du(i))] += du1[i].T
            # dual regularization: `jac.reg_du .-= β κ` spans ALL dual diagonals, once per stage (App. C.2)
            dd = np.arange(H * L.nr, L.n)
            R[dd, dd] -= beta * self.kappa
        return R

    # update_traj!  newton_residual.jl:140-176
    def update_traj(self, cand: ContactTraj, traj: ContactTraj, nu_cand, nu, delta, alpha):
        L, H = self.lay, self.H
        for t in range(H):
            cand.q[t + 2] = traj.q[t + 2] - alpha * delta[L.pr(t, L.iq)]
            cand.u[t] = traj.u[t] - alpha * delta[L.pr(t, L.iu)]
            if self.mode == "configurationforce":
                cand.gamma[t] = traj.gamma[t] - alpha * delta[L.pr(t, L.ig)]
                cand.b[t] = traj.b[t] - alpha * delta[L.pr(t, L.ib)]
            nu_cand[t] = nu[t] - alpha * delta[L.du(t)]
        cand.update_z()
        cand.update_theta()

    # newton_solve!  newton.jl:169-288
    def solve(self, dyn, q0, q1, window, ref: ContactTraj, warm_start=False):
        o = self.opts
        self.reset(ref, q0, q1, warm_start)
        out = dyn(window, self.traj)
        self.stats["ip_sweeps"] += 1
        self.res = self.residual(self.nu_, out, self.traj, ref)
        r_norm = np.abs(self.res).sum()
        for _l in range(o.max_iter):
            if r_norm / len(self.res) < o.r_tol:
                break
            self.stats["iters"] += 1
            R = self.jacobian(out, self.beta)
            delta = np.linalg.solve(R, self.res)
            alpha, it = 1.0, 0
            self.update_traj(self.traj_cand, self.traj, self.nu_cand, self.nu_, delta, alpha)
            out_c = dyn(window, self.traj_cand)
            self.stats["ip_sweeps"] += 1
            res_c = self.residual(self.nu_cand, out_c, self.traj_cand, ref)
            r_cand = np.abs(res_c).sum()
            while r_cand ** 2 >= (1.0 - 0.001 * alpha) * r_norm ** 2:
                alpha *= 0.5
                it += 1
                if it > 6:
                    break
                self.update_traj(self.traj_cand, self.traj, self.nu_cand, self.nu_, delta, alpha)
                out_c = dyn(window, self.traj_cand)
                self.stats["ip_sweeps"] += 1
                res_c = self.residual(self.nu_cand, out_c, self.traj_cand, ref)
                r_cand = np.abs(res_c).sum()
            # accept (note: with iter > 6 the last α is applied although its candidate was never evaluated;
            # res ← res_cand of the previous trial — the reference's behaviour, newton.jl:271-275)
            self.update_traj(self.traj, self.traj, self.nu_, self.nu_, delta, alpha)
            self.res = res_c
            out = out_c
            r_norm = r_cand
            self.beta = min(self.beta * 1.3, 1.0e2) if it > 6 else max(1.0e1, self.beta / 1.3)
        return self.traj.u[0].copy()


# ---------------------------------------------------------------------------------------------- policy glue
def get_stride(m, traj: ContactTraj):  # mpc_utils.jl:103-107
    s = np.zeros(m.nq)
    s[0] = traj.q[-2][0] - traj.q[0][0]
    return s


def rot_n_stride(traj: ContactTraj, stride):  # mpc_utils.jl:1-101
    H = traj.H
    first = {n: getattr(traj, n)[0].copy() for n in ("q", "u", "w", "gamma", "b", "z", "theta")}
    traj.q[:H + 1] = traj.q[1:H + 2].copy()
    for n in ("u", "w", "gamma", "b", "z", "theta"):
        a = getattr(traj, n)
        a[:H - 1] = a[1:H].copy()
        a[H - 1] = first[n]
    traj.q[H + 1] = first["q"]
    i = traj.idx
    for t in (H, H + 1):  # 0-based positions of Julia's t = H+1, H+2
        traj.q[t] = traj.q[t - H] + stride
        tau = t - 2
        traj.z[tau, i.q2] = traj.q[tau + 2]
        traj.theta[tau, i.q0] = traj.q[tau]
        traj.theta[tau, i.q1] = traj.q[tau + 1]
        tau = t - 3
        traj.theta[tau, i.q0] = traj.q[tau]
        traj.theta[tau, i.q1] = traj.q[tau + 1]


def update_window(window, max_window):  # policy.jl:162-171 (0-based knots)
    return [(w + 1) % max_window for w in window]
