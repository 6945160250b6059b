"""The linear algebra behind the general device Newton kernel (`csrc/newton_general.cuh`), pinned on the CPU:
the augmented dual Schur complement — zero-weight force rows eliminated, velocity cost written with slacks and
multipliers — yields the same Newton direction as a dense solve of the reference's KKT matrix
`R = [Q Cᵀ; C −ρI]` (`jacobian!`, src/controller/newton_jacobian.jl:148-198, assembled by oracle/newton.py) for both modes
of `ImplicitTrajectory` and both objectives.  The kernel implements exactly this elimination with a banded block
Cholesky; tests/test_gpu_newton.py then compares the kernel with the oracle end to end."""
import numpy as np
import pytest

from common import SIZES, load_gait, load_lin


def structured_solve(L, H, nq, nu, nyd, obj_q, obj_u, obj_v, dq0, dq1, du1, res, rho):
    vel = obj_v is not None
    yidx = np.concatenate([L.ig, L.ib]).astype(int)
    r_u = np.array([res[L.pr(t, L.iu)] for t in range(H)])
    r_x = np.array([res[L.pr(t, L.iq)] for t in range(H)])
    r_nu = np.array([res[L.du(t)] for t in range(H)])
    Wu = 1.0 / obj_u
    dense_q = obj_q.ndim == 3  # per-stage full matrices (relative_state_cost): P⁻¹ is block diagonal with Q_t⁻¹
    A, B, U = dq1[:, :nq, :], dq0[:, :nq, :], du1[:, :nq, :]
    dnu_y = np.zeros((H, nyd))
    if nyd:  # zero-weight (γ, b) rows: Δν^y = −r_y, its coupling moves to the right-hand side
        dnu_y = -np.array([res[L.pr(t, yidx)] for t in range(H)])
        Ay, By, Uy = dq1[:, nq:, :], dq0[:, nq:, :], du1[:, nq:, :]
        for t in range(H):
            r_u[t] -= Uy[t].T @ dnu_y[t]
            if t >= 1:
                r_x[t - 1] -= Ay[t].T @ dnu_y[t]
            if t >= 2:
                r_x[t - 2] -= By[t].T @ dnu_y[t]
    nb, npr = (2 * nq if vel else nq), nu + nq
    n = H * nb
    C = np.zeros((n, H * npr))
    for t in range(H):
        r0 = t * nb
        C[r0:r0 + nq, t * npr:t * npr + nu] = U[t]
        C[r0:r0 + nq, t * npr + nu:(t + 1) * npr] = -np.eye(nq)
        if t >= 1:
            C[r0:r0 + nq, (t - 1) * npr + nu:t * npr] = A[t]
        if t >= 2:
            C[r0:r0 + nq, (t - 2) * npr + nu:(t - 1) * npr] = B[t]
        if vel:  # slack rows  x_t − x_{t−1} − s_t = 0
            C[r0 + nq:r0 + nb, t * npr + nu:(t + 1) * npr] = np.eye(nq)
            if t >= 1:
                C[r0 + nq:r0 + nb, (t - 1) * npr + nu:t * npr] = -np.eye(nq)
    W = np.zeros((H * npr, H * npr))
    for t in range(H):
        W[t * npr:t * npr + nu, t * npr:t * npr + nu] = np.diag(Wu[t])
        W[t * npr + nu:(t + 1) * npr, t * npr + nu:(t + 1) * npr] = np.linalg.inv(obj_q[t]) if dense_q else np.diag(1.0 / obj_q[t])
    reg = np.zeros(n)
    for t in range(H):
        reg[t * nb:t * nb + nq] = rho
        if vel:
            reg[t * nb + nq:(t + 1) * nb] = 1.0 / obj_v[t]
    Y = C @ W @ C.T + np.diag(reg)
    # block-pentadiagonal: nothing outside the band of three stage blocks
    for t1 in range(H):
        for t2 in range(H):
            if abs(t1 - t2) > 2:
                assert not Y[t1 * nb:(t1 + 1) * nb, t2 * nb:(t2 + 1) * nb].any()
    rp = np.concatenate([np.concatenate([r_u[t], r_x[t]]) for t in range(H)])
    rd = np.zeros(n)
    for t in range(H):
        rd[t * nb:t * nb + nq] = r_nu[t][:nq]
    dmu = np.linalg.solve(Y, C @ (W @ rp) - rd)
    np.linalg.cholesky(Y)  # SPD: the kernel factorises it without pivoting
    dp = W @ (rp - C.T @ dmu)
    delta = np.zeros(L.n)
    for t in range(H):
        delta[L.pr(t, L.iu)] = dp[t * npr:t * npr + nu]
        delta[L.pr(t, L.iq)] = dp[t * npr + nu:(t + 1) * npr]
        delta[L.du(t)[:nq]] = dmu[t * nb:t * nb + nq]
    if nyd:
        for t in range(H):
            gy = Uy[t] @ delta[L.pr(t, L.iu)]
            if t >= 1:
                gy += Ay[t] @ delta[L.pr(t - 1, L.iq)]
            if t >= 2:
                gy += By[t] @ delta[L.pr(t - 2, L.iq)]
            delta[L.pr(t, yidx)] = gy - rho * dnu_y[t] - r_nu[t][nq:]
            delta[L.du(t)[nq:]] = dnu_y[t]
    return delta


@pytest.mark.parametrize("robot,mode,vel,dense", [("quadruped", "configuration", False, False), ("quadruped", "configuration", True, False),
                                                  ("flamingo", "configurationforce", False, False), ("flamingo", "configurationforce", True, False),
                                                  ("quadruped", "configuration", False, True), ("quadruped", "configuration", True, True)])
def test_augmented_dual_schur_equals_dense_kkt(robot, mode, vel, dense):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.models import get_model
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    from oracle.trajectory import ContactTraj
    m = get_model(robot)
    lin, gait = load_lin(robot), load_gait(robot)
    ref = ContactTraj(m, lin["z0"].shape[0], gait["h"])
    ref.q[:] = gait["q"]; ref.u[:] = gait["u"]; ref.gamma[:] = gait["gamma"]; ref.b[:] = gait["b"]
    ref.z[:] = lin["z0"]; ref.theta[:] = lin["th0"]
    H, kappa = (10 if robot == "quadruped" else 15), 2e-4
    nq, nc, nb_, nu = m.nq, m.nc, m.nb, m.nu
    nyd = nc + nb_ if mode == "configurationforce" else 0
    co = COracle(*SIZES[robot], lin, mode=mode, solver="lu")
    ipo = IPOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)

    def dyn(window, traj):
        knot = np.array(window[:H], dtype=np.int32)
        z, dz, st, _ = co.solve(knot, traj.theta[:H], traj.q[2:H + 2], ipo)
        d = z[:, :nq + nyd].copy()
        d[:, :nq] -= traj.q[2:H + 2]
        if nyd:
            d[:, nq:nq + nc] -= traj.gamma[:H]
            d[:, nq + nc:] -= traj.b[:H]
        return d, dz[:, :, :nq], dz[:, :, nq:2 * nq], dz[:, :, 2 * nq:]

    rng = np.random.default_rng(0)
    obj = TrackingObjective(q=np.tile(0.1 * (0.5 + rng.random(nq)), (H, 1)), u=np.tile(0.3 * (0.5 + rng.random(nu)), (H, 1)),
                            gamma=np.full((H, nc), 1e-100), b=np.full((H, nb_), 1e-100),
                            v=np.tile(1e-3 * (1 + 100 * rng.random(nq)), (H, 1)) if vel else None)
    if dense:  # a full SPD weight per stage and a velocity target (objective.jl:34-46)
        Qs = []
        for t in range(H):
            Mq = rng.standard_normal((nq, nq))
            Qs.append(0.05 * (Mq @ Mq.T / nq + np.diag(0.5 + rng.random(nq))))
        obj.q = np.array(Qs)
        if vel:
            obj.v_target = 0.01 * rng.standard_normal((H, nq))
            obj.__post_init__()
    nw = Newton(m, H, gait["h"], obj, kappa, NewtonOptions(r_tol=3e-4, max_iter=5), mode=mode)
    nw.reset(ref, ref.q[0], ref.q[1] + 0.01 * rng.standard_normal(nq), False)
    nw.nu_[:] = 0.01 * rng.standard_normal(nw.nu_.shape)  # non-trivial duals, force duals included
    out = dyn(list(range(H + 2)), nw.traj)
    r = nw.residual(nw.nu_, out, nw.traj, ref)
    for beta in (1e-5, 10.0):  # first Newton iteration and the later ones
        d_ref = np.linalg.solve(nw.jacobian(out, beta), r)
        d_st = structured_solve(nw.lay, H, nq, nu, nyd, obj.q, obj.u, obj.v, out[1], out[2], out[3], r, H * beta * kappa)
        assert np.abs(d_st - d_ref).max() <= 1e-9 * max(1.0, np.abs(d_ref).max())
