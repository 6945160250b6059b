"""Device `newton_solve!` (one MPC step for a batch of rollouts) vs the numpy oracle (`oracle/newton.py`:
dense KKT + LAPACK LU exactly as the reference's `:lu_solver` path), run with `-m gpu`.

The oracle's implicit-dynamics backend is the C restatement with accurate LU solves and `max_ls = 0`
(the deterministic setting of the parity protocol); the device side uses the same options."""
import numpy as np
import pytest

from common import SIZES, load_gait, load_lin

pytestmark = pytest.mark.gpu

H_MPC = 10
KAPPA = 1.0e-4


def _reference_traj(robot):
    from oracle.models import get_model
    from oracle.trajectory import ContactTraj
    m = get_model(robot)
    lin, gait = load_lin(robot), load_gait(robot)
    ref = ContactTraj(m, lin["z0"].shape[0], gait["h"])
    ref.q[:] = gait["q"]; ref.u[:] = gait["u"]; ref.gamma[:] = gait["gamma"]; ref.b[:] = gait["b"]
    ref.z[:] = lin["z0"]; ref.theta[:] = lin["th0"]
    return m, lin, gait, ref


def _objective(m):
    # examples/quadruped/monte_carlo.jl:33-37
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (m.nq - 3)), (H_MPC, 1))
    ou = np.tile(3e-2 * np.ones(m.nu), (H_MPC, 1))
    return oq, ou


def _oracle_newton(m, lin, gait, ip_kw, n_opts):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    co = COracle(*SIZES["quadruped"], lin, mode="configuration", solver="lu")
    ipo = IPOptions(diff_sol=True, **ip_kw)
    nq = m.nq

    def dyn(window, traj):
        knot = np.array(window[:H_MPC], dtype=np.int32)
        z, dz, st, it = co.solve(knot, traj.theta[:H_MPC], traj.q[2:H_MPC + 2], ipo)
        return z[:, :nq] - traj.q[2:H_MPC + 2], dz[:, :, :nq], dz[:, :, nq:2 * nq], dz[:, :, 2 * nq:]

    oq, ou = _objective(m)
    obj = TrackingObjective(q=oq, u=ou, gamma=np.full((H_MPC, m.nc), 1e-100), b=np.full((H_MPC, m.nb), 1e-100))
    return Newton(m, H_MPC, gait["h"], obj, KAPPA, NewtonOptions(**n_opts)), dyn


def test_newton_solve_matches_oracle(cuda_device):
    import torch
    import cimpc_b200 as cb
    m, lin, gait, ref = _reference_traj("quadruped")
    ip_kw = dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, max_ls=0)  # monte_carlo.jl:51-58, deterministic ls
    n_opts = dict(r_tol=3e-4, max_iter=5)                              # monte_carlo.jl:44-48
    R = 24
    rng = np.random.default_rng(21)
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.01 * rng.standard_normal((R, m.nq))
    q1[0] = ref.q[1]  # one unperturbed rollout
    window = np.arange(H_MPC + 2, dtype=np.int32)

    im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode="configuration", opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    oq, ou = _objective(m)
    nw = cb.Newton(im, H_MPC, R, oq, ou, KAPPA, cb.NewtonOptions(**n_opts))
    u, q, info = nw.solve(window, ref.q[:H_MPC + 2], ref.u[:H_MPC], gait["mu"], gait["h"],
                          torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1).to(cuda_device), want_q=True)
    torch.cuda.synchronize()
    u, q, info = u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()

    worst_u = worst_q = 0.0
    agree = 0
    for r in range(R):
        core, dyn = _oracle_newton(m, lin, gait, ip_kw, n_opts)
        uo = core.solve(dyn, q0[r], q1[r], list(window), ref, warm_start=False)
        same_path = (core.stats["iters"] == info[r, 0]) and (core.stats["ip_sweeps"] == info[r, 1])
        if same_path:
            agree += 1
            worst_u = max(worst_u, np.abs(u[r] - uo).max())
            worst_q = max(worst_q, np.abs(q[r] - core.traj.q).max())
        conv = np.abs(core.res).sum() / len(core.res) < n_opts["r_tol"]
        assert bool(info[r, 2]) == bool(conv)
    assert agree >= R - 1, f"only {agree}/{R} rollouts followed the oracle's iteration path"
    assert worst_u <= 1e-7 and worst_q <= 1e-7, (worst_u, worst_q)


def test_newton_warm_start_and_mpc_loop(cuda_device):
    """Three consecutive MPC steps with `rot_n_stride!` / `update_window!` between them (policy.jl:119-131),
    warm-started after the first, against the oracle driven the same way."""
    import torch
    import cimpc_b200 as cb
    from oracle.newton import get_stride, rot_n_stride, update_window
    m, lin, gait, ref = _reference_traj("quadruped")
    ip_kw = dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, max_ls=0)
    n_opts = dict(r_tol=3e-4, max_iter=5)
    R = 6
    rng = np.random.default_rng(22)
    H_ref = ref.H
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode="configuration", opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    oq, ou = _objective(m)
    nw = cb.Newton(im, H_MPC, R, oq, ou, KAPPA, cb.NewtonOptions(**n_opts))
    cores = [_oracle_newton(m, lin, gait, ip_kw, n_opts) for _ in range(R)]
    ptraj = ref.copy()
    stride = get_stride(m, ref)
    window = list(range(H_MPC + 2))
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.005 * rng.standard_normal((R, m.nq))
    for step in range(3):
        u, q, info = nw.solve(np.array(window, dtype=np.int32), ptraj.q[:H_MPC + 2], ptraj.u[:H_MPC], gait["mu"],
                              gait["h"], torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1).to(cuda_device),
                              warm_start=step > 0, want_q=True)
        torch.cuda.synchronize()
        u, q, info = u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()
        q1_next = np.zeros_like(q1)
        for r, (core, dyn) in enumerate(cores):
            it0, sw0 = core.stats["iters"], core.stats["ip_sweeps"]
            uo = core.solve(dyn, q0[r], q1[r], window, ptraj, warm_start=step > 0)
            if core.stats["iters"] - it0 == info[r, 0] and core.stats["ip_sweeps"] - sw0 == info[r, 1]:
                assert np.abs(u[r] - uo).max() <= 1e-6 and np.abs(q[r] - core.traj.q).max() <= 1e-6
            q1_next[r] = core.traj.q[2]  # "measured" next configuration = the oracle's plan (same for both)
        rot_n_stride(ptraj, stride)
        window = update_window(window, H_ref)
        q0, q1 = q1, q1_next
