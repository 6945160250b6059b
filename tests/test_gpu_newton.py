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


def _oracle_newton(m, lin, gait, ip_kw, n_opts, alt=None):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    co = COracle(*SIZES["quadruped"], lin, mode="configuration", solver="lu")
    ipo = IPOptions(diff_sol=True, **ip_kw)
    nq = m.nq

    def dyn(window, traj):
        knot = np.array(window[:H_MPC], dtype=np.int32)
        # `set_altitude!`: the same offsets on every stage of the policy (implicit_dynamics.jl:141-154)
        z, dz, st, it = co.solve(knot, traj.theta[:H_MPC], traj.q[2:H_MPC + 2], ipo,
                                 alt=None if alt is None else np.tile(alt, (H_MPC, 1)))
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


def test_newton_solve_with_altitude_matches_oracle(cuda_device):
    """`policy` with `altitude_update` (policy.jl:111-115): every rollout carries its own altitude vector, applied to the
    impact rows of all H_mpc subproblems; device `newton_solve!` vs the oracle with the same offsets — and through the
    HOST entry point (`cimpc_newton_solve_batch_host`), which must return the very same controls."""
    import torch
    import cimpc_b200 as cb
    m, lin, gait, ref = _reference_traj("quadruped")
    ip_kw = dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, max_ls=0)
    n_opts = dict(r_tol=3e-4, max_iter=5)
    R = 10
    rng = np.random.default_rng(31)
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.005 * rng.standard_normal((R, m.nq))
    alt = 0.01 * rng.random((R, m.nc))          # terrain up to 1 cm above the reference's ground
    alt[0] = 0.0
    window = np.arange(H_MPC + 2, dtype=np.int32)
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode="configuration", opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    oq, ou = _objective(m)
    nw = cb.Newton(im, H_MPC, R, oq, ou, KAPPA, cb.NewtonOptions(**n_opts))
    u, q, info = nw.solve(window, ref.q[:H_MPC + 2], ref.u[:H_MPC], gait["mu"], gait["h"],
                          torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1).to(cuda_device), want_q=True,
                          alt=torch.from_numpy(alt).to(cuda_device))
    torch.cuda.synchronize()
    u, q, info = u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()
    u_flat, _, _ = nw.solve(window, ref.q[:H_MPC + 2], ref.u[:H_MPC], gait["mu"], gait["h"],
                            torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1).to(cuda_device))
    u_flat = u_flat.cpu().numpy()
    assert np.array_equal(u_flat[0], u[0]) and np.abs(u_flat[1:] - u[1:]).max() > 1e-4  # the offsets matter
    agree = 0
    for r in range(R):
        core, dyn = _oracle_newton(m, lin, gait, ip_kw, n_opts, alt=alt[r])
        uo = core.solve(dyn, q0[r], q1[r], list(window), ref, warm_start=False)
        if core.stats["iters"] == info[r, 0] and core.stats["ip_sweeps"] == info[r, 1]:
            agree += 1
            assert np.abs(u[r] - uo).max() <= 1e-7 and np.abs(q[r] - core.traj.q).max() <= 1e-7
    assert agree >= R - 1
    uh, ih = nw.solve_host(window, ref.q[:H_MPC + 2], ref.u[:H_MPC], gait["mu"], gait["h"], q0, q1, alt=alt)
    assert np.array_equal(uh, u) and np.array_equal(ih, info)


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


# ------------------------------------------------------------------------------------------------------------
# General device Newton (newton_general.cuh): :configurationforce and / or TrackingVelocityObjective
# ------------------------------------------------------------------------------------------------------------
def _general_case(robot, mode, vel):
    """Objective of test/controller/mpc_flamingo.jl:27-41 (flamingo) / monte_carlo.jl:33-37 (quadruped) with optional
    velocity weights; returns everything both sides need."""
    from oracle.newton import TrackingObjective
    m, lin, gait, ref = _reference_traj(robot)
    H = 15 if robot == "flamingo" else H_MPC
    if robot == "flamingo":
        oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H, 1))
        ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H, 1))
        ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H, 1))
    else:
        oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (m.nq - 3)), (H, 1))
        ou = np.tile(3e-2 * np.ones(m.nu), (H, 1))
        ov = np.tile(1e-3 * np.array([1.0, 1.0, 1e2] + [1.0] * (m.nq - 3)), (H, 1))
    og, ob = np.full((H, m.nc), 1e-100), np.full((H, m.nb), 1e-100)
    obj = TrackingObjective(q=oq, u=ou, gamma=og, b=ob, v=ov if vel else None)
    return m, lin, gait, ref, H, obj


def _oracle_newton_general(robot, mode, m, lin, gait, H, obj, kappa, ip_kw, n_opts):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.newton import Newton, NewtonOptions
    co = COracle(*SIZES[robot], lin, mode=mode, solver="lu")
    ipo = IPOptions(diff_sol=True, **ip_kw)
    nq, nc, nb = m.nq, m.nc, m.nb
    force = mode == "configurationforce"

    def dyn(window, traj):
        knot = np.array(window[:H], dtype=np.int32)
        z, dz, st, it = co.solve(knot, traj.theta[:H], traj.q[2:H + 2], ipo)
        nd = nq + (nc + nb if force else 0)
        d = z[:, :nd].copy()
        d[:, :nq] -= traj.q[2:H + 2]
        if force:
            d[:, nq:nq + nc] -= traj.gamma[:H]
            d[:, nq + nc:] -= traj.b[:H]
        return d, dz[:, :, :nq], dz[:, :, nq:2 * nq], dz[:, :, 2 * nq:]

    return Newton(m, H, gait["h"], obj, kappa, NewtonOptions(**n_opts), mode=mode), dyn


@pytest.mark.parametrize("robot,mode,vel", [("quadruped", "configuration", True),
                                            ("flamingo", "configurationforce", False),
                                            ("flamingo", "configurationforce", True),
                                            ("flamingo", "configuration", True),
                                            ("centroidal_quadruped", "configuration", True)])
def test_general_newton_matches_oracle(cuda_device, robot, mode, vel):
    """`newton_solve!` with γ, b as Newton variables (:configurationforce) and / or a TrackingVelocityObjective — the
    flamingo policy of test/controller/mpc_flamingo.jl:27-56 — against the numpy oracle (dense KKT + LAPACK)."""
    import torch
    import cimpc_b200 as cb
    m, lin, gait, ref, H, obj = _general_case(robot, mode, vel)
    kappa = 2.0e-4
    ip_kw = dict(r_tol=1e-8, kappa_tol=kappa, max_iter=100, max_ls=0, undercut=5.0)  # mpc_flamingo.jl:48-54, deterministic ls
    n_opts = dict(r_tol=3e-4, max_iter=5)
    R = 12
    rng = np.random.default_rng(31)
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.005 * rng.standard_normal((R, m.nq))
    q1[0] = ref.q[1]
    window = np.arange(H + 2, dtype=np.int32)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                               opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    nw = cb.Newton(im, H, R, obj.q, obj.u, kappa, cb.NewtonOptions(**n_opts), obj_gamma=obj.gamma, obj_b=obj.b, obj_v=obj.v)
    out = nw.solve(window, ref.q[:H + 2], ref.u[:H], gait["mu"], gait["h"], torch.from_numpy(q0).to(cuda_device),
                   torch.from_numpy(q1).to(cuda_device), want_q=True, ref_gamma=ref.gamma[:H], ref_b=ref.b[:H], want_y=True)
    torch.cuda.synchronize()
    u, q, info, y = (a.cpu().numpy() if a is not None else None for a in out)
    agree, worst = 0, 0.0
    for r in range(R):
        core, dyn = _oracle_newton_general(robot, mode, m, lin, gait, H, obj, kappa, ip_kw, n_opts)
        uo = core.solve(dyn, q0[r], q1[r], list(window), ref, warm_start=False)
        if core.stats["iters"] == info[r, 0] and core.stats["ip_sweeps"] == info[r, 1]:
            agree += 1
            worst = max(worst, np.abs(u[r] - uo).max() / max(1.0, np.abs(uo).max()), np.abs(q[r] - core.traj.q).max())
            if y is not None:
                yo = np.concatenate([core.traj.gamma, core.traj.b], axis=1)
                worst = max(worst, np.abs(y[r] - yo).max() / max(1.0, np.abs(yo).max()))
        conv = np.abs(core.res).sum() / len(core.res) < n_opts["r_tol"]
        assert bool(info[r, 2]) == bool(conv)
    assert agree >= R - 1, f"only {agree}/{R} rollouts followed the oracle's iteration path"
    # quadruped: 1e-7-level agreement; flamingo's Schur complement is ill-conditioned (cond(R) ≈ 5e5, DESIGN.md §5: two
    # faithful CPU restatements of its IP solve differ by 2e-5), and five Newton iterations compound it (observed 3e-6)
    assert worst <= (2e-5 if robot == "flamingo" else 1e-6), worst


def relative_state_cost(qbody, qorientation, qfoot):
    """src/dynamics/centroidal_quadruped/model.jl:168-183, restated: ½ bodyᵀQb body + ½ oriᵀQo ori + Σ ½ (foot − body)ᵀ Qf (foot − body)."""
    Q = np.zeros((18, 18))
    Q[0:3, 0:3] = np.diag(qbody)
    Q[3:6, 3:6] = np.diag(qorientation)
    for i in range(1, 5):
        f = slice(3 + 3 * i, 6 + 3 * i)
        Q[0:3, 0:3] += np.diag(qfoot)
        Q[f, f] += np.diag(qfoot)
        Q[0:3, f] += -np.diag(qfoot)
        Q[f, 0:3] += -np.diag(qfoot)
    return Q


@pytest.mark.parametrize("case", ["flat_trot", "flat_trot_v_target", "dense_no_velocity", "diagonal_as_dense"])
def test_dense_weight_newton_matches_oracle(cuda_device, case):
    """The REAL objective of examples/centroidal_quadruped/flat_trot.jl:37-42 (BASELINE config 5's robot):
    `TrackingVelocityObjective` with q = relative_state_cost(1e-0·[1e-2, 1e-2, 1], 3e-1·[1, 1, 1], 1e-0·[0.2, 0.2, 1]) — a
    NON-diagonal matrix coupling the body position with every foot —, v = 1e-3·[1 1 1; 1e3·[1 1 1]; 1 …], u = 3e-3, and
    v_target (zero in the example since v0 = 0; a non-zero target is exercised as well), H_mpc = 10, κ_mpc = 2e-4, IP
    r_tol 1e-4 / κ_tol κ_mpc / undercut 5, Newton r_tol 3e-5 / max_iter 5 — device dense-weight kernel vs the oracle's
    dense KKT solve.  `diagonal_as_dense` feeds a diagonal objective through the dense kernel and must reproduce the
    general (diagonal) kernel."""
    import torch
    import cimpc_b200 as cb
    from oracle.newton import TrackingObjective
    robot, mode, H, kappa = "centroidal_quadruped", "configuration", 10, 2.0e-4
    m, lin, gait, ref = _reference_traj(robot)
    nq = m.nq
    Q = relative_state_cost(1e-0 * np.array([1e-2, 1e-2, 1]), 3e-1 * np.array([1.0, 1, 1]), 1e-0 * np.array([0.2, 0.2, 1]))
    ov = np.tile(1e-3 * np.array([1.0, 1, 1, 1e3, 1e3, 1e3] + [1.0] * 12), (H, 1))
    ou = np.tile(3e-3 * np.ones(m.nu), (H, 1))
    vt = None
    if case == "flat_trot_v_target":
        v0 = 0.05 * gait["h"]  # flat_trot.jl:42 passes 1/h·[v0 …]-shaped targets; any (H, nq) array is legal
        vt = np.tile(np.array([v0, 0, 0, 0, 0, 0] + [v0, 0, 0] * 4), (H, 1))
    if case == "dense_no_velocity":
        ov = None
    if case == "diagonal_as_dense":
        Q = np.diag(1e-1 * (0.5 + np.random.default_rng(2).random(nq)))
    oq = np.tile(Q, (H, 1, 1))
    og, ob = np.full((H, m.nc), 1e-100), np.full((H, m.nb), 1e-100)
    obj = TrackingObjective(q=oq, u=ou, gamma=og, b=ob, v=ov, v_target=vt)
    ip_kw = dict(r_tol=1e-4, kappa_tol=kappa, max_iter=100, max_ls=0, undercut=5.0)
    n_opts = dict(r_tol=3e-5, max_iter=5)
    R = 8
    rng = np.random.default_rng(77)
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.003 * rng.standard_normal((R, nq))
    q1[0] = ref.q[1]
    window = np.arange(H + 2, dtype=np.int32)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                               opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    nw = cb.Newton(im, H, R, oq, ou, kappa, cb.NewtonOptions(**n_opts), obj_v=ov, v_target=vt)
    mu = float(lin["th0"][0, -2])
    u, q, info = nw.solve(window, ref.q[:H + 2], ref.u[:H], mu, gait["h"], torch.from_numpy(q0).to(cuda_device),
                          torch.from_numpy(q1).to(cuda_device), want_q=True)
    torch.cuda.synchronize()
    u, q, info = u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()
    assert info[:, 0].max() >= 1, "the test must exercise at least one Newton iteration"
    agree, worst = 0, 0.0
    for r in range(R):
        core, dyn = _oracle_newton_general(robot, mode, m, lin, gait, H, obj, kappa, ip_kw, n_opts)
        uo = core.solve(dyn, q0[r], q1[r], list(window), ref, warm_start=False)
        if core.stats["iters"] == info[r, 0] and core.stats["ip_sweeps"] == info[r, 1]:
            agree += 1
            worst = max(worst, np.abs(u[r] - uo).max() / max(1.0, np.abs(uo).max()), np.abs(q[r] - core.traj.q).max())
        conv = np.abs(core.res).sum() / len(core.res) < n_opts["r_tol"]
        assert bool(info[r, 2]) == bool(conv)
    assert agree >= R - 1, f"only {agree}/{R} rollouts followed the oracle's iteration path"
    assert worst <= 1e-6, worst
    if case == "diagonal_as_dense":
        dq = np.tile(np.diag(Q), (H, 1))
        im2 = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                                    opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
        nw2 = cb.Newton(im2, H, R, dq, ou, kappa, cb.NewtonOptions(**n_opts), obj_v=ov)
        u2, q2, info2 = nw2.solve(window, ref.q[:H + 2], ref.u[:H], mu, gait["h"], torch.from_numpy(q0).to(cuda_device),
                                  torch.from_numpy(q1).to(cuda_device), want_q=True)
        same = (info2.cpu().numpy()[:, :2] == info[:, :2]).all(axis=1)
        assert same.mean() >= 0.85
        assert np.abs(u2.cpu().numpy() - u)[same].max() <= 1e-8 and np.abs(q2.cpu().numpy() - q)[same].max() <= 1e-8
    # a matrix that is not positive definite is refused
    bad = oq.copy()
    bad[3] = -bad[3]
    with pytest.raises(cb.CimpcError):
        cb.Newton(im, H, R, bad, ou, kappa, cb.NewtonOptions(**n_opts), obj_v=ov)


def test_general_kernel_equals_specialised_kernel(cuda_device):
    """(:configuration, TrackingObjective) through the GENERAL kernel — selected by passing velocity weights of 1e-300,
    which change nothing numerically — reproduces the specialised kernel: same iteration / sweep counts, u and q to
    round-off."""
    import torch
    import cimpc_b200 as cb
    m, lin, gait, ref = _reference_traj("quadruped")
    ip_kw = dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, max_ls=0)
    R = 64
    rng = np.random.default_rng(5)
    q0 = torch.from_numpy(np.tile(ref.q[0], (R, 1))).to(cuda_device)
    q1 = torch.from_numpy(ref.q[1] + 0.01 * rng.standard_normal((R, m.nq))).to(cuda_device)
    window = np.arange(H_MPC + 2, dtype=np.int32)
    oq, ou = _objective(m)
    res = []
    for tiny_v in (None, np.full((H_MPC, m.nq), 1e-300)):  # a velocity weight of 1e-300 changes nothing numerically
        im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                   mode="configuration", opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
        nw = cb.Newton(im, H_MPC, R, oq, ou, KAPPA, cb.NewtonOptions(r_tol=3e-4, max_iter=5), obj_v=tiny_v)
        u, q, info = nw.solve(window, ref.q[:H_MPC + 2], ref.u[:H_MPC], gait["mu"], gait["h"], q0, q1, want_q=True)
        torch.cuda.synchronize()
        res.append((u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()))
    same = (res[0][2][:, :2] == res[1][2][:, :2]).all(axis=1)
    assert same.mean() >= 0.95
    assert np.abs(res[0][0][same] - res[1][0][same]).max() <= 1e-8
    assert np.abs(res[0][1][same] - res[1][1][same]).max() <= 1e-8


@pytest.mark.parametrize("robot,H,kappa", [("hopper_2D", 10, 2.0e-4), ("centroidal_quadruped", 20, 2.0e-4)])
def test_newton_other_robots_match_oracle(cuda_device, robot, H, kappa):
    """The specialised device Newton on the other two BASELINE robots (hopper: examples/hopper/flat.jl:24-51, H_mpc = 10;
    centroidal quadruped: examples/centroidal_quadruped/flat_trot.jl:31-62, H_mpc = 20) against the numpy oracle."""
    import torch
    import cimpc_b200 as cb
    from oracle.newton import TrackingObjective
    m, lin, gait, ref = _reference_traj(robot)
    rng = np.random.default_rng(41)
    oq = np.tile(1e-1 * (0.5 + rng.random(m.nq)), (H, 1))
    ou = np.tile(1e-1 * (0.5 + rng.random(m.nu)), (H, 1))
    obj = TrackingObjective(q=oq, u=ou, gamma=np.full((H, m.nc), 1e-100), b=np.full((H, m.nb), 1e-100))
    ip_kw = dict(r_tol=1e-8, kappa_tol=kappa, max_iter=100, max_ls=0, undercut=5.0)
    n_opts = dict(r_tol=3e-4, max_iter=5)
    R = 6
    q0 = np.tile(ref.q[0], (R, 1))
    q1 = ref.q[1] + 0.003 * rng.standard_normal((R, m.nq))
    window = np.arange(H + 2, dtype=np.int32)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration",
                               opts=cb.InteriorPointOptions(diff_sol=True, **ip_kw))
    nw = cb.Newton(im, H, R, oq, ou, kappa, cb.NewtonOptions(**n_opts))
    mu = float(lin["th0"][0, -2])
    u, q, info = nw.solve(window, ref.q[:H + 2], ref.u[:H], mu, gait["h"], torch.from_numpy(q0).to(cuda_device),
                          torch.from_numpy(q1).to(cuda_device), want_q=True)
    torch.cuda.synchronize()
    u, q, info = u.cpu().numpy(), q.cpu().numpy(), info.cpu().numpy()
    agree, worst = 0, 0.0
    for r in range(R):
        core, dyn = _oracle_newton_general(robot, "configuration", m, lin, gait, H, obj, kappa, ip_kw, n_opts)
        uo = core.solve(dyn, q0[r], q1[r], list(window), ref, warm_start=False)
        if core.stats["iters"] == info[r, 0] and core.stats["ip_sweeps"] == info[r, 1]:
            agree += 1
            worst = max(worst, np.abs(u[r] - uo).max() / max(1.0, np.abs(uo).max()), np.abs(q[r] - core.traj.q).max())
    assert agree >= R - 1, f"only {agree}/{R} rollouts followed the oracle's iteration path"
    assert worst <= 1e-6, worst


def test_newton_entry_points_reject_bad_arguments(cuda_device):
    """Error behaviour at the ABI: nothing throws inside the library, every misuse is a status code."""
    import ctypes as C
    import cimpc_b200 as cb
    from cimpc_b200 import package
    capi = package.capi
    m, lin, gait, ref = _reference_traj("flamingo")
    H = 15
    im = cb.ImplicitTrajectory(*SIZES["flamingo"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode="configurationforce", opts=cb.InteriorPointOptions(diff_sol=True))
    oq, ou = np.ones((H, m.nq)), np.ones((H, m.nu))
    with pytest.raises(ValueError):  # host mirror: force mode needs the γ, b weights
        cb.Newton(im, H, 4, oq, ou, 2e-4)
    with pytest.raises(cb.CimpcError) as e:  # force weights that are not negligible: unsupported, not silently wrong
        cb.Newton(im, H, 4, oq, ou, 2e-4, obj_gamma=np.full((H, m.nc), 1e-3), obj_b=np.full((H, m.nb), 1e-100))
    assert e.value.code == 2
    with pytest.raises(cb.CimpcError) as e:  # non-positive velocity weight
        cb.Newton(im, H, 4, oq, ou, 2e-4, obj_gamma=np.full((H, m.nc), 1e-100), obj_b=np.full((H, m.nb), 1e-100),
                  obj_v=np.zeros((H, m.nq)))
    assert e.value.code == 1
    # the :configuration-only entry point refuses a :configurationforce context
    lib = capi.load_library()
    no, io = cb.NewtonOptions().to_c(), cb.InteriorPointOptions().to_c()
    rc = lib.cimpc_newton_create(im._ctx, H, 4, oq.ctypes.data, ou.ctypes.data, 2e-4, C.byref(no), C.byref(io))
    assert rc == 2
