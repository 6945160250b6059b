"""CPU tests (`-m "not gpu"`): the oracle against the reference's own tests and fixtures.

Each test names the reference test it restates (file:line under the reference tree)."""
import numpy as np
import pytest

from common import SIZES, FixtureIndex, load_gait, load_lin, make_batch, oracle_problems, oracle_solve_batch
from oracle.ip import IPOptions, interior_point_solve
from oracle.linearized import MGS, LinProblem, Schur, lin_blocks


def test_mgs_qr_solve():
    """test/solver/qr.jl:1-25 — MGS-QR vector solve, n = 16, ‖A x − b‖∞ < 1e-10."""
    rng = np.random.default_rng(0)
    n = 16
    A = rng.random((n, n))
    b = rng.random(n)
    gs = MGS(n)
    gs.factorize(A)
    x = gs.solve(b)
    assert np.abs(A @ x - b).max() < 1e-10
    # packed R is upper triangular with positive diagonal, Q orthonormal
    Q = gs.qs.T
    assert np.abs(Q.T @ Q - np.eye(n)).max() < 1e-10


def test_schur_solve_random():
    """test/solver/schur.jl:1-17 — n = 11, m = 16, factorize with a *new* D, residual < 1e-9."""
    rng = np.random.default_rng(1)
    n, m = 11, 16
    M = rng.random((n + m, n + m))
    A, B, C, D = M[:n, :n], M[:n, n:], M[n:, :n], M[n:, n:]
    S = Schur(A, B, C, D)
    D2 = rng.random((m, m))
    S.factorize(D2)
    u, v = rng.random(n), rng.random(m)
    x, y = S.solve(u, v)
    M2 = np.block([[A, B], [C, D2]])
    assert np.abs(M2 @ np.concatenate([x, y]) - np.concatenate([u, v])).max() < 1e-9


def test_schur_ill_conditioned_flamingo():
    """test/solver/schur.jl:19-62 — Schur built from the flamingo Jacobian at max.(1e-6, ref_traj.z[10]),
    RHS u = r0[dyn], v = r0[rst] − r0[bil] ./ z0[y1]; reduced-system residual < 1e-7."""
    from oracle.residual import get_residual
    res = get_residual("flamingo")
    idx = res.idx
    lin = load_lin("flamingo")
    t = 9  # Julia ref_traj.z[10]
    z0 = np.maximum(1e-6, lin["z0"][t])
    th0 = lin["th0"][t]
    r0, rz0 = res.r(z0, th0, 0.0), res.rz(z0, th0)
    A, B = rz0[np.ix_(idx.dyn, idx.x)], rz0[np.ix_(idx.dyn, idx.y1)]
    Cm = rz0[np.ix_(idx.rst, idx.x)]
    D = rz0[np.ix_(idx.rst, idx.y1)] - rz0[np.ix_(idx.rst, idx.y2)] @ np.diag(z0[idx.y2] / z0[idx.y1])
    S = Schur(A, B, Cm, D)
    u = r0[idx.dyn]
    v = r0[idx.rst] - r0[idx.bil] / z0[idx.y1]
    S.factorize(D)
    x, y = S.solve(u, v)
    M1 = np.block([[A, B], [Cm, D]])
    assert np.abs(M1 @ np.concatenate([x, y]) - np.concatenate([u, v])).max() < 1e-7


@pytest.fixture(scope="module")
def hopper_residual():
    from oracle.residual import get_residual
    return get_residual("hopper_2D")


def test_linearized_solver_hooks(hopper_residual):
    """test/controller/linearized_solver.jl:1-68 — hopper_2D at random z0, θ0."""
    res = hopper_residual
    idx = res.idx
    rng = np.random.default_rng(3)
    z0, th0, kappa = rng.random(idx.nz), rng.random(idx.ntheta), 1e-4
    r0, rz0, rth0 = res.r(z0, th0, kappa), res.rz(z0, th0), res.rth(z0, th0)
    blk = lin_blocks(idx, res.model.nc, z0, th0, r0, rz0, rth0)
    # :38-44 block slicing
    assert np.array_equal(blk.Dx, rz0[np.ix_(idx.dyn, idx.x)])
    assert np.array_equal(blk.Ry2, np.diag(rz0[np.ix_(idx.rst, idx.y2)]))
    assert np.allclose(np.diag(rz0[np.ix_(idx.bil, idx.y1)]), z0[idx.y2], atol=1e-12)
    assert np.allclose(np.diag(rz0[np.ix_(idx.bil, idx.y2)]), z0[idx.y1], atol=1e-12)
    p = LinProblem(blk, idx)
    # :47-52 rlin! reproduces r0 at the linearization point
    p.r(z0, th0, kappa)
    assert np.abs(r0[idx.dyn] - p.rdyn).max() < 1e-10
    assert np.abs(r0[idx.rst] - p.rrst).max() < 1e-10
    assert np.abs(r0[idx.bil] - p.rbil).max() < 1e-10
    # :55-57 linear_solve!(Δ, rz, r) == rz0 \ r0
    p.rz(z0)
    assert np.abs(p.solve() - np.linalg.solve(rz0, r0)).max() < 1e-10
    # :61 rθ0[ibil, :] == 0 ; :63-67 linear_solve!(δz, rz, rθ) == rz0 \ rθ0
    assert np.abs(rth0[idx.bil, :]).max() < 1e-10
    assert np.abs(p.solve_sensitivity() - np.linalg.solve(rz0, rth0)).max() < 1e-10


@pytest.mark.parametrize("robot,tol", [("quadruped", 1e-4), ("flamingo", 1e-4)])
def test_gait_residual_identity(robot, tol):
    """test/simulator/quadruped.jl:16-19 — ‖residual(z_t, θ_t, 0)‖ < 1e-4 on the shipped gait, with
    `update_friction_coefficient!` (μ of the gait in both z and θ).  Pins the sympy restatement of
    model + dynamics + residual against the reference's own fixture."""
    from oracle.residual import get_residual
    from oracle.trajectory import phi_numeric
    res = get_residual(robot)
    m = res.model
    g = load_gait(robot)
    for t in range(g["u"].shape[0]):
        q2 = g["q"][t + 2]
        s2 = g["mu"] * g["gamma"][t] - g["b"][t].reshape(m.nc, m.nf).sum(1)
        z = np.concatenate([q2, g["gamma"][t], g["b"][t], g["psi"][t], phi_numeric(m, q2), g["eta"][t], s2])
        th = np.concatenate([g["q"][t], g["q"][t + 1], g["u"][t], np.zeros(m.nw), [g["mu"]], [g["h"]]])
        assert np.linalg.norm(res.r(z, th, 0.0)) < tol
    # the committed linearization fixture is what this model produces (spot-check three knots)
    lin = load_lin(robot)
    for t in (0, 17, lin["z0"].shape[0] - 1):
        assert np.abs(res.rz(lin["z0"][t], lin["th0"][t]) - lin["rz0"][t]).max() < 1e-9
        assert np.abs(res.rth(lin["z0"][t], lin["th0"][t]) - lin["rth0"][t]).max() < 1e-9
        assert np.abs(res.r(lin["z0"][t], lin["th0"][t], float(lin["kappa"])) - lin["r0"][t]).max() < 1e-12


def test_jacobians_vs_finite_differences(hopper_residual):
    """test/dynamics/*.jl style cross-check: symbolic rz, rθ equal central differences."""
    res = hopper_residual
    rng = np.random.default_rng(4)
    z, th = rng.random(res.idx.nz) + 0.1, rng.random(res.idx.ntheta) + 0.1
    eps = 1e-6
    for J, n, which in ((res.rz(z, th), res.idx.nz, 0), (res.rth(z, th), res.idx.ntheta, 1)):
        for j in range(n):
            a, b = [z.copy(), th.copy()], [z.copy(), th.copy()]
            a[which][j] += eps
            b[which][j] -= eps
            fd = (res.r(a[0], a[1], 0.0) - res.r(b[0], b[1], 0.0)) / (2 * eps)
            assert np.abs(fd - J[:, j]).max() < 1e-6


def test_implicit_dynamics_reproduces_gait():
    """test/controller/implicit_dynamics.jl:1-25 — quadruped gait2, κ = 1e-4, :configuration,
    r_tol = 1e-8, κ_tol = 2e-4: after implicit_dynamics!(im_traj, ref_traj), ‖dq2[t]‖∞ < 1e-2."""
    lin = load_lin("quadruped")
    nq = SIZES["quadruped"][0]
    H = lin["z0"].shape[0]
    knot = np.arange(H, dtype=np.int32)
    theta, q2 = lin["th0"], lin["z0"][:, :nq]
    z, dz, st, it = oracle_solve_batch("quadruped", lin, knot, theta, q2,
                                       IPOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True))
    assert st.all()
    assert np.abs(z[:H - 1, :nq] - q2[:H - 1]).max() < 1e-2
    assert np.abs(dz).max() > 0  # test/solver/ldl.jl: δz ≠ 0


def test_ip_solves_random_equality_qp_like_problem():
    """test/solver/ldl.jl:1-112 in spirit: status true and ‖r‖∞ < r_tol, κ-violation < κ_tol on every
    knot of two robots from the cold start, checked with the independently recomputed residual."""
    from common import independent_violation
    for robot in ("quadruped", "flamingo"):
        lin, gait = load_lin(robot), load_gait(robot)
        knot, theta, q2 = make_batch(robot, lin, gait, lin["z0"].shape[0], seed=5)
        o = IPOptions(r_tol=1e-8, kappa_tol=1e-5, diff_sol=False)
        z, _, st, it = oracle_solve_batch(robot, lin, knot, theta, q2, o)
        assert st.all() and it.max() < 40
        rv, kv = independent_violation(robot, lin, knot, theta, z)
        assert rv.max() < 1e-8 * 1.001 and kv.max() < 1e-5 * 1.001


def test_c_restatement_matches_numpy_oracle():
    """The C restatement (CPU baseline) and the numpy oracle are the same algorithm.
    * With the accurate LU variant and the noise-sensitive residual back-tracking disabled (max_ls = 0)
      the two agree to ~1e-11 on EVERY problem: the iteration is well defined.
    * With the reference's MGS-QR they agree only to the reference's own round-off floor
      (≈ cond(S)²·eps: 1e-8 quadruped, 1e-5 flamingo) — documented here, used by the GPU parity tests."""
    from oracle.c_oracle import COracle
    floors = {}
    for robot, mode in (("quadruped", "configuration"), ("flamingo", "configurationforce"),
                        ("centroidal_quadruped", "configuration")):
        lin, gait = load_lin(robot), load_gait(robot)
        knot, theta, q2 = make_batch(robot, lin, gait, 90, seed=6)
        o = IPOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True, max_ls=0)
        for solver, ztol, dztol in (("lu", 1e-10, 1e-9), ("mgs", 1e-4, 1e-3)):
            z, dz, st, it = oracle_solve_batch(robot, lin, knot, theta, q2, o, mode=mode, solver=solver)
            zc, dzc, stc, itc = COracle(*SIZES[robot], lin, mode=mode, solver=solver).solve(knot, theta, q2, o)
            assert st.all() and np.array_equal(st, stc)
            assert np.array_equal(it, itc)
            ez = np.abs(z - zc).max()
            assert ez < ztol, (robot, solver, ez)
            assert (np.abs(dz - dzc).max() / np.abs(dz).max()) < dztol
            floors[robot, solver] = ez
    # the MGS floor really is orders of magnitude above the LU floor on the ill-conditioned robot
    assert floors["flamingo", "mgs"] > 100 * floors["flamingo", "lu"]


def test_hopper_single_policy_call():
    """BASELINE config 1 — hopper flat (examples/hopper/flat.jl:15-51): ONE `policy` call of the CI-MPC on
    the CPU (plumbing / correctness, no GPU): ImplicitTrajectory from the shipped gait (a serialized
    ContactTraj, :joint_traj), H_mpc = 10, κ = 2e-4, IP r_tol = 1e-8 / κ_tol = 2e-4, Newton r_tol = 3e-4,
    max_iter = 5, then rot_n_stride! / update_window! exactly as policy.jl:119-131."""
    from oracle.c_oracle import COracle
    from oracle.models import get_model
    from oracle.newton import (Newton, NewtonOptions, TrackingObjective, get_stride, rot_n_stride, update_window)
    from oracle.trajectory import ContactTraj
    robot = "hopper_2D"
    m = get_model(robot)
    lin, gait = load_lin(robot), load_gait(robot)
    H_ref, H = lin["z0"].shape[0], 10
    ref = ContactTraj(m, H_ref, gait["h"])
    ref.q[:] = gait["q"]; ref.u[:] = gait["u"]; ref.w[:] = gait["w"]; ref.gamma[:] = gait["gamma"]; ref.b[:] = gait["b"]
    ref.z[:] = lin["z0"]; ref.theta[:] = lin["th0"]
    # the implicit dynamics reproduce the gait (the hopper analogue of test/controller/implicit_dynamics.jl:24)
    co = COracle(*SIZES[robot], lin, mode="configuration", solver="mgs")
    ipo = IPOptions(r_tol=1e-8, kappa_tol=2e-4, undercut=5.0, diff_sol=True)
    z, dz, st, it = co.solve(np.arange(H_ref, dtype=np.int32), ref.theta, ref.q[2:], ipo)
    assert st.all() and np.abs(z[:, :m.nq] - ref.q[2:]).max() < 1e-2

    def dyn(window, traj):
        knot = np.array(window[:H], dtype=np.int32)
        zz, dzz, stt, _ = co.solve(knot, traj.theta[:H], traj.q[2:H + 2], ipo)
        assert stt.all()
        nq = m.nq
        return zz[:, :nq] - traj.q[2:H + 2], dzz[:, :, :nq], dzz[:, :, nq:2 * nq], dzz[:, :, 2 * nq:]

    obj = TrackingObjective(q=np.tile(1e-1 * np.array([0.1, 3.0, 1.0, 3.0]), (H, 1)),   # flat.jl:31-35
                            u=np.tile(np.array([1e-3, 1.0]), (H, 1)),
                            gamma=np.full((H, m.nc), 1e-100), b=np.full((H, m.nb), 1e-100))
    core = Newton(m, H, gait["h"], obj, 2.0e-4, NewtonOptions(r_tol=3e-4, max_iter=5))
    ptraj, window = ref.copy(), list(range(H + 2))
    q0, q1 = ref.q[0].copy(), ref.q[1].copy()
    u = core.solve(dyn, q0, q1, window, ptraj, warm_start=False)   # policy(p, traj, 1)
    assert np.isfinite(u).all()
    # started on the reference: the plan stays on it and the control equals the reference control
    assert np.abs(core.traj.q - ref.q[:H + 2]).max() < 5e-3
    assert np.abs(u - ref.u[0]).max() < 5e-2 * max(1.0, np.abs(ref.u[0]).max())
    assert np.abs(core.res).sum() / len(core.res) < 3e-4 or core.stats["iters"] == 5
    rot_n_stride(ptraj, get_stride(m, ref))
    window = update_window(window, H_ref)
    assert window[0] == 1 and np.allclose(ptraj.q[0], ref.q[1])
    assert np.allclose(ptraj.q[H_ref + 1] - ptraj.q[1], get_stride(m, ref))


def test_piecewise_terrain_and_its_approx_jacobian():
    """`piecewise1_2D_lc` (src/simulation/environments/piecewise.jl:40-125): the surface is C¹ across the smoothed kinks
    (the reference asserts the cubic fits at :57-58, :77-78), the rotation of environment.jl:81-96 is the pair
    (1, s')/√(1+s'²), and on the straight pieces — where the rotation does not change — the `approx` Jacobian of
    residual_approx.jl:14-99 IS the Jacobian of the residual (central differences)."""
    from oracle.residual import get_residual
    from oracle.terrain import get_terrain
    ter = get_terrain("piecewise1_2D_lc")
    m_ss = np.tan(np.deg2rad(10.0))
    assert ter.height(0.2) == 0.0 and abs(ter.height(1.0) - 0.5 * m_ss) < 1e-15 and abs(ter.height(3.0) - 1.25 * m_ss) < 1e-15
    for x in (0.4, 0.6, 1.9, 2.1):
        assert abs(ter.height(x - 1e-9) - ter.height(x + 1e-9)) < 1e-8 and abs(ter.slope(x - 1e-9) - ter.slope(x + 1e-9)) < 1e-7
    for x in (0.1, 0.5, 1.0, 2.0, 3.0):
        c, s = ter.rotation(x)
        g = ter.slope(x)
        assert abs(c - 1 / np.hypot(1, g)) < 1e-15 and abs(s - g / np.hypot(1, g)) < 1e-15
    res = get_residual("hopper_2D_piecewise")
    m = res.model
    rng = np.random.default_rng(0)
    z = np.abs(rng.standard_normal(m.nz)) + 0.1
    th = rng.standard_normal(m.ntheta); th[-1] = 0.01; th[-2] = 0.8
    for foot_x, straight in ((0.2, True), (1.0, True), (3.0, True), (0.5, False)):
        z[0] = foot_x - z[3] * np.sin(z[2])
        J = res.rz(z, th)
        Jf = np.zeros_like(J)
        for j in range(m.nz):
            e = np.zeros(m.nz); e[j] = 1e-6
            Jf[:, j] = (res.r(z + e, th, 0.0) - res.r(z - e, th, 0.0)) / 2e-6
        if straight:
            assert np.abs(J - Jf).max() < 1e-6
        else:  # inside a kink the neglected curvature terms are there
            assert np.abs(J - Jf).max() > 1e-2
