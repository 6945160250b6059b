"""End-to-end CPU test of the restated stack against the reference's closed-loop band.

test/controller/mpc_quadruped.jl:1-69 — quadruped, gait2, `update_friction_coefficient!`, H_mpc = 10,
N_sample = 5, κ_mpc = 2e-4, TrackingObjective, `:configuration`, Newton r_tol = 3e-4 / max_iter = 5, IP
r_tol = 1e-8 / κ_tol = κ_mpc / undercut 5 / γ_reg 0.1, nonlinear simulator at h/5 for H_sim = 1000 steps:
    status,  q < 0.0201·1.5,  u < 0.0437·1.5,  γ < 0.374·1.5,  b < 0.0789·1.5.
Everything the hot path depends on is exercised in one loop — sympy models / dynamics / residual, the
linearization, the reconstructed RoboDojo interior-point loop (on both the linearized and the nonlinear
residual), `newton_solve!`, `policy`, `rot_n_stride!` — so this is the strongest pin the reference offers for
the parts whose source is not in the reference tree."""
import numpy as np
import pytest

from common import SIZES, load_gait

H_MPC, N_SAMPLE, KAPPA = 10, 5, 2.0e-4


def _setup(H_sim):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.linearized import linearized_step
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    from oracle.residual import get_residual
    from oracle.simulator import CIMPC, update_friction_coefficient
    from oracle.trajectory import trajectory_from_gait
    robot = "quadruped"
    res = get_residual(robot)
    m = res.model
    gait = load_gait(robot)
    ref = trajectory_from_gait(m, gait)
    update_friction_coefficient(ref, m)  # mpc_quadruped.jl:11
    H = ref.H
    r0 = np.zeros((H, m.nz)); rz0 = np.zeros((H, m.nz, m.nz)); rth0 = np.zeros((H, m.nz, m.ntheta))
    for t in range(H):
        r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], KAPPA)
    lin = dict(z0=ref.z, th0=ref.theta, r0=r0, rz0=rz0, rth0=rth0)
    co = COracle(*SIZES[robot], lin, mode="configuration", solver="mgs")
    ipo = IPOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, gamma_reg=0.1, diff_sol=True)
    nq = m.nq

    def dyn(window, traj):
        knot = np.array(window[:H_MPC], dtype=np.int32)
        z, dz, st, _ = co.solve(knot, traj.theta[:H_MPC], traj.q[2:H_MPC + 2], ipo)
        return z[:, :nq] - traj.q[2:H_MPC + 2], dz[:, :, :nq], dz[:, :, nq:2 * nq], dz[:, :, 2 * nq:]

    obj = TrackingObjective(q=np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.25] * (nq - 3)), (H_MPC, 1)),
                            u=np.tile(3e-2 * np.ones(m.nu), (H_MPC, 1)),
                            gamma=np.full((H_MPC, m.nc), 1e-100), b=np.full((H_MPC, m.nb), 1e-100))
    newton = Newton(m, H_MPC, gait["h"], obj, KAPPA, NewtonOptions(r_tol=3e-4, max_iter=5))
    policy = CIMPC(m, ref, newton, dyn, H_MPC, N_SAMPLE)
    return res, m, ref, policy, gait


@pytest.mark.parametrize("H_sim", [1000])
def test_mpc_quadruped_tracking_error_band(H_sim):
    from oracle.simulator import simulate, tracking_error
    res, m, ref, policy, gait = _setup(H_sim)
    h = gait["h"]
    q1, v1 = ref.q[1].copy(), (ref.q[1] - ref.q[0]) / h  # initial_conditions, trajectory.jl:219-224
    ok, q, u, gam, b = simulate(res, policy, q1, v1, H_sim, h / N_SAMPLE, m.mu_world)
    assert ok
    qe, ue, ge, be = tracking_error(ref, m, q, u, gam, b, N_SAMPLE, idx_shift=(0,))
    print(f"tracking errors q {qe:.4f} (ref 0.0201)  u {ue:.4f} (0.0437)  γ {ge:.4f} (0.374)  b {be:.4f} (0.0789)")
    assert qe < 0.0201 * 1.5
    assert ue < 0.0437 * 1.5
    assert ge < 0.374 * 1.5
    assert be < 0.0789 * 1.5


def test_mpc_flamingo_tracking_error_band():
    """test/controller/mpc_flamingo.jl:1-80 — flamingo, gait_forward_36_4, H_mpc = 15, default mode
    :configurationforce, TrackingVelocityObjective, simulator γ_reg = 0 / ϵ_min = 0.05:
        q < 0.0154·1.5, u < 0.0829·1.5, γ < 0.444·1.5, b < 0.0169·1.5  and the initial-condition identities :66-69."""
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.linearized import linearized_step
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    from oracle.residual import get_residual
    from oracle.simulator import CIMPC, simulate, tracking_error
    from oracle.trajectory import trajectory_from_gait
    robot, H_mpc, N, kappa, H_sim = "flamingo", 15, 5, 2.0e-4, 1000
    res = get_residual(robot)
    m = res.model
    gait = load_gait(robot)
    ref = trajectory_from_gait(m, gait)  # no update_friction_coefficient! here (mpc_flamingo.jl:8-10)
    H = ref.H
    r0 = np.zeros((H, m.nz)); rz0 = np.zeros((H, m.nz, m.nz)); rth0 = np.zeros((H, m.nz, m.ntheta))
    for t in range(H):
        r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], kappa)
    lin = dict(z0=ref.z, th0=ref.theta, r0=r0, rz0=rz0, rth0=rth0)
    co = COracle(*SIZES[robot], lin, mode="configurationforce", solver="mgs")
    ipo = IPOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
    nq, nc, nb = m.nq, m.nc, m.nb

    def dyn(window, traj):
        knot = np.array(window[:H_mpc], dtype=np.int32)
        z, dz, st, _ = co.solve(knot, traj.theta[:H_mpc], traj.q[2:H_mpc + 2], ipo)
        d = z[:, :nq + nc + nb].copy()
        d[:, :nq] -= traj.q[2:H_mpc + 2]
        d[:, nq:nq + nc] -= traj.gamma[:H_mpc]
        d[:, nq + nc:] -= traj.b[:H_mpc]
        return d, dz[:, :, :nq], dz[:, :, nq:2 * nq], dz[:, :, 2 * nq:]

    one = np.ones
    obj = TrackingObjective(
        q=np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_mpc, 1)),
        u=np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_mpc, 1)),
        gamma=np.full((H_mpc, nc), 1e-100), b=np.full((H_mpc, nb), 1e-100),
        v=np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_mpc, 1)))
    newton = Newton(m, H_mpc, gait["h"], obj, kappa, NewtonOptions(r_tol=3e-4, max_iter=5), mode="configurationforce")
    policy = CIMPC(m, ref, newton, dyn, H_mpc, N)
    h = gait["h"]
    q1, v1 = ref.q[1].copy(), (ref.q[1] - ref.q[0]) / h
    sim_opts = IPOptions(r_tol=1e-8, kappa_tol=1e-8, undercut=np.inf, gamma_reg=0.0, eps_min=0.05, diff_sol=False)
    ok, q, u, gam, b = simulate(res, policy, q1, v1, H_sim, h / N, m.mu_world, sim_opts)
    assert ok
    # :66-69
    assert np.linalg.norm(q[0] - ref.q[1] * (1 - 1 / N) - ref.q[0] / N) < 1e-8
    assert np.linalg.norm(q[1] - ref.q[1]) < 1e-8
    assert np.linalg.norm(ref.q[0][1:] - ref.q[-2][1:]) < 1e-8 and np.linalg.norm(ref.q[1][1:] - ref.q[-1][1:]) < 1e-8
    qe, ue, ge, be = tracking_error(ref, m, q, u, gam, b, N, idx_shift=(0,))
    print(f"flamingo tracking errors q {qe:.4f} (ref 0.0154)  u {ue:.4f} (0.0829)  γ {ge:.4f} (0.444)  b {be:.4f} (0.0169)")
    assert qe < 0.0154 * 1.5 and ue < 0.0829 * 1.5 and ge < 0.444 * 1.5 and be < 0.0169 * 1.5
