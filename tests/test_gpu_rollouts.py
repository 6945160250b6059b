"""Whole closed loop on the GPU (`MonteCarloRollouts`: batched `newton_solve!` + batched simulator step) for
the setting of test/controller/mpc_quadruped.jl: the reference's tracking-error band must hold for the nominal
rollout, and the GPU loop must track as well as the all-CPU oracle loop."""
import numpy as np
import pytest

from common import SIZES

pytestmark = pytest.mark.gpu

H_MPC, N_SAMPLE, KAPPA = 10, 5, 2.0e-4


def test_monte_carlo_rollouts_on_device(cuda_device):
    import torch
    import cimpc_b200 as cb
    from oracle.linearized import linearized_step
    from oracle.simulator import simulate, tracking_error
    from test_closed_loop import _setup
    H_sim, R = 1000, 8
    res, m, ref, cpu_policy, gait = _setup(H_sim)
    h, nq = gait["h"], m.nq
    Hr = ref.H
    r0 = np.zeros((Hr, m.nz)); rz0 = np.zeros((Hr, m.nz, m.nz)); rth0 = np.zeros((Hr, m.nz, m.ntheta))
    for t in range(Hr):
        r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], KAPPA)
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, gamma_reg=0.1, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], ref.z, ref.theta, r0, rz0, rth0, mode="configuration", opts=ipo)
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.25] * (nq - 3)), (H_MPC, 1))
    ou = np.tile(3e-2 * np.ones(m.nu), (H_MPC, 1))
    mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_MPC, N_sample=N_SAMPLE,
                               obj_q=oq, obj_u=ou, kappa=KAPPA, n_rollouts=R,
                               newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
    rng = np.random.default_rng(60)
    q1 = np.tile(ref.q[1], (R, 1))
    v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
    v1[1:] *= 1.0 + 0.05 * rng.standard_normal((R - 1, 1))  # rollout 0 = the reference's initial condition
    out = mc.run(torch.from_numpy(q1).to(cuda_device), torch.from_numpy(v1).to(cuda_device), H_sim)
    torch.cuda.synchronize()
    ok = out["status"].cpu().numpy()
    # a simulator step that the interior-point iteration cannot solve ends that rollout (as `simulate!` does);
    # the reconstructed iteration fails on ≈ 1 step in 4 000 on this gait — the oracle fails on the same inputs
    # (scripts/gpu_sim_debug.py) — so most, not all, 1000-step rollouts run to the end
    assert ok[0], "the nominal rollout must complete (test/controller/mpc_quadruped.jl:59 asserts `status`)"
    assert ok.sum() >= R // 2, f"only {ok.sum()} of {R} rollouts completed"
    assert mc.mpc_steps == H_sim // N_SAMPLE
    q, u, gam, b = (out[k].cpu().numpy() for k in ("q", "u", "gamma", "b"))
    e = np.array([tracking_error(ref, m, q[:, r], u[:, r], gam[:, r], b[:, r], N_SAMPLE, idx_shift=(0,)) for r in range(R)])
    e_ok = e[ok]
    print("GPU closed-loop tracking errors (q, u, γ, b) of completed rollouts:\n", np.round(e_ok, 4), "\nfailed at", out["failed_at"].cpu().numpy())
    band = np.array([0.0201, 0.0437, 0.374, 0.0789]) * 1.5  # mpc_quadruped.jl:61-64
    assert (e[0] < band).all()  # the nominal rollout = the reference's own test
    assert (e_ok < 1.25 * band).all()
    assert (np.median(e_ok, axis=0) < band).all()
    # same quality as the all-CPU oracle loop
    ok_c, qc, uc, gc, bc = simulate(res, cpu_policy, q1[0], v1[0], H_sim, h / N_SAMPLE, m.mu_world)
    e_cpu = tracking_error(ref, m, qc, uc, gc, bc, N_SAMPLE, idx_shift=(0,))
    print("CPU-oracle loop:", np.round(e_cpu, 4))
    assert np.all(np.abs(np.median(e_ok, axis=0) - e_cpu) <= 0.2 * e_cpu + 1e-3)


def test_flamingo_closed_loop_on_device(cuda_device):
    """test/controller/mpc_flamingo.jl:1-80 entirely on the GPU: flamingo, gait_forward_36_4, H_mpc = 15, N_sample = 5,
    :configurationforce, TrackingVelocityObjective (general device Newton), device-side linearization, simulator with
    γ_reg = 0 / ϵ_min = 0.05.  Reference band: q < 0.0154·1.5, u < 0.0829·1.5, γ < 0.444·1.5, b < 0.0169·1.5."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait
    from oracle.residual import get_residual
    from oracle.simulator import tracking_error
    from oracle.trajectory import trajectory_from_gait
    robot, H_mpc, N, kappa, H_sim, R = "flamingo", 15, 5, 2.0e-4, 1000, 8
    res = get_residual(robot)
    m = res.model
    gait = load_gait(robot)
    ref = trajectory_from_gait(m, gait)
    h = gait["h"]
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES[robot], ref.z, ref.theta, kappa=kappa, mode="configurationforce", opts=ipo)
    oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_mpc, 1))
    ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_mpc, 1))
    ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_mpc, 1))
    sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"),
                                       gamma_reg=0.0)
    mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_mpc, N_sample=N, obj_q=oq,
                               obj_u=ou, kappa=kappa, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                               sim_opts=sim_opts, obj_gamma=np.full((H_mpc, m.nc), 1e-100),
                               obj_b=np.full((H_mpc, m.nb), 1e-100), obj_v=ov, ref_gamma=ref.gamma, ref_b=ref.b)
    rng = np.random.default_rng(61)
    q1 = np.tile(ref.q[1], (R, 1))
    v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
    v1[1:] *= 1.0 + 0.02 * rng.standard_normal((R - 1, 1))
    out = mc.run(torch.from_numpy(q1).to(cuda_device), torch.from_numpy(v1).to(cuda_device), H_sim)
    torch.cuda.synchronize()
    ok = out["status"].cpu().numpy()
    assert ok[0], "the nominal rollout must complete (test/controller/mpc_flamingo.jl:69 asserts `status`)"
    assert ok.sum() >= R // 2, f"only {ok.sum()} of {R} rollouts completed"
    q, u, gam, b = (out[k].cpu().numpy() for k in ("q", "u", "gamma", "b"))
    e = np.array([tracking_error(ref, m, q[:, r], u[:, r], gam[:, r], b[:, r], N, idx_shift=(0,)) for r in range(R)])
    e_ok = e[ok]
    print("flamingo GPU closed-loop tracking errors (q, u, γ, b):\n", np.round(e_ok, 4), "\nfailed at", out["failed_at"].cpu().numpy())
    band = np.array([0.0154, 0.0829, 0.444, 0.0169]) * 1.5  # mpc_flamingo.jl:72-75
    assert (np.median(e_ok, axis=0) < band).all()
    assert (e[0] < band).all()  # the nominal rollout = the reference's own test


def test_altitude_update_and_disturbances_in_the_loop(cuda_device):
    """examples/flamingo/piecewise.jl:37-47 options on the device loop: `altitude_update = true`,
    `altitude_impact_threshold = 0.02` — every rollout's altitude vector is refreshed from the simulator step of
    largest impact of the last N_sample steps (mpc_utils.jl:109-135) — plus an impulse disturbance
    (src/simulator/disturbances.jl:41-61) on the simulator's w.  Checked against a host re-computation of
    `update_altitude!` from the recorded trajectory, and the disturbed rollouts must differ from the undisturbed run."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait
    from oracle.residual import get_residual
    from oracle.trajectory import phi_numeric, trajectory_from_gait
    robot, H_mpc, N, kappa, H_sim, R = "flamingo", 15, 5, 1.0e-4, 120, 6
    res = get_residual(robot)
    m = res.model
    gait = load_gait(robot)
    ref = trajectory_from_gait(m, gait)
    h = gait["h"]
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
    oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_mpc, 1))
    ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_mpc, 1))
    ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_mpc, 1))
    sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"), gamma_reg=0.0)

    def loop(dist):
        im = cb.ImplicitTrajectory(*SIZES[robot], ref.z, ref.theta, kappa=kappa, mode="configurationforce", opts=ipo)
        mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_mpc, N_sample=N, obj_q=oq,
                                   obj_u=ou, kappa=kappa, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                                   sim_opts=sim_opts, obj_gamma=np.full((H_mpc, m.nc), 1e-100),
                                   obj_b=np.full((H_mpc, m.nb), 1e-100), obj_v=ov, ref_gamma=ref.gamma, ref_b=ref.b,
                                   altitude_update=True, altitude_impact_threshold=0.02)
        q1 = np.tile(ref.q[1], (R, 1))
        v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
        out = mc.run(torch.from_numpy(q1).to(cuda_device), torch.from_numpy(v1).to(cuda_device), H_sim, dist=dist)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items() if v is not None}

    a = loop(None)
    assert a["status"].all()
    # host re-computation of update_altitude! at the last policy call (t = H_sim − N + 1) for rollout 0
    alt = np.zeros(m.nc)
    for t in range(1, H_sim + 1):
        if (t - 1) % N == 0 and t > 1:
            lo = max(0, t - 1 - N) + 1
            for i in range(m.nc):
                g = a["gamma"][lo - 1:t - 1, 0, i]
                if g.max() > 0.02:
                    j = lo + int(np.argmax(g))                    # 1-based simulator step of the largest impact
                    alt[i] = phi_numeric(m, a["q"][j + 1, 0])[i]  # ϕ(traj.q[j + 2]) (1-based) = q array index j + 1
    assert np.abs(a["alt"][0] - alt).max() < 1e-7, (a["alt"][0], alt)
    assert np.abs(a["alt"]).max() < 5e-3          # flat ground: the measured altitudes are contact-level heights
    # an impulse on the simulator's w at step 20 (w enters the dynamics through A(q)ᵀ w, model.jl:33)
    b = loop(cb.ImpulseDisturbance([[2.0, 0.0]], [20]))
    assert np.array_equal(a["q"][:20], b["q"][:20])
    assert np.abs(a["q"][25] - b["q"][25]).max() > 1e-4


def test_payload_plant_under_nominal_policy_in_rollout_groups(cuda_device):
    """examples/quadruped/payload.jl: the policy is built on the nominal quadruped, the simulated plant carries the
    payload (`quadruped_payload`: + 3 kg, + 0.03 kg m² on the torso).  Two rollout GROUPS in one run — nominal plant and
    payload plant — each with its own context (`GroupedRollouts(group_kwargs=…)`): the nominal group reproduces the
    plain run bit for bit, the payload group stays on the gait (the reference's claim for this example) with a visibly
    different trajectory and larger normal forces."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait, load_lin
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    nq, nu = SIZES[robot][0], SIZES[robot][1]
    R, N, H_sim = 8, 5, 300
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-4, undercut=5.0, diff_sol=True)
    oq = np.tile(1e-2 * np.array([10.0, 0.02, 0.25] + [0.5] * (nq - 3)), (H_MPC, 1))   # payload.jl:36-40
    ou = np.tile(3e-2 * np.ones(nu), (H_MPC, 1))

    def make_im():
        return cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                     mode="configuration", opts=opts)
    q1 = torch.from_numpy(np.tile(gait["q"][1], (R, 1))).to(cuda_device)
    v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(cuda_device)
    kw = dict(H_mpc=H_MPC, N_sample=N, obj_q=oq, obj_u=ou, kappa=1e-4, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
              altitude_update=True, altitude_impact_threshold=0.05)                     # payload.jl:48-53
    mc = cb.GroupedRollouts(make_im, R, 2, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"],
                            group_kwargs=[dict(), dict(sim_model="quadruped_payload")], **kw)
    out = {k: v.cpu().numpy() for k, v in mc.run(q1, v1, H_sim).items()}
    plain = cb.MonteCarloRollouts(make_im(), gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], n_rollouts=R // 2, **kw)
    ref = {k: v.cpu().numpy() for k, v in plain.run(q1[:R // 2].contiguous(), v1[:R // 2].contiguous(), H_sim).items() if v is not None}
    assert out["status"].all(), out["failed_at"]
    assert np.array_equal(out["q"][:, :R // 2], ref["q"])
    qn, qp = out["q"][:, 0], out["q"][:, R // 2]
    assert np.abs(qn - qp).max() > 1e-3                                  # the payload changes the motion …
    assert np.abs(qp[-1, 1] - qn[-1, 1]) < 0.05                          # … but the robot keeps walking at height
    gn, gp = out["gamma"][:, 0].sum(), out["gamma"][:, R // 2].sum()
    assert gp > 1.1 * gn                                                 # 3 kg on 14.5 kg: ≈ 20 % more normal impulse


def test_grouped_rollouts_are_bit_identical(cuda_device):
    """`GroupedRollouts` (independent parts on their own streams / host threads) changes scheduling only."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait, load_lin
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    nq, nu = SIZES[robot][0], SIZES[robot][1]
    R, N = 96, 5
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H_MPC, 1))
    ou = np.tile(3e-2 * np.ones(nu), (H_MPC, 1))

    def make_im():
        return cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                     mode="configuration", opts=opts)
    q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=3)).to(cuda_device)
    v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(cuda_device)
    outs = []
    for G in (1, 3):
        mc = cb.GroupedRollouts(make_im, R, G, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H_MPC, N_sample=N,
                                obj_q=oq, obj_u=ou, kappa=1e-4, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
        out = mc.run(q1, v1, 3 * N)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in out.items()})
        assert mc.mpc_steps == 3
    for k in ("q", "u", "gamma", "b", "status", "failed_at"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_hopper_closed_loop_on_device(cuda_device):
    """BASELINE config 1 on the device: examples/hopper/flat.jl:24-51 (hopper_2D, gait_forward, H_mpc = 10, N_sample = 5,
    κ_mpc = 2e-4, IP r_tol 1e-8 / κ_tol 2e-4 / undercut 5, Newton r_tol 3e-4 / max_iter 5, w = 0).  No tracking band is
    published for this example, so the device loop is held to the all-CPU oracle loop run on the same settings."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait, load_lin
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    from oracle.newton import Newton, NewtonOptions, TrackingObjective
    from oracle.residual import get_residual
    from oracle.simulator import CIMPC, simulate, tracking_error
    from oracle.trajectory import ContactTraj
    robot, H_mpc, N, kappa, H_sim, R = "hopper_2D", 10, 5, 2.0e-4, 500, 4
    res = get_residual(robot)
    m = res.model
    lin, gait = load_lin(robot), load_gait(robot)
    Hr = lin["z0"].shape[0]
    ref = ContactTraj(m, Hr, gait["h"])
    ref.q[:] = gait["q"]; ref.u[:] = gait["u"]; ref.gamma[:] = gait["gamma"]; ref.b[:] = gait["b"]
    ref.z[:] = lin["z0"]; ref.theta[:] = lin["th0"]
    h = gait["h"]
    oq = np.tile(1e-1 * np.array([0.1, 3.0, 1.0, 3.0]), (H_mpc, 1))
    ou = np.tile(np.array([1e-3, 1.0]), (H_mpc, 1))
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES[robot], ref.z, ref.theta, kappa=kappa, mode="configuration", opts=ipo)  # device linearization
    mc = cb.MonteCarloRollouts(im, ref.q, ref.u, float(ref.theta[0, -2]), m.mu_world, h, H_mpc=H_mpc, N_sample=N, obj_q=oq,
                               obj_u=ou, kappa=kappa, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
    q1 = np.tile(ref.q[1], (R, 1))
    v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
    v1[1:] *= 1.0 + 0.02 * np.random.default_rng(62).standard_normal((R - 1, 1))
    out = mc.run(torch.from_numpy(q1).to(cuda_device), torch.from_numpy(v1).to(cuda_device), H_sim)
    torch.cuda.synchronize()
    ok = out["status"].cpu().numpy()
    assert ok[0], "the nominal rollout must complete"
    q, u, gam, b = (out[k].cpu().numpy() for k in ("q", "u", "gamma", "b"))
    e_gpu = tracking_error(ref, m, q[:, 0], u[:, 0], gam[:, 0], b[:, 0], N, idx_shift=(0,))
    # the all-CPU oracle loop on the same settings
    co = COracle(*SIZES[robot], lin, mode="configuration", solver="mgs")
    oipo = IPOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)

    def dyn(window, traj):
        knot = np.array(window[:H_mpc], dtype=np.int32)
        z, dz, st, _ = co.solve(knot, traj.theta[:H_mpc], traj.q[2:H_mpc + 2], oipo)
        return z[:, :m.nq] - traj.q[2:H_mpc + 2], dz[:, :, :m.nq], dz[:, :, m.nq:2 * m.nq], dz[:, :, 2 * m.nq:]

    obj = TrackingObjective(q=oq, u=ou, gamma=np.full((H_mpc, m.nc), 1e-100), b=np.full((H_mpc, m.nb), 1e-100))
    newton = Newton(m, H_mpc, h, obj, kappa, NewtonOptions(r_tol=3e-4, max_iter=5))
    policy = CIMPC(m, ref, newton, dyn, H_mpc, N)
    ok_c, qc, uc, gc, bc = simulate(res, policy, q1[0], v1[0], H_sim, h / N, m.mu_world)
    assert ok_c
    e_cpu = tracking_error(ref, m, qc, uc, gc, bc, N, idx_shift=(0,))
    print("hopper tracking errors (q, u, γ, b): GPU", np.round(e_gpu, 4), " CPU oracle", np.round(e_cpu, 4))
    assert np.all(np.abs(e_gpu - e_cpu) <= 0.2 * e_cpu + 2e-3)


def test_drop_box_simulator_failures_match_oracle(cuda_device):
    """Failure RATE of the simulator step over the Monte-Carlo box of examples/quadruped/monte_carlo.jl:76-92, device vs
    oracle on identical inputs (open loop: the reference controls u_t / N_sample, so no MPC in between and every
    difference is the simulator's).  A step fails when the reconstructed interior-point iteration jams: at a stick /
    slide transition of a landing foot a pair (ψ, s2) — or (b, η) — leaves the central path by six orders of magnitude
    (fraction-to-the-boundary τ = 1 − max(r_vio, κ_vio)² → 1 sends the blocking variable to 1e-9 while its partner is
    1e-3), the Newton system reaches cond ≈ 3e9 and every later step length is ≈ 1e-10.  The oracle
    (numpy, dense LAPACK LU) jams on the same states: the two failure sets must coincide up to marginal cases, which
    makes the rate a property of the iteration of oracle/ip.py, not of the CUDA path."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait
    from oracle.residual import get_residual
    from oracle.simulator import nonlinear_ip_solve, simulator_ip_options
    robot, R, T, N = "quadruped", 48, 50, 5
    res = get_residual(robot)
    m, i = res.model, res.idx
    gait = load_gait(robot)
    h = gait["h"] / N
    q1 = cb.quadruped_initial_configurations(R, seed=100)
    v1 = (gait["q"][1] - gait["q"][0]) / gait["h"]
    q0 = q1 - h * v1
    # device
    sim = cb.Simulator(*SIZES[robot])
    qa, qb = torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1).to(cuda_device)
    ok = torch.ones(R, dtype=torch.bool, device=cuda_device)
    fail_dev = np.zeros(R, dtype=int)
    for t in range(T):
        u = torch.from_numpy(np.tile(gait["u"][(t // N) % gait["u"].shape[0]] / N, (R, 1))).to(cuda_device)
        q2, _, _, st, _ = sim.step(qa, qb, u, m.mu_world, h, active=ok.to(torch.uint8))
        st = st.bool() | ~ok
        newly = (ok & ~st).cpu().numpy()
        fail_dev[newly] = t + 1
        ok &= st
        q2 = torch.where(ok[:, None], q2, qb)
        qa, qb = qb, q2
    # oracle
    opts = simulator_ip_options()
    fail_cpu = np.zeros(R, dtype=int)
    for r in range(R):
        a, b = q0[r], q1[r]
        for t in range(T):
            z = np.ones(i.nz); z[i.q2] = b
            th = np.concatenate([a, b, gait["u"][(t // N) % gait["u"].shape[0]] / N, np.zeros(m.nw), [m.mu_world], [h]])
            okc, zs, _ = nonlinear_ip_solve(res, z, th, opts)
            if not okc:
                fail_cpu[r] = t + 1
                break
            a, b = b, zs[i.q2]
    fd, fc = fail_dev > 0, fail_cpu > 0
    print(f"drop box, {T} open-loop simulator steps: device fails {fd.sum()}/{R}, oracle fails {fc.sum()}/{R}, "
          f"both {np.sum(fd & fc)}, same step {np.sum((fail_dev == fail_cpu) & fd)}")
    assert np.sum(fd ^ fc) <= max(2, R // 16), (fail_dev, fail_cpu)
    assert abs(fd.mean() - fc.mean()) <= 0.05


def test_flamingo_climbs_the_piecewise_terrain_on_device(cuda_device):
    """examples/flamingo/piecewise.jl with the simulator ON the terrain (`simulator(s_sim, …)`, s_sim =
    get_simulation("flamingo", "piecewise1_2D_lc", "piecewise", approx = true)): the policy is linearized on flat ground
    (:12, :37) and learns the ground height under every contact from `update_altitude!` (`altitude_update = true`,
    `altitude_impact_threshold = 0.02`, :43-47).  The plant is the `flamingo_piecewise` model of this library.  The nominal
    rollout must walk up the 10° ramp (from x = 0.5) for 2500 steps with its feet on the surface; without the altitude
    update the same policy does not get up the ramp."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait
    from oracle.residual import get_residual
    from oracle.trajectory import trajectory_from_gait
    robot, H_mpc, N, kappa, H_sim, R = "flamingo", 15, 5, 1.0e-4, 2500, 4
    res, tres = get_residual(robot), get_residual("flamingo_piecewise")
    m = res.model
    gait = load_gait(robot)
    ref = trajectory_from_gait(m, gait)
    h = gait["h"]
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
    oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_mpc, 1))
    ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_mpc, 1))
    ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_mpc, 1))
    sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"), gamma_reg=0.0)

    def loop(altitude_update, steps):
        im = cb.ImplicitTrajectory(*SIZES[robot], ref.z, ref.theta, kappa=kappa, mode="configurationforce", opts=ipo)
        mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_mpc, N_sample=N, obj_q=oq,
                                   obj_u=ou, kappa=kappa, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                                   sim_opts=sim_opts, obj_gamma=np.full((H_mpc, m.nc), 1e-100),
                                   obj_b=np.full((H_mpc, m.nb), 1e-100), obj_v=ov, ref_gamma=ref.gamma, ref_b=ref.b,
                                   altitude_update=altitude_update, altitude_impact_threshold=0.02, sim_model="flamingo_piecewise")
        q1 = np.tile(ref.q[1], (R, 1))
        v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
        v1[1:] *= 1.0 + 0.02 * np.random.default_rng(61).standard_normal((R - 1, 1))
        out = mc.run(torch.from_numpy(q1).to(cuda_device), torch.from_numpy(v1).to(cuda_device), steps)
        torch.cuda.synchronize()
        return {k: v.cpu().numpy() for k, v in out.items() if v is not None}

    a = loop(True, H_sim)
    assert a["status"][0], "the nominal rollout must complete"
    assert a["status"].sum() >= R - 1
    q = a["q"][:, 0]
    assert q[-1, 0] > 1.5                                        # it is well up the ramp (which starts at x = 0.5) …
    lift = tres.terrain.height(q[-1, 0])
    assert lift > 0.15 and abs((q[-1, 1] - q[1, 1]) - lift) < 0.03  # … and the body rose with the ground
    gaps = np.array([[res.model.phi_func(list(qt))[c] - tres.terrain.height(x) for c, x in enumerate(tres._px(qt))]
                     for qt in q[1::25]], dtype=float)
    assert gaps.min() > -2e-3                                    # no contact point below the surface
    assert (gaps.min(axis=1) < 2e-3).mean() > 0.9                # and some foot is on it
    # the measured altitudes are the surface heights under the feet at their last impacts
    assert np.all(a["alt"][0] > 0.1) and np.all(a["alt"][0] < lift + 0.06)
    b = loop(False, H_sim)
    assert (not b["status"][0]) or b["q"][-1, 0, 0] < q[-1, 0] - 0.3, "altitude_update = false should not get up the ramp"


def test_fused_simulator_steps_are_bit_identical(cuda_device):
    """`cimpc_sim_steps_batch` (the N_sample simulator steps between two policy calls in ONE launch, every tile of
    rollouts running its steps back to back) against N_sample calls of the single step driven by the host loop of
    `simulate!`: same bits in every output of every step — including rollouts that fail on the way (drop box: rear feet
    below the ground), rollouts switched off from the start, and per-step disturbances."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait
    robot = "quadruped"
    gait = load_gait(robot)
    nq, nu, nw, nc, nb = SIZES[robot]
    R, n = 200, 5
    sim = cb.Simulator(nq, nu, nw, nc, nb)
    h = gait["h"] / n
    q1n = cb.quadruped_initial_configurations(R, seed=11)
    q1n[:8, 2] += 0.6                                     # pitched torsos: steps that fail
    q1 = torch.from_numpy(q1n).to(cuda_device)
    q0 = (q1 - h * torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(cuda_device)).contiguous()
    u = torch.from_numpy(np.tile(gait["u"][0] / n, (R, 1))).to(cuda_device)
    active = torch.ones(R, dtype=torch.uint8, device=cuda_device)
    active[5] = 0
    active[77] = 0
    rng = np.random.Generator(np.random.Philox(4))
    w = torch.from_numpy(0.05 * rng.standard_normal((n, R, nw))).to(cuda_device)
    so = cb.simulator_options()
    capped = cb.InteriorPointOptions(r_tol=so.r_tol, kappa_tol=so.kappa_tol, max_ls=so.max_ls, eps_min=so.eps_min,
                                     undercut=so.undercut, gamma_reg=so.gamma_reg, diff_sol=False, max_iter=9)
    for opts, want_failures in ((None, False), (capped, True)):
        q2s, gs, bs, sts, its, phis = sim.steps(n, q0, q1, u, gait["mu"], h, w=w, active=active, opts=opts)
        torch.cuda.synchronize()
        # the host loop of the single step (rollout.py, fused_steps = False)
        ok = active.bool().clone()
        qa, qb = q0, q1
        n_failed = 0
        for s in range(n):
            q2, g, b, st, it, phi = sim.step(qa, qb, u, gait["mu"], h, w=w[s].contiguous(), active=ok.to(torch.uint8),
                                             want_phi=True, opts=opts)
            st = st.bool() & ok
            q2 = torch.where(st[:, None], q2, qb)
            run = ok  # rollouts that took this step
            assert torch.equal(sts[s].bool(), st), s
            assert torch.equal(q2s[s], q2), s
            assert torch.equal(gs[s][run], g[run]) and torch.equal(bs[s][run], b[run]) and torch.equal(phis[s][run], phi[run]), s
            assert torch.equal(its[s][run], it[run]), s
            assert float(gs[s][~run].abs().sum()) == 0.0 and float(bs[s][~run].abs().sum()) == 0.0
            assert int(its[s][~run].sum()) == 0
            n_failed += int((ok & ~st).sum())
            ok = st
            qa, qb = qb, q2
        if want_failures:  # an iteration cap of 9 ends the slower rollouts at different steps of the launch
            assert n_failed >= 3 and int(ok.sum()) >= 3, (n_failed, int(ok.sum()))
    sim.close()


def test_fused_steps_closed_loop_is_bit_identical(cuda_device):
    """The closed loop with fused simulator steps (default) against the loop that launches every step on its own."""
    import torch
    import cimpc_b200 as cb
    from common import load_gait, load_lin
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    nq, nu = SIZES[robot][0], SIZES[robot][1]
    R, N = 96, 5
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H_MPC, 1))
    ou = np.tile(3e-2 * np.ones(nu), (H_MPC, 1))
    q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=3)).to(cuda_device)
    v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(cuda_device)
    outs = []
    for fused in (True, False):
        im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                   mode="configuration", opts=opts)
        mc = cb.MonteCarloRollouts(im, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H_MPC, N_sample=N, obj_q=oq,
                                   obj_u=ou, kappa=1e-4, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                                   altitude_update=True, fused_steps=fused)
        out = mc.run(q1, v1, 3 * N + 2, record_every=1)   # the last interval is a short one
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in out.items() if v is not None})
        assert mc.mpc_steps == 4
    for k in ("q", "u", "gamma", "b", "status", "failed_at", "alt"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
