"""Drop-in check (`-m gpu`): the CUDA `newton_solve!` batch drives the closed loop of
test/controller/mpc_quadruped.jl (policy every N_sample steps, nonlinear simulator from the oracle on the
CPU in between) for a few rollouts in lock-step, and achieves the same tracking errors as the all-CPU oracle
loop and the reference's band."""
import numpy as np
import pytest

from common import SIZES, load_gait

pytestmark = pytest.mark.gpu

H_MPC, N_SAMPLE, KAPPA = 10, 5, 2.0e-4


def test_gpu_policy_closed_loop_quadruped(cuda_device):
    import torch
    import cimpc_b200 as cb
    from oracle.ip import IPOptions
    from oracle.linearized import linearized_step
    from oracle.newton import get_stride, rot_n_stride, update_window
    from oracle.residual import get_residual
    from oracle.simulator import nonlinear_ip_solve, simulator_ip_options, tracking_error, update_friction_coefficient
    from oracle.trajectory import trajectory_from_gait
    from test_closed_loop import _setup
    from oracle.simulator import simulate

    H_sim, R = 300, 3
    res, m, ref, cpu_policy, gait = _setup(H_sim)
    h = gait["h"]
    nq = m.nq
    # all-CPU oracle loop (nominal initial condition) for comparison
    q1, v1 = ref.q[1].copy(), (ref.q[1] - ref.q[0]) / h
    ok, qc, uc, gc, bc = simulate(res, cpu_policy, q1, v1, H_sim, h / N_SAMPLE, m.mu_world)
    assert ok
    e_cpu = tracking_error(ref, m, qc, uc, gc, bc, N_SAMPLE, idx_shift=(0,))

    # device policy: same linearization (friction-updated reference), same options
    Hr = ref.H
    r0 = np.zeros((Hr, m.nz)); rz0 = np.zeros((Hr, m.nz, m.nz)); rth0 = np.zeros((Hr, m.nz, m.ntheta))
    for t in range(Hr):
        r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], KAPPA)
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, gamma_reg=0.1, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], ref.z, ref.theta, r0, rz0, rth0, mode="configuration", opts=ipo)
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.25] * (nq - 3)), (H_MPC, 1))
    ou = np.tile(3e-2 * np.ones(m.nu), (H_MPC, 1))
    nw = cb.Newton(im, H_MPC, R, oq, ou, KAPPA, cb.NewtonOptions(r_tol=3e-4, max_iter=5))

    rng = np.random.default_rng(40)
    q = np.zeros((R, H_sim + 2, nq)); u = np.zeros((R, H_sim, m.nu)); gam = np.zeros((R, H_sim, m.nc)); b = np.zeros((R, H_sim, m.nb))
    for r in range(R):
        v = v1 * (1.0 + (0.0 if r == 0 else 0.05 * rng.standard_normal()))
        q[r, 1] = q1
        q[r, 0] = q1 - (h / N_SAMPLE) * v
    ptraj, window, stride = ref.copy(), list(range(H_MPC + 2)), get_stride(m, ref)
    q0 = np.tile(ref.q[0], (R, 1))
    u_cur = np.zeros((R, m.nu))
    sim_opts = simulator_ip_options()
    i = res.idx
    cnt = N_SAMPLE
    for t in range(1, H_sim + 1):
        if cnt == N_SAMPLE:  # policy.jl:109-135 for every rollout at once
            q1m = q[:, t].copy()
            uu, _, info = nw.solve(np.array(window, dtype=np.int32), ptraj.q[:H_MPC + 2], ptraj.u[:H_MPC], ref.theta[0, -2],
                                   h, torch.from_numpy(q0).to(cuda_device), torch.from_numpy(q1m).to(cuda_device),
                                   warm_start=t > 1)
            u_cur = uu.cpu().numpy() / N_SAMPLE
            rot_n_stride(ptraj, stride)
            window = update_window(window, Hr)
            q0 = q1m
            cnt = 0
        cnt += 1
        for r in range(R):
            u[r, t - 1] = u_cur[r]
            z = np.ones(i.nz)
            z[i.q2] = q[r, t]
            th = np.concatenate([q[r, t - 1], q[r, t], u[r, t - 1], np.zeros(m.nw), [m.mu_world], [h / N_SAMPLE]])
            okk, z, _ = nonlinear_ip_solve(res, z, th, sim_opts)
            assert okk
            q[r, t + 1], gam[r, t - 1], b[r, t - 1] = z[i.q2], z[i.g1], z[i.b1]
    e_gpu = np.array([tracking_error(ref, m, q[r], u[r], gam[r], b[r], N_SAMPLE, idx_shift=(0,)) for r in range(R)])
    print("CPU-oracle loop :", np.round(e_cpu, 4))
    print("GPU-policy loops:", np.round(e_gpu, 4))
    band = np.array([0.0201, 0.0437, 0.374, 0.0789]) * 1.5  # mpc_quadruped.jl:61-64
    assert (e_gpu[0] < band).all()
    assert np.all(np.abs(e_gpu[0] - e_cpu) <= 0.15 * e_cpu + 1e-3)  # same controller, same plant: same tracking quality
    assert (e_gpu[1:] < 2.0 * band).all()  # perturbed rollouts stay on the gait
