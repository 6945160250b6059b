import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait
from oracle.linearized import linearized_step
from test_closed_loop import _setup, H_MPC, N_SAMPLE, KAPPA
dev = torch.device("cuda:0")
res, m, ref, cpu_policy, gait = _setup(1000)
h, nq = gait["h"], m.nq
Hr = ref.H
r0 = np.zeros((Hr, m.nz)); rz0 = np.zeros((Hr, m.nz, m.nz)); rth0 = np.zeros((Hr, m.nz, m.ntheta))
for t in range(Hr):
    r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], KAPPA)
ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, gamma_reg=0.1, diff_sol=True)
im = cb.ImplicitTrajectory(*SIZES["quadruped"], ref.z, ref.theta, r0, rz0, rth0, mode="configuration", opts=ipo)
oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.25] * (nq - 3)), (H_MPC, 1)); ou = np.tile(3e-2 * np.ones(m.nu), (H_MPC, 1))
R = 8
mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_MPC, N_sample=N_SAMPLE, obj_q=oq, obj_u=ou, kappa=KAPPA, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
rng = np.random.default_rng(60)
q1 = np.tile(ref.q[1], (R, 1)); v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1)); v1[1:] *= 1.0 + 0.05 * rng.standard_normal((R - 1, 1))
# instrumented copy of run()
h_sim = h / N_SAMPLE
qa = torch.from_numpy(q1 - h_sim * v1).to(dev); qb = torch.from_numpy(q1).to(dev)
mc.ref.reset(); q0_mpc = torch.from_numpy(np.tile(mc.ref.q0[0], (R, 1))).to(dev); cnt = N_SAMPLE
dump = {}
for t in range(1, 1001):
    if cnt == N_SAMPLE:
        u_mpc, _, info = mc.newton.solve(mc.ref.window, mc.ref.q[:H_MPC + 2], mc.ref.u[:H_MPC], mc.mu_mpc, h, q0_mpc, qb, warm_start=t > 1)
        u_sim = (u_mpc / N_SAMPLE).contiguous(); mc.ref.advance(); q0_mpc = qb; cnt = 0
    cnt += 1
    q2, gam, b, st, it = mc.sim.step(qa, qb, u_sim, mc.mu_sim, h_sim)
    if not bool(st.all()):
        bad = (~st.bool()).nonzero().flatten().tolist()
        print("step", t, "failed rollouts", bad, "iters", it[bad].tolist())
        dump = dict(t=t, q0=qa.cpu().numpy(), q1=qb.cpu().numpy(), u=u_sim.cpu().numpy(), q2=q2.cpu().numpy(), st=st.cpu().numpy(), it=it.cpu().numpy(), mu=mc.mu_sim, h=h_sim)
        break
    qa, qb = qb, q2
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "sim_fail.npz"), **dump)
print("max iters seen", int(it.max()))
