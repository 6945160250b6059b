"""Closed loop (bench setting: drop box, 16 384 rollouts, 50 simulator steps): CUDA-event time of every newton.solve and
every fused cimpc_sim_steps_batch call, Newton iteration / sweep statistics per MPC step."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin
dev = torch.device("cuda:0")
robot = "quadruped"; lin, gait = load_lin(robot), load_gait(robot); nq, nu = 11, 8
R, H, N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 10, 5
opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1)); ou = np.tile(3e-2 * np.ones(nu), (H, 1))
mc = cb.MonteCarloRollouts(im, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H, N_sample=N, obj_q=oq, obj_u=ou, kappa=1e-4, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
for label, q1 in (("monte-carlo box", cb.quadruped_initial_configurations(R, seed=100)), ("on-gait", np.tile(gait["q"][1], (R, 1)))):
    for rep in range(2):
        q1t = torch.from_numpy(q1).to(dev); v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(dev)
        h_sim = gait["h"] / N
        qa = (q1t - h_sim * v1).contiguous(); qb = q1t.contiguous()
        mc.ref.reset(); q0 = torch.from_numpy(np.tile(mc.ref.q0[0], (R, 1))).to(dev)
        ok = torch.ones(R, dtype=torch.bool, device=dev)
        tn, ts, sw, its = [], [], [], []
        for k in range(10):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            okb = ok.to(torch.uint8)
            e0.record()
            u_mpc, _, info = mc.newton.solve(mc.ref.window, mc.ref.q[:H + 2], mc.ref.u[:H], mc.mu_mpc, gait["h"], q0, qb, warm_start=k > 0, active=okb)
            e1.record()
            u_sim = (u_mpc / N).contiguous(); mc.ref.advance(); q0 = qb
            q2s, gs, bs, sts, it, phis = mc.sim.steps(N, qa, qb, u_sim, 1.0, h_sim, active=okb)
            e2.record(); torch.cuda.synchronize()
            tn.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
            act = ok
            sw.append((round(info[act, 0].float().mean().item(), 2), round(info[act, 1].float().mean().item(), 1), int(info[act, 1].max())))
            its.append((round(it[:, act].float().mean().item(), 1), int(it.max())))
            ok = ok & sts[N - 1].bool()
            qa, qb = q2s[N - 2].contiguous(), q2s[N - 1].contiguous()
    print(label, "alive", ok.float().mean().item(), "total ms", round(sum(tn) + sum(ts), 1))
    print("  newton ms", [round(x, 1) for x in tn], "Σ", round(sum(tn), 1))
    print("  newton (mean iterations, mean sweeps, max sweeps)", sw)
    print("  5 sim steps ms", [round(x, 1) for x in ts], "Σ", round(sum(ts), 1))
    print("  sim iters (mean,max)", its)
