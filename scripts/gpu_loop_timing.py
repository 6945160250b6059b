"""Where does the closed loop spend its time?  Per-call CUDA-event timings of newton.solve and sim.step."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin
dev = torch.device("cuda:0")
robot = "quadruped"; lin, gait = load_lin(robot), load_gait(robot); nq, nu = 11, 8
R, H, N = 16384, 10, 5
opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1)); ou = np.tile(3e-2 * np.ones(nu), (H, 1))
mc = cb.MonteCarloRollouts(im, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H, N_sample=N, obj_q=oq, obj_u=ou, kappa=1e-4, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
for label, q1 in (("on-gait", np.tile(gait["q"][1], (R, 1))), ("monte-carlo box", cb.quadruped_initial_configurations(R, seed=100))):
    q1 = torch.from_numpy(q1).to(dev); v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(dev)
    h_sim = gait["h"] / N
    qa = (q1 - h_sim * v1).contiguous(); qb = q1.contiguous()
    mc.ref.reset(); q0 = torch.from_numpy(np.tile(mc.ref.q0[0], (R, 1))).to(dev)
    ok = torch.ones(R, dtype=torch.bool, device=dev); cnt = N
    tn, ts, sw, its = [], [], [], []
    for t in range(1, 31):
        if cnt == N:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); u_mpc, _, info = mc.newton.solve(mc.ref.window, mc.ref.q[:H + 2], mc.ref.u[:H], mc.mu_mpc, gait["h"], q0, qb, warm_start=t > 1, active=ok.to(torch.uint8)); e1.record(); torch.cuda.synchronize()
            tn.append(e0.elapsed_time(e1)); sw.append((mc.newton.last_sweeps, info[:, 1].float().mean().item()))
            u_sim = (u_mpc / N).contiguous(); mc.ref.advance(); q0 = qb; cnt = 0
        cnt += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); q2, gam, b, st, it = mc.sim.step(qa, qb, u_sim, 1.0, h_sim, active=ok.to(torch.uint8)); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); its.append((it.float().mean().item(), int(it.max().item())))
        st = st.bool() | ~ok; ok &= st; q2 = torch.where(ok[:, None], q2, qb); qa, qb = qb, q2
    print(label, "alive", ok.float().mean().item())
    print("  newton ms", [round(x, 1) for x in tn], "sweeps(max, mean)", [(a, round(b_, 1)) for a, b_ in sw])
    print("  sim ms", [round(x, 1) for x in ts])
    print("  sim iters (mean,max)", [(round(a, 1), b_) for a, b_ in its])
