"""Closed loop (Monte-Carlo drop box, 16 384 rollouts, 50 simulator steps) as 1 / 2 / 4 / 8 independent parts."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin
dev = torch.device("cuda:0")
robot = "quadruped"; lin, gait = load_lin(robot), load_gait(robot); nq, nu = 11, 8
R, H, N = 16384, 10, 5
opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1)); ou = np.tile(3e-2 * np.ones(nu), (H, 1))
make_im = lambda: cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=100)).to(dev)
v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(dev)
ref = None
for G in (1, 2, 4, 8):
    mc = cb.GroupedRollouts(make_im, R, G, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H, N_sample=N, obj_q=oq, obj_u=ou,
                            kappa=1e-4, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
    mc.run(q1, v1, N); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = mc.run(q1, v1, 50, record_every=N); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    q = out["q"].cpu().numpy(); ok = out["status"].cpu().numpy()
    if ref is None: ref = (q, ok)
    same = np.array_equal(ok, ref[1]) and np.array_equal(q, ref[0])
    print(f"groups {G}: {ms:.1f} ms  {R * mc.mpc_steps / ms:.1f} k MPC steps/s  ok {ok.mean():.4f}  identical to 1 group: {same}", flush=True)
