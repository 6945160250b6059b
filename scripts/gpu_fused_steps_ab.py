"""Closed loop of the bench (quadruped Monte-Carlo drop box, 16 384 rollouts, 50 simulator steps = 10 MPC steps) with the
simulator steps launched one by one vs fused per policy interval (cimpc_sim_steps_batch), 1 / 2 / 4 rollout groups."""
import json, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin
dev = torch.device("cuda:0")
robot = "quadruped"; lin, gait = load_lin(robot), load_gait(robot); nq, nu = 11, 8
R = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H, N, T = 10, 5, 50
opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1)); ou = np.tile(3e-2 * np.ones(nu), (H, 1))
def make_im():
    return cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=100)).to(dev)
v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(dev)
ref = None
for groups in (1, 2, 4):
    for fused in (False, True):
        mc = cb.GroupedRollouts(make_im, R, groups, gait["q"], gait["u"], gait["mu"], 1.0, gait["h"], H_mpc=H, N_sample=N, obj_q=oq, obj_u=ou,
                                kappa=1e-4, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5), fused_steps=fused)
        mc.run(q1, v1, T, record_every=N); torch.cuda.synchronize()   # warm-up (graphs, allocations)
        ms = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = mc.run(q1, v1, T, record_every=N); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        o = {k: out[k].cpu().numpy() for k in ("q", "u", "gamma", "b", "status", "failed_at")}
        same = None if ref is None else all(np.array_equal(o[k], ref[k]) for k in o)
        ref = ref or o
        print(json.dumps({"rollouts": R, "groups": groups, "fused_steps": fused, "ms": [round(x, 1) for x in ms],
                          "mpc_steps_per_s": R * mc.mpc_steps / min(ms) * 1e3, "ok_frac": float(o["status"].mean()),
                          "identical_to_first": same}), flush=True)
        del mc
