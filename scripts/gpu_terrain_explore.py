"""Flamingo under the flat-linearized CI-MPC policy with altitude updates, simulated ON the piecewise terrain
(examples/flamingo/piecewise.jl with `sim = simulator(s_sim, ...)`): does it climb the ramp?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import cimpc_b200 as cb  # noqa: E402
from common import SIZES, load_gait  # noqa: E402
from oracle.residual import get_residual  # noqa: E402
from oracle.trajectory import trajectory_from_gait  # noqa: E402

robot, H_mpc, N, kappa, R = "flamingo", 15, 5, 1.0e-4, 8
H_sim = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
sim_model = sys.argv[2] if len(sys.argv) > 2 else "flamingo_piecewise"
alt_upd = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
res = get_residual(robot)
tres = get_residual("flamingo_piecewise")
m = res.model
gait = load_gait(robot)
ref = trajectory_from_gait(m, gait)
h = gait["h"]
ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=kappa, undercut=5.0, diff_sol=True)
im = cb.ImplicitTrajectory(*SIZES[robot], ref.z, ref.theta, kappa=kappa, mode="configurationforce", opts=ipo)
oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_mpc, 1))
ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_mpc, 1))
ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_mpc, 1))
sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"), gamma_reg=0.0)
mc = cb.MonteCarloRollouts(im, ref.q, ref.u, ref.theta[0, -2], m.mu_world, h, H_mpc=H_mpc, N_sample=N, obj_q=oq,
                           obj_u=ou, kappa=kappa, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                           sim_opts=sim_opts, obj_gamma=np.full((H_mpc, m.nc), 1e-100),
                           obj_b=np.full((H_mpc, m.nb), 1e-100), obj_v=ov, ref_gamma=ref.gamma, ref_b=ref.b,
                           altitude_update=alt_upd, altitude_impact_threshold=0.02, sim_model=sim_model)
dev = torch.device("cuda", 0)
q1 = np.tile(ref.q[1], (R, 1))
v1 = np.tile((ref.q[1] - ref.q[0]) / h, (R, 1))
v1[1:] *= 1.0 + 0.02 * np.random.default_rng(61).standard_normal((R - 1, 1))
import time
t0 = time.time()
out = mc.run(torch.from_numpy(q1).to(dev), torch.from_numpy(v1).to(dev), H_sim)
torch.cuda.synchronize()
print("wall", time.time() - t0)
ok = out["status"].cpu().numpy()
q = out["q"].cpu().numpy()
print("status", ok, "failed_at", out["failed_at"].cpu().numpy())
for t in range(0, H_sim + 1, max(1, H_sim // 15)):
    x = q[t + 1, 0, 0]
    px = tres._px(q[t + 1, 0])
    phi = [q_ for q_ in res.model.phi_func(list(q[t + 1, 0]))]
    print(f"t={t:5d} x={x:7.3f} z={q[t + 1, 0, 1]:6.3f} terrain(x)={tres.terrain.height(x):6.3f} torso={q[t + 1, 0, 2]:6.3f} "
          f"gap={[round(float(p) - tres.terrain.height(xx), 3) for p, xx in zip(phi, px)]}")
print("alt", out["alt"].cpu().numpy()[0] if out.get("alt") is not None else None)
