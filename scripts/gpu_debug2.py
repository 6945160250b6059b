"""Scratch: device-entry IP solve at several batch sizes (convergence / iteration statistics)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import *
robot = "quadruped"
lin, gait = load_lin(robot), load_gait(robot)
opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
dev = torch.device("cuda:0")
for nb in [int(a) for a in sys.argv[1:]] or (64, 4096, 32768, 40960, 163840, 655360):
    knot, theta, q2 = make_batch(robot, lin, gait, nb, seed=1)
    for srt in (False, True):
        if srt:
            o = np.argsort(knot, kind="stable"); knot, theta, q2 = knot[o], theta[o], q2[o]
        kd, td, qd = torch.from_numpy(knot).to(dev), torch.from_numpy(theta).to(dev), torch.from_numpy(q2).to(dev)
        z, dz, st, it = im.solve_device(kd, td, qd)
        torch.cuda.synchronize()
        print(f"n={nb} sorted={srt}: conv {st.float().mean().item():.4f} iters {it.float().mean().item():.2f} max {it.max().item()} nan {torch.isnan(z).any().item()}", flush=True)
