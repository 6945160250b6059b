"""Solve seeded batches on every robot / mode with the library CIMPC_B200_LIB points at and save all outputs (npz), so
that two builds can be compared bit for bit.  usage: gpu_ip_dump.py OUT.npz"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
out = {}
for robot, mode, n, kw in (("quadruped", "configuration", 65536, dict(r_tol=1e-4, kappa_tol=1e-4)),
                           ("quadruped", "configurationforce", 16384, dict(r_tol=1e-8, kappa_tol=1e-8)),
                           ("flamingo", "configurationforce", 32768, dict(r_tol=1e-8, kappa_tol=2e-4)),
                           ("flamingo", "configuration", 16384, dict(r_tol=1e-8, kappa_tol=1e-8, max_ls=0)),
                           ("centroidal_quadruped", "configuration", 32768, dict(r_tol=1e-8, kappa_tol=2e-4)),
                           ("hopper_2D", "configuration", 16384, dict(r_tol=1e-8, kappa_tol=2e-4))):
    lin, gait = load_lin(robot), load_gait(robot)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                               opts=cb.InteriorPointOptions(diff_sol=True, **kw))
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=21)
    kd, td, qd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (knot, theta, q2))
    z, dz, st, it = im.solve_device(kd, td, qd)
    torch.cuda.synchronize()
    for name, t in (("z", z), ("dz", dz), ("st", st), ("it", it)):
        out[f"{robot}/{mode}/{name}"] = t.cpu().numpy()
np.savez(sys.argv[1], **out)
print("saved", sys.argv[1], len(out))
