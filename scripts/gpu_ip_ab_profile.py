"""One launch of each ip_solve kernel version on the same quadruped batch (for ncu: -k regex:ip_solve)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
robot = sys.argv[1] if len(sys.argv) > 1 else "quadruped"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 81920
H = 10
lin, gait = load_lin(robot), load_gait(robot)
opts = cb.InteriorPointOptions(diff_sol=True, r_tol=1e-4, kappa_tol=1e-4)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
knot, theta, q2 = make_batch(robot, lin, gait, n, seed=1)
R = max(n // H, 1)
stage = (np.arange(n) // R).astype(np.int32) % lin["z0"].shape[0]
nq = SIZES[robot][0]
theta = theta - lin["th0"][knot] + lin["th0"][stage]; q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
kd, td, qd = torch.from_numpy(stage).to(dev), torch.from_numpy(np.ascontiguousarray(theta)).to(dev), torch.from_numpy(np.ascontiguousarray(q2)).to(dev)
for ver in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("v2", "v3")):
    os.environ["CIMPC_IP_KERNEL"] = ver
    out = im.solve_device(kd, td, qd)
    torch.cuda.synchronize()
print("done")
