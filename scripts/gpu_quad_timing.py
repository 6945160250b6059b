"""Device-resident ip_solve_kernel throughput on the headline quadruped batch (655 360 subproblems); prints one line."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
robot, H, n = "quadruped", 10, 655360
lin, gait = load_lin(robot), load_gait(robot)
im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration",
                           opts=cb.InteriorPointOptions(diff_sol=True, r_tol=1e-4, kappa_tol=1e-4))
knot, theta, q2 = make_batch(robot, lin, gait, n, seed=1)
stage = (np.arange(n) // (n // H)).astype(np.int32) % lin["z0"].shape[0]
theta = theta - lin["th0"][knot] + lin["th0"][stage]; q2 = q2 - lin["z0"][knot, :11] + lin["z0"][stage, :11]
kd, td, qd = torch.from_numpy(stage).to(dev), torch.from_numpy(np.ascontiguousarray(theta)).to(dev), torch.from_numpy(np.ascontiguousarray(q2)).to(dev)
out = im.solve_device(kd, td, qd)
for _ in range(3): im.solve_device(kd, td, qd, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): im.solve_device(kd, td, qd, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{os.environ.get('CIMPC_B200_LIB', 'default')}: {ms:.3f} ms {n / ms / 1e3:.2f} M/s conv {out[2].float().mean().item():.4f} iters {out[3].float().mean().item():.3f}")
