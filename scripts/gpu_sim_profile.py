"""Three simulator-step launches for ncu: two on-gait steps and one step from the Monte-Carlo drop box."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait
dev = torch.device("cuda:0")
robot = "quadruped"; gait = load_gait(robot); nq, nu = 11, 8
R = int(os.environ.get("R", 16384)); N = 5
sim = cb.Simulator(*SIZES[robot])
h = gait["h"] / N
v1 = (gait["q"][1] - gait["q"][0]) / gait["h"]
u = torch.from_numpy(np.tile(gait["u"][0] / N, (R, 1))).to(dev)
for label, q1 in (("on-gait", np.tile(gait["q"][1], (R, 1))), ("box", cb.quadruped_initial_configurations(R, seed=100))):
    qb = torch.from_numpy(q1).to(dev); qa = (qb - h * torch.from_numpy(v1).to(dev)).contiguous()
    for t in range(2 if label == "on-gait" else 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); q2, gam, b, st, it = sim.step(qa, qb, u, 1.0, h); e1.record(); torch.cuda.synchronize()
        print(label, t, "ms", round(e0.elapsed_time(e1), 2), "iters mean/max", it.float().mean().item(), it.max().item(), "ok", st.float().mean().item())
        qa, qb = qb, q2
