"""Where does a Monte-Carlo drop-box simulator step spend its time?  Same inputs, varying max_iter / max_ls."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait
dev = torch.device("cuda:0")
robot = "quadruped"; gait = load_gait(robot); nq, nu = 11, 8
R, N = 16384, 5
h_sim = gait["h"] / N
q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=100)).to(dev)
v1 = torch.from_numpy(np.tile((gait["q"][1] - gait["q"][0]) / gait["h"], (R, 1))).to(dev)
qa = (q1 - h_sim * v1).contiguous(); qb = q1.contiguous()
u = torch.from_numpy(np.tile(gait["u"][0] / N, (R, 1))).to(dev)
for max_iter, max_ls in ((100, 25), (50, 25), (25, 25), (12, 25), (100, 0), (100, 3)):
    o = cb.simulator_options(); o.max_iter = max_iter; o.max_ls = max_ls
    sim = cb.Simulator(*SIZES[robot], opts=o)
    sim.step(qa, qb, u, 1.0, h_sim); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); q2, g, b, st, it = sim.step(qa, qb, u, 1.0, h_sim); e1.record(); torch.cuda.synchronize()
    itc = it.cpu().numpy(); stc = st.cpu().numpy()
    tiles = itc.reshape(-1, 32).max(axis=1)
    print(f"max_iter {max_iter:3d} max_ls {max_ls:2d}: {e0.elapsed_time(e1):6.2f} ms  conv {stc.mean():.4f} iters mean {itc.mean():.1f} "
          f"hist(>=25,>=50,>=100) {(itc >= 25).sum()},{(itc >= 50).sum()},{(itc >= 100).sum()}  tiles with max>=100: {(tiles >= 100).sum()} of {len(tiles)}; sum of tile maxima {tiles.sum()}")
