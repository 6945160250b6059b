"""One batched newton_solve! workload (the MPC leg of bench.py) on its own: timing, or a target for ncu.
    python scripts/gpu_mpc_solve.py [--rollouts 16384] [--solves 3]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from common import SIZES, load_gait  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts", type=int, default=16384)
    ap.add_argument("--solves", type=int, default=3)
    args = ap.parse_args()
    import torch
    import cimpc_b200 as cb
    cfg = dict(bench.CONFIGS[4], id=4)
    robot, H = cfg["robot"], cfg["H"]
    nq, nu, nw, nc, nb = SIZES[robot]
    lin = bench.build_workload(0, 8, cfg)[0]
    opts = cb.InteriorPointOptions(**cfg["ip"])
    im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                               mode=cfg["mode"], opts=opts, device=0)
    gait = load_gait(robot)
    R = args.rollouts
    oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1))
    ou = np.tile(3e-2 * np.ones(nu), (H, 1))
    newton = cb.Newton(im, H, R, oq, ou, 1.0e-4, cb.NewtonOptions(r_tol=3e-4, max_iter=5), ip_opts=opts)
    rng = np.random.Generator(np.random.Philox(1000))
    dev = torch.device("cuda", 0)
    q0 = torch.from_numpy(np.tile(gait["q"][0], (R, 1))).to(dev)
    q1 = torch.from_numpy(gait["q"][1] + 0.01 * rng.standard_normal((R, nq))).to(dev)
    win = np.arange(H + 2, dtype=np.int32)
    a = (win, gait["q"][:H + 2], gait["u"][:H], gait["mu"], gait["h"], q0, q1)
    u, _, info = newton.solve(*a)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.solves + 1)]
    ev[0].record()
    for i in range(args.solves):
        u, _, info = newton.solve(*a)
        ev[i + 1].record()
    torch.cuda.synchronize()
    each = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.solves)]
    ms = sum(each) / max(args.solves, 1)
    print("per solve ms:", " ".join(f"{x:.2f}" for x in each), "hostloop" if os.environ.get("CIMPC_NEWTON_HOSTLOOP") else "graph")
    print(f"kernel={os.environ.get('CIMPC_NEWTON_KERNEL', 'cta')} rollouts={R} ms_per_solve={ms:.3f} "
          f"mpc_steps_per_s={R / ms * 1e3:.0f} info_mean={info.double().mean(0).cpu().numpy()}")


if __name__ == "__main__":
    main()
