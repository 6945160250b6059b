"""Batched newton_solve! (cold start) throughput for the kernel variants: one JSON line per case."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ONLY = os.environ.get("CIMPC_ONLY")
CASES = [("quadruped", "configuration", False, 10, 1e-4, dict(r_tol=1e-4, kappa_tol=1e-4)),
         ("quadruped", "configuration", True, 10, 1e-4, dict(r_tol=1e-4, kappa_tol=1e-4)),
         ("flamingo", "configurationforce", True, 15, 2e-4, dict(r_tol=1e-8, kappa_tol=2e-4)),
         ("centroidal_quadruped", "configuration", False, 20, 2e-4, dict(r_tol=1e-8, kappa_tol=2e-4))]
for robot, mode, vel, H, kappa, kw in CASES:
    if ONLY and robot != ONLY:
        continue
    lin, gait = load_lin(robot), load_gait(robot)
    nq, nu, nw, nc, nb = SIZES[robot]
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode,
                               opts=cb.InteriorPointOptions(diff_sol=True, max_iter=100, **kw))
    rng = np.random.default_rng(1)
    if robot == "flamingo":
        oq = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H, 1)); ou = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H, 1))
        ov = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H, 1))
    else:
        oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1)); ou = np.tile(3e-2 * np.ones(nu), (H, 1))
        ov = np.tile(1e-3 * np.ones(nq), (H, 1))
    force = mode == "configurationforce"
    nwt = cb.Newton(im, H, R, oq, ou, kappa, cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                    obj_gamma=np.full((H, nc), 1e-100) if force else None, obj_b=np.full((H, nb), 1e-100) if force else None,
                    obj_v=ov if vel else None)
    q0 = torch.from_numpy(np.tile(gait["q"][0], (R, 1))).to(dev)
    # the centroidal stand-in-place gait needs a larger push before the first residual exceeds the Newton tolerance
    sigma = float(os.environ.get("CIMPC_SIGMA", "0.05" if robot == "centroidal_quadruped" else "0.005"))
    q1 = torch.from_numpy(gait["q"][1] + sigma * rng.standard_normal((R, nq))).to(dev)
    window = np.arange(H + 2, dtype=np.int32)
    mu = float(lin["th0"][0, -2])
    args = (window, gait["q"][:H + 2], gait["u"][:H], mu, gait["h"], q0, q1)
    kwargs = dict(ref_gamma=gait["gamma"][:H], ref_b=gait["b"][:H]) if force else {}
    for _ in range(2): nwt.solve(*args, **kwargs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 3
    e0.record()
    for _ in range(K): u, q, info = nwt.solve(*args, **kwargs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    info = info.cpu().numpy()
    print(json.dumps({"robot": robot, "mode": mode, "velocity_objective": vel, "H_mpc": H, "rollouts": R, "ms_per_batch": ms,
                      "mpc_steps_per_s": R / ms * 1e3, "mean_newton_iterations": float(info[:, 0].mean()),
                      "mean_sweeps": float(info[:, 1].mean()), "converged": float(info[:, 2].mean()),
                      "kernel": "specialised" if (mode == "configuration" and not vel) else "general"}))
