"""Device-resident throughput of ip_solve_kernel on every BASELINE config's robot (one JSON line each)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
CONFIGS = [  # robot, mode, H_mpc, n, ip options  (BASELINE.json configs 1-5)
    ("hopper_2D", "configuration", 10, 65536 * 10, dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("quadruped", "configuration", 10, 4096 * 10, dict(r_tol=1e-4, kappa_tol=1e-4)),
    ("quadruped", "configuration", 10, 65536 * 10, dict(r_tol=1e-4, kappa_tol=1e-4)),
    ("flamingo", "configurationforce", 15, 16384, dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("flamingo", "configurationforce", 15, 16384 * 15, dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("centroidal_quadruped", "configuration", 20, 262144, dict(r_tol=1e-8, kappa_tol=2e-4)),
]
for robot, mode, H, n, kw in CONFIGS:
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(diff_sol=True, **kw)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=1)
    R = n // H
    stage = (np.arange(n) // max(R, 1)).astype(np.int32) % lin["z0"].shape[0]
    nq = SIZES[robot][0]
    theta = theta - lin["th0"][knot] + lin["th0"][stage]; q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
    kd, td, qd = torch.from_numpy(stage).to(dev), torch.from_numpy(np.ascontiguousarray(theta)).to(dev), torch.from_numpy(np.ascontiguousarray(q2)).to(dev)
    out = im.solve_device(kd, td, qd)
    for _ in range(3): im.solve_device(kd, td, qd, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    e0.record()
    for _ in range(K): im.solve_device(kd, td, qd, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    nu_, nw_, nc_, nb_ = SIZES[robot][1:]
    nth = 2 * nq + nu_ + nw_ + 2; nd = nq if mode == "configuration" else nq + nc_ + nb_
    B = 8 * (nth + nq + nc_) + 8 * (nd + nd * (2 * nq + nu_)) + 5
    nx, ny, mit = nq, 2 * nc_ + nb_, out[3].float().mean().item()
    f_it = 2 * ny ** 3 + ny ** 2 + 4 * (2 * nx * ny + 1.5 * ny ** 2 + nx ** 2) + 6 * (nx ** 2 + 2 * nx * ny + ny ** 2)
    f_d = 2 * ny ** 3 + (2 * nq + nu_) * 2 * (2 * nx * ny + 1.5 * ny ** 2 + nx ** 2)
    flops = mit * f_it + f_d  # BASELINE.md §3 (reference algorithm's flop count)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    f64 = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["dfma_tflops"]
    print(json.dumps({"robot": robot, "mode": mode, "H_mpc": H, "subproblems": n, "ms": ms, "subproblems_per_s": n / ms * 1e3,
                      "algorithmic_bytes": B, "achieved_GBps": n * B / ms / 1e6, "hbm_frac": n * B / ms / 1e6 / peaks["hbm_gbs"],
                      "fp64_tflops_reference_count": n * flops / ms / 1e9, "fp64_frac": n * flops / ms / 1e9 / f64, "group_lanes": im.group,
                      "mean_iters": out[3].float().mean().item(), "converged": out[2].float().mean().item(), **kw}))
