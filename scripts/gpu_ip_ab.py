"""A/B of ip_solve_kernel (v2, one row per lane) and ip_solve2_kernel (v3, two rows per lane): bitwise comparison of every
output and device-resident timing on each BASELINE config's robot.  CIMPC_IP_KERNEL is read per launch."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
CONFIGS = [  # robot, mode, H_mpc, n, ip options
    ("quadruped", "configuration", 10, 1003, dict(r_tol=1e-4, kappa_tol=1e-4)),
    ("quadruped", "configuration", 10, 65536 * 10, dict(r_tol=1e-4, kappa_tol=1e-4)),
    ("quadruped", "configurationforce", 10, 16384 * 10, dict(r_tol=1e-8, kappa_tol=1e-8)),
    ("flamingo", "configurationforce", 15, 16384 * 15, dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("flamingo", "configuration", 15, 16384 * 15, dict(r_tol=1e-4, kappa_tol=1e-4, max_ls=0)),
    ("centroidal_quadruped", "configuration", 20, 262160, dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("centroidal_quadruped", "configurationforce", 20, 20000, dict(r_tol=1e-8, kappa_tol=2e-4)),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for robot, mode, H, n, kw in CONFIGS:
    if only and only not in robot: continue
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(diff_sol=True, **kw)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=1)
    R = max(n // H, 1)
    stage = (np.arange(n) // R).astype(np.int32) % lin["z0"].shape[0]
    nq = SIZES[robot][0]
    theta = theta - lin["th0"][knot] + lin["th0"][stage]; q2 = q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq]
    kd, td, qd = torch.from_numpy(stage).to(dev), torch.from_numpy(np.ascontiguousarray(theta)).to(dev), torch.from_numpy(np.ascontiguousarray(q2)).to(dev)
    res = {}
    for ver in ("v2", "v3"):
        os.environ["CIMPC_IP_KERNEL"] = ver
        out = im.solve_device(kd, td, qd)
        for _ in range(2): im.solve_device(kd, td, qd, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for _ in range(K): im.solve_device(kd, td, qd, out=out)
        e1.record(); torch.cuda.synchronize()
        res[ver] = (e0.elapsed_time(e1) / K, [o.clone() for o in out])
    a, b = res["v2"][1], res["v3"][1]
    same = [bool(torch.equal(x, y)) for x, y in zip(a, b)]
    nan_eq = [bool(torch.equal(torch.nan_to_num(x.double(), nan=12345.0), torch.nan_to_num(y.double(), nan=12345.0))) for x, y in zip(a, b)]
    dmax = [float((x.double() - y.double()).abs().nan_to_num(nan=0.0).max()) for x, y in zip(a, b)]
    print(json.dumps({"robot": robot, "mode": mode, "n": n, "ms_v2": res["v2"][0], "ms_v3": res["v3"][0],
                      "M_per_s_v2": n / res["v2"][0] / 1e3, "M_per_s_v3": n / res["v3"][0] / 1e3,
                      "speedup": res["v2"][0] / res["v3"][0], "bitwise_equal[z,dz,status,iters]": same, "equal_modulo_nan": nan_eq,
                      "max_abs_diff": dmax, "mean_iters": float(a[3].float().mean()), "converged": float(a[2].float().mean()), **kw}), flush=True)
