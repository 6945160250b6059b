"""Print the parity numbers of one strict-parity configuration (debug aid): python scripts/gpu_parity_debug.py [robot] [mode]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
from oracle.c_oracle import COracle
from oracle.ip import IPOptions
robot = sys.argv[1] if len(sys.argv) > 1 else "quadruped"
mode = sys.argv[2] if len(sys.argv) > 2 else "configuration"
lin, gait = load_lin(robot), load_gait(robot)
for kw in (dict(r_tol=1e-4, kappa_tol=1e-4), dict(r_tol=1e-8, kappa_tol=2e-4), dict(r_tol=1e-8, kappa_tol=1e-8)):
    opts = cb.InteriorPointOptions(diff_sol=True, max_ls=0, **kw)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
    n = 20 * lin["z0"].shape[0] + 7
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    oo = IPOptions(r_tol=opts.r_tol, kappa_tol=opts.kappa_tol, max_iter=opts.max_iter, max_ls=0, diff_sol=True)
    zo, dzo, sto, ito = COracle(*SIZES[robot], lin, mode=mode, solver="lu").solve(knot, theta, q2, oo)
    ez = np.abs(z - zo).max(axis=1) / np.maximum(1.0, np.abs(zo).max(axis=1))
    edz = np.abs(dz - dzo).max(axis=(1, 2)) / np.maximum(1.0, np.abs(dzo).max(axis=(1, 2)))
    same = it == ito
    print(robot, mode, kw, "status eq", np.array_equal(st, sto), "conv", st.mean(), sto.mean(), "iters same", same.mean(),
          "ez max/median(same)", ez[same].max() if same.any() else None, np.median(ez), "edz max/median", edz[same].max() if same.any() else None, np.median(edz),
          "mean iters gpu/oracle", it.mean(), ito.mean())
    nq = SIZES[robot][0]
    i = int(np.argmax(ez))
    print("   worst problem", i, "iters", it[i], ito[i], "dz err", edz[i], "\n   z gpu", np.round(z[i], 6), "\n   z ora", np.round(zo[i], 6))
