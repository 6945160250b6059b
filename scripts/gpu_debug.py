"""Scratch GPU run: dump CUDA vs oracle results for offline analysis + first timing."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import *
from oracle.c_oracle import COracle
from oracle.ip import IPOptions
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = {}
for robot, mode, kw in [("quadruped", "configuration", dict(r_tol=1e-4, kappa_tol=1e-4)),
                        ("quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),
                        ("flamingo", "configurationforce", dict(r_tol=1e-8, kappa_tol=2e-4)),
                        ("centroidal_quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),]:
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(diff_sol=True, **kw)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
    n = 2000
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    co = COracle(*SIZES[robot], lin, mode=mode)
    oo = IPOptions(diff_sol=True, **kw)
    zc, dzc, stc, itc = co.solve(knot, theta, q2, oo)
    tag = f"{robot}_{mode}_{kw['r_tol']}"
    same = it == itc
    sel = same & stc
    print(tag, "conv", st.mean(), stc.mean(), "status eq", np.array_equal(st, stc), "iter mismatch", (~same).mean(),
          "z err", np.abs(z - zc)[sel].max(), "dz relerr", (np.abs(dz - dzc)[sel].max() / np.abs(dzc).max()))
    out[tag + "_z"] = z; out[tag + "_zc"] = zc; out[tag + "_it"] = it; out[tag + "_itc"] = itc
    out[tag + "_st"] = st; out[tag + "_stc"] = stc
    # timing
    dev = torch.device("cuda:0")
    for nb in (4096 * 10, 65536 * 10):
        knot, theta, q2 = make_batch(robot, lin, gait, nb, seed=1)
        kd, td, qd = torch.from_numpy(knot).to(dev), torch.from_numpy(theta).to(dev), torch.from_numpy(q2).to(dev)
        outb = im.solve_device(kd, td, qd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            im.solve_device(kd, td, qd, out=outb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"   n={nb}: {ms:.2f} ms  -> {nb / ms * 1e3 / 1e6:.2f} M subproblems/s; conv {outb[2].float().mean().item():.4f} iters {outb[3].float().mean().item():.2f}")
    t0 = time.time(); co.solve(knot[:40000], theta[:40000], q2[:40000], oo); dt = time.time() - t0
    print(f"   C oracle: {40000 / dt:.0f} subproblems/s on {co.max_threads} threads")
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "debug1.npz"), **out)
