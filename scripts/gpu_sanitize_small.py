"""Tiny invocation of every kernel family (target of compute-sanitizer racecheck / synccheck / initcheck, which run
50-500x slower than native): IP solves on every robot size and mode, the three Newton kernels (CTA-per-rollout,
general, dense weights) through the one-graph MPC step, the warp-per-rollout Newton kernel, the simulator step with a
straggler (thread-per-trial line search), the device linearization.  Prints one line per family; results are compared
with a second identical call (determinism) so that a race that changes a result is seen even without the tool."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb
from common import SIZES, load_gait, load_lin, make_batch
dev = torch.device("cuda:0")
which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"ip", "newton", "sim", "lin"}

def same(a, b):
    return all(torch.equal(torch.nan_to_num(x.double(), nan=7.0), torch.nan_to_num(y.double(), nan=7.0)) for x, y in zip(a, b))

if "ip" in which:
    for robot, mode, n in (("quadruped", "configuration", 70), ("quadruped", "configurationforce", 40), ("flamingo", "configurationforce", 40),
                           ("centroidal_quadruped", "configuration", 12), ("hopper_2D", "configuration", 40)):
        lin, gait = load_lin(robot), load_gait(robot)
        opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-6, diff_sol=True)
        im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
        knot, theta, q2 = make_batch(robot, lin, gait, n, seed=3)
        kd, td, qd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (knot, theta, q2))
        o1 = [t.clone() for t in im.solve_device(kd, td, qd)]
        o2 = im.solve_device(kd, td, qd)
        torch.cuda.synchronize()
        print(f"ip {robot} {mode}: converged {float(o1[2].float().mean()):.3f} deterministic {same(o1, o2)}", flush=True)
        im.close()

if "newton" in which:
    robot = "quadruped"
    nq, nu, nw, nc, nb = SIZES[robot]
    lin, gait = load_lin(robot), load_gait(robot)
    H, R = 10, 6
    for label, env, kw in (("cta", {}, {}), ("warp", {"CIMPC_NEWTON_KERNEL": "warp"}, {}),
                           ("general+velocity", {}, {"obj_v": np.tile(1e-3 * np.ones(nq), (H, 1))}),
                           ("dense", {}, {"dense": True})):
        os.environ.pop("CIMPC_NEWTON_KERNEL", None)
        os.environ.update(env)
        opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, diff_sol=True)
        im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode="configuration", opts=opts)
        oq = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H, 1))
        ou = np.tile(3e-2 * np.ones(nu), (H, 1))
        if kw.pop("dense", False):
            oq = np.stack([np.diag(q) + 1e-4 * np.ones((nq, nq)) for q in oq])
        newton = cb.Newton(im, H, R, oq, ou, 1.0e-4, cb.NewtonOptions(r_tol=3e-4, max_iter=3), ip_opts=opts, **kw)
        rng = np.random.Generator(np.random.Philox(5))
        q0 = torch.from_numpy(np.tile(gait["q"][0], (R, 1))).to(dev)
        q1 = torch.from_numpy(gait["q"][1] + 0.01 * rng.standard_normal((R, nq))).to(dev)
        win = np.arange(H + 2, dtype=np.int32)
        a = (win, gait["q"][:H + 2], gait["u"][:H], gait["mu"], gait["h"], q0, q1)
        u1, _, i1 = newton.solve(*a)
        u1 = u1.clone(); i1 = i1.clone()
        u2, _, i2 = newton.solve(*a)
        torch.cuda.synchronize()
        print(f"newton {label}: iterations {i1[:, 0].tolist() if i1.ndim > 1 else i1.tolist()} deterministic {same([u1, i1], [u2, i2])}", flush=True)
        im.close()
    os.environ.pop("CIMPC_NEWTON_KERNEL", None)

if "sim" in which:
    for robot in ("quadruped", "hopper_2D"):
        nq, nu, nw, nc, nb = SIZES[robot]
        gait = load_gait(robot)
        sim = cb.Simulator(nq, nu, nw, nc, nb)
        R = 40
        rng = np.random.Generator(np.random.Philox(9))
        q0n = np.tile(gait["q"][0], (R, 1)); q1n = np.tile(gait["q"][1], (R, 1))
        q1n[3:, 1] += 0.05 * rng.random(R - 3)       # lifted: free fall
        q1n[0, 2:] += 0.3 * rng.standard_normal(nq - 2)  # a violent one: back-tracking / straggler
        q0, q1 = torch.from_numpy(q0n).to(dev), torch.from_numpy(q1n).to(dev)
        u = torch.from_numpy(np.tile(gait["u"][0], (R, 1))).to(dev)
        o1 = [t.clone() for t in sim.step(q0, q1, u, gait["mu"], gait["h"])]
        o2 = sim.step(q0, q1, u, gait["mu"], gait["h"])
        torch.cuda.synchronize()
        print(f"sim {robot}: ok {float(o1[3].float().mean()):.3f} max iterations {int(o1[4].max())} deterministic {same(o1, o2)}", flush=True)
        sim.close()

if "lin" in which:
    robot = "flamingo"
    lin = load_lin(robot)
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], mode="configuration")
    im.linearize(lin["z0"], lin["th0"], 0.0)
    r0, rz0, rth0 = im.get_linearization()
    print("linearize flamingo: max |rz0 - fixture|", float(np.abs(rz0 - lin["rz0"]).max()), flush=True)
    im.close()
