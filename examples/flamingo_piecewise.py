"""examples/flamingo/piecewise.jl on the B200, with the plant ON the terrain: the flamingo CI-MPC policy is linearized on
flat ground (:configurationforce, TrackingVelocityObjective) and learns the ground height under each foot from
`update_altitude!` (`altitude_update = true`, `altitude_impact_threshold = 0.02`); the simulator is the
`flamingo_piecewise` model (`piecewise1_2D_lc`: flat, a 10° ramp from x = 0.5, kinks smoothed; `approx` Jacobians).

    python examples/flamingo_piecewise.py [--steps 2500] [--rollouts 8] [--no-altitude-update]

Prints where the nominal rollout is on the terrain every 250 steps.  Product only."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cimpc_b200 as cb  # noqa: E402
from cimpc_b200 import package  # noqa: E402

SIZES = (9, 6, 2, 4, 8)
H_MPC, N_SAMPLE, KAPPA = 15, 5, 1.0e-4  # piecewise.jl:24-29


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2500)  # H_sim, piecewise.jl:28
    ap.add_argument("--rollouts", type=int, default=8)
    ap.add_argument("--no-altitude-update", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    with np.load(os.path.join(ROOT, "tests", "golden", "flamingo_gait.npz")) as f:
        gait = {k: (f[k] if f[k].ndim else float(f[k])) for k in f.files}
    ref = cb.ContactTraj.from_gait("flamingo", gait, kappa=KAPPA)
    nq, nu, nw, nc, nb = SIZES
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES, ref.z, ref.theta, kappa=KAPPA, mode="configurationforce", opts=ipo)  # flat ground
    obj_q = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_MPC, 1))   # piecewise.jl:31-36
    obj_u = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_MPC, 1))
    obj_v = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_MPC, 1))
    sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"), gamma_reg=0.0)
    R = args.rollouts
    mc = cb.MonteCarloRollouts(im, ref.q, ref.u, float(gait["mu"]), 1.0, ref.h, H_mpc=H_MPC, N_sample=N_SAMPLE, obj_q=obj_q,
                               obj_u=obj_u, kappa=KAPPA, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                               sim_opts=sim_opts, obj_gamma=np.full((H_MPC, nc), 1e-100), obj_b=np.full((H_MPC, nb), 1e-100),
                               obj_v=obj_v, ref_gamma=ref.gamma, ref_b=ref.b,
                               altitude_update=not args.no_altitude_update, altitude_impact_threshold=0.02,  # piecewise.jl:43-47
                               sim_model="flamingo_piecewise")
    q1 = np.tile(ref.q[1], (R, 1))
    v1 = np.tile((ref.q[1] - ref.q[0]) / ref.h, (R, 1))
    v1[1:] *= 1.0 + 0.02 * np.random.default_rng(0).standard_normal((R - 1, 1))  # rollout 0 = the nominal initial condition
    out = mc.run(torch.from_numpy(q1).to(dev), torch.from_numpy(v1).to(dev), args.steps)
    torch.cuda.synchronize()
    ok = out["status"].cpu().numpy()
    q = out["q"].cpu().numpy()
    terrain = package.modelgen.terrain.get_terrain("piecewise1_2D_lc") if hasattr(package, "modelgen") else None
    if terrain is None:
        import importlib
        terrain = importlib.import_module(package.__name__ + ".modelgen.terrain").get_terrain("piecewise1_2D_lc")
    print(f"{int(ok.sum())} of {R} rollouts completed {args.steps} simulator steps; failed at {out['failed_at'].cpu().numpy()}")
    for t in range(0, args.steps + 1, 250):
        x, z = q[t + 1, 0, 0], q[t + 1, 0, 1]
        print(f"step {t:5d}: body x = {x:6.3f} m, z = {z:6.3f} m, ground under the body = {terrain.surf(x):6.3f} m")
    if out.get("alt") is not None:
        print("altitudes the policy plans with (last update, nominal rollout):", np.round(out["alt"].cpu().numpy()[0], 3))


if __name__ == "__main__":
    main()
