"""examples/flamingo/flat.jl (= test/controller/mpc_flamingo.jl) on the B200: the flamingo CI-MPC policy
(:configurationforce, TrackingVelocityObjective) and its simulator in closed loop, entirely on the device.

    python examples/flamingo_flat.py [--steps 1000] [--rollouts 8] [--gait path/to/gait_forward_36_4.jld2]

Prints the tracking errors next to the reference's nominal values (mpc_flamingo.jl:72-75).  Product only."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cimpc_b200 as cb  # noqa: E402

SIZES = (9, 6, 2, 4, 8)
H_MPC, N_SAMPLE, KAPPA = 15, 5, 2.0e-4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--rollouts", type=int, default=8)
    ap.add_argument("--gait", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.gait and args.gait.endswith(".jld2"):
        gait = cb.load_gait(args.gait, "split_traj_alt")
    else:
        with np.load(args.gait or os.path.join(ROOT, "tests", "golden", "flamingo_gait.npz")) as f:
            gait = {k: (f[k] if f[k].ndim else float(f[k])) for k in f.files}
    ref = cb.ContactTraj.from_gait("flamingo", gait, kappa=KAPPA)
    nq, nu, nw, nc, nb = SIZES
    ipo = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=KAPPA, undercut=5.0, diff_sol=True)
    im = cb.ImplicitTrajectory(*SIZES, ref.z, ref.theta, kappa=KAPPA, mode="configurationforce", opts=ipo)
    obj_q = np.tile(1e-1 * np.array([3e2, 1e-6, 3e2, 1, 1, 1, 1, 0.1, 0.1]), (H_MPC, 1))   # flat.jl:34-41
    obj_u = np.tile(3e-1 * np.array([0.1, 0.1, 0.3, 0.3, 2, 2]), (H_MPC, 1))
    obj_v = np.tile(1e-3 * np.array([1e0, 1, 1e4, 1, 1, 1, 1, 1e4, 1e4]), (H_MPC, 1))
    sim_opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_ls=25, eps_min=0.05, undercut=float("inf"), gamma_reg=0.0)
    R = args.rollouts
    mc = cb.MonteCarloRollouts(im, ref.q, ref.u, float(gait["mu"]), 1.0, ref.h, H_mpc=H_MPC, N_sample=N_SAMPLE, obj_q=obj_q,
                               obj_u=obj_u, kappa=KAPPA, n_rollouts=R, newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5),
                               sim_opts=sim_opts, obj_gamma=np.full((H_MPC, nc), 1e-100), obj_b=np.full((H_MPC, nb), 1e-100),
                               obj_v=obj_v, ref_gamma=ref.gamma, ref_b=ref.b)
    q1 = np.tile(ref.q[1], (R, 1))
    v1 = np.tile((ref.q[1] - ref.q[0]) / ref.h, (R, 1))
    v1[1:] *= 1.0 + 0.02 * np.random.default_rng(0).standard_normal((R - 1, 1))  # rollout 0 = the nominal initial condition
    out = mc.run(torch.from_numpy(q1).to(dev), torch.from_numpy(v1).to(dev), args.steps)
    torch.cuda.synchronize()
    ok = out["status"].cpu().numpy()
    q, u, gam, b = (out[k].cpu().numpy() for k in ("q", "u", "gamma", "b"))
    print(f"{int(ok.sum())} of {R} rollouts completed {args.steps} simulator steps")
    print("reference (mpc_flamingo.jl:72-75)      q 0.0154  u 0.0829  γ 0.444  b 0.0169")
    for r in range(min(R, 4)):
        if ok[r]:
            e = cb.tracking_error(ref, q[:, r], u[:, r], gam[:, r], b[:, r], N_SAMPLE)
            print(f"rollout {r}{' (nominal)' if r == 0 else '          '}                    q {e[0]:.4f}  u {e[1]:.4f}  γ {e[2]:.3f}  b {e[3]:.4f}")


if __name__ == "__main__":
    main()
