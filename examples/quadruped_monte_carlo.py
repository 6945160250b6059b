"""examples/quadruped/monte_carlo.jl on the B200: CI-MPC closed loop of many quadruped rollouts, entirely on the device.

    python examples/quadruped_monte_carlo.py [--rollouts 4096] [--steps 250] [--gait path/to/gait2.jld2] [--groups 2]

Reference trajectory -> device linearization (`ImplicitTrajectory(ref_traj, s; κ)`) -> `ci_mpc_policy` (batched
`newton_solve!`) + simulator (`step!`) for every rollout, initial configurations drawn from the box of
monte_carlo.jl:76-116; prints the survival rate, the tracking errors of the surviving rollouts (trajectory.jl:188-217)
and the throughput.  Uses only the product (`cimpc_b200`): no oracle, no reference tree."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cimpc_b200 as cb  # noqa: E402

SIZES = (11, 8, 2, 4, 8)  # nq, nu, nw, nc, nb of the quadruped
H_MPC, N_SAMPLE, KAPPA = 10, 5, 1.0e-4  # monte_carlo.jl:27-31


def load_reference(path):
    if path and path.endswith(".jld2"):
        return cb.load_gait(path, "split_traj_alt")
    with np.load(path or os.path.join(ROOT, "tests", "golden", "quadruped_gait.npz")) as f:
        return {k: (f[k] if f[k].ndim else float(f[k])) for k in f.files}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=250, help="simulator steps (N_sample = 5 per MPC step)")
    ap.add_argument("--gait", default=None, help="gait2.jld2 of the reference, or an .npz with the same arrays")
    ap.add_argument("--groups", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    gait = load_reference(args.gait)
    ref = cb.ContactTraj.from_gait("quadruped", gait, kappa=KAPPA)
    nq, nu = SIZES[0], SIZES[1]
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)  # monte_carlo.jl:51-58

    def make_im():  # `ImplicitTrajectory(ref_traj, s; κ = κ_mpc, opts)`, linearized on the device
        return cb.ImplicitTrajectory(*SIZES, ref.z, ref.theta, kappa=KAPPA, mode="configuration", opts=opts)

    obj_q = np.tile(1e-2 * np.array([1.0, 0.02, 0.25] + [0.75] * (nq - 3)), (H_MPC, 1))  # monte_carlo.jl:33-37
    obj_u = np.tile(3e-2 * np.ones(nu), (H_MPC, 1))
    R = args.rollouts
    mc = cb.GroupedRollouts(make_im, R, args.groups, ref.q, ref.u, float(gait["mu"]), 1.0, ref.h, H_mpc=H_MPC,
                            N_sample=N_SAMPLE, obj_q=obj_q, obj_u=obj_u, kappa=KAPPA,
                            newton_opts=cb.NewtonOptions(r_tol=3e-4, max_iter=5))
    q1 = torch.from_numpy(cb.quadruped_initial_configurations(R, seed=100)).to(dev)
    v1 = torch.from_numpy(np.tile((ref.q[1] - ref.q[0]) / ref.h, (R, 1))).to(dev)
    mc.run(q1, v1, N_SAMPLE)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = mc.run(q1, v1, args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = out["status"].cpu().numpy()
    q, u, gam, b = (out[k].cpu().numpy() for k in ("q", "u", "gamma", "b"))
    idx = np.flatnonzero(ok)[:64]
    err = np.array([cb.tracking_error(ref, q[:, r], u[:, r], gam[:, r], b[:, r], N_SAMPLE) for r in idx])
    print(f"{R} rollouts x {args.steps} simulator steps ({mc.mpc_steps} MPC steps each) in {dt:.2f} s: "
          f"{R * mc.mpc_steps / dt / 1e3:.1f} k MPC steps/s, {R * args.steps / dt / 1e6:.2f} M simulator steps/s")
    print(f"survived the drop: {ok.mean() * 100:.1f} %")
    if len(idx):
        print("tracking errors of surviving rollouts (median q, u, γ, b):", np.round(np.median(err, axis=0), 4))


if __name__ == "__main__":
    main()
