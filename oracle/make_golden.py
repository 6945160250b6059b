"""Regenerate the committed fixtures under tests/golden/ (run in the BUILD container only).

TEST INFRASTRUCTURE ONLY.  Reads the reference's gait files from /root/reference (read-only) with
`oracle/jld2_gait.py`, evaluates the oracle's nonlinear residual / Jacobians (`oracle/residual.py`)
at every reference knot — the `LinearizedStep` data of `ImplicitTrajectory`
(src/controller/implicit_dynamics.jl:56, linearized_step.jl:10-29) — and stores

  tests/golden/<robot>_gait.npz   q, u, gamma, b, psi, eta, mu, h     (datasets qm, um, γm, bm, ψm, ηm, μm, hm)
  tests/golden/<robot>_lin.npz    z0, th0, r0, rz0, rth0, kappa        (one entry per knot)

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

from .jld2_gait import load_joint_traj, load_split_traj_alt
from .linearized import linearized_step
from .residual import get_residual
from .trajectory import trajectory_from_gait

REF = "/root/reference/src/dynamics"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# robot -> (gait file, κ used by the MPC example that defines the config)
CONFIGS = {
    # examples/quadruped/monte_carlo.jl:20-31 (κ_mpc = 1e-4)
    "quadruped": ("quadruped/gaits/gait2.jld2", 1.0e-4),
    # examples/flamingo/piecewise.jl:14-28 (κ_mpc = 2e-4)
    "flamingo": ("flamingo/gaits/gait_forward_36_4.jld2", 2.0e-4),
    # examples/centroidal_quadruped/flat_trot.jl:21-35 (κ_mpc = 2e-4)
    "centroidal_quadruped": ("centroidal_quadruped/gaits/inplace_trot_v4.jld2", 2.0e-4),
}


def hopper():
    """examples/hopper/flat.jl:15-27: `gait_forward.jld2` is a serialized ContactTraj (:joint_traj), κ_mpc = 2e-4."""
    tr = load_joint_traj(os.path.join(REF, "hopper_2D/gaits/gait_forward.jld2"))
    res = get_residual("hopper_2D")
    H, kappa = tr["H"], 2.0e-4
    nq, nc, nb = res.model.nq, res.model.nc, res.model.nb
    z, th = tr["z"], tr["theta"]
    gait = dict(q=tr["q"], u=tr["u"], gamma=tr["gamma"], b=tr["b"], psi=z[:, nq + nc + nb:nq + 2 * nc + nb],
                eta=z[:, nq + 3 * nc + nb:nq + 3 * nc + 2 * nb], mu=float(th[0, -2]), h=tr["h"], w=tr["w"])
    np.savez_compressed(os.path.join(OUT, "hopper_2D_gait.npz"), **gait)
    r0 = np.zeros((H, res.idx.nz)); rz0 = np.zeros((H, res.idx.nz, res.idx.nz)); rth0 = np.zeros((H, res.idx.nz, res.idx.ntheta))
    for t in range(H):
        r0[t], rz0[t], rth0[t] = linearized_step(res, z[t], th[t], kappa)
    np.savez_compressed(os.path.join(OUT, "hopper_2D_lin.npz"), z0=z, th0=th, r0=r0, rz0=rz0, rth0=rth0, kappa=kappa)
    worst = max(np.linalg.norm(res.r(z[t], th[t], 0.0)) for t in range(H))
    print("hopper_2D H =", H, "max ||r(z_t, θ_t, 0)|| =", worst)


# extra reference gaits kept only as MODEL PINS (tests/test_model_pins.py): a stand pins mass, gravity, B_func and the
# contact stack unambiguously; inplace_trot_v0 was generated with the undamped model and pins the translational rows.
# (Of the quadruped gaits only gait2 — the one the reference's tests use — satisfies the identity with the shipped model:
# gait1 0.65, gait3 0.42, calipso_gait8 2.9, calipso_gait11 4.4 — generated with other model versions.)
PIN_GAITS = {
    "centroidal_quadruped_stand_euler_v0": "centroidal_quadruped/gaits/stand_euler_v0.jld2",
    "centroidal_quadruped_inplace_trot_v0": "centroidal_quadruped/gaits/inplace_trot_v0.jld2",
}


def pins():
    for name, gait_file in PIN_GAITS.items():
        gait = load_split_traj_alt(os.path.join(REF, gait_file))
        np.savez_compressed(os.path.join(OUT, f"{name}_gait.npz"), **gait)
        print("pin", name, "H =", gait["u"].shape[0], "h =", gait["h"], "mu =", gait["mu"])


def main(robots=None):
    os.makedirs(OUT, exist_ok=True)
    if not robots or "pins" in robots:
        pins()
        if robots == ["pins"]:
            return
    if not robots or "hopper_2D" in robots:
        hopper()
    for name, (gait_file, kappa) in CONFIGS.items():
        if robots and name not in robots:
            continue
        gait = load_split_traj_alt(os.path.join(REF, gait_file))
        np.savez_compressed(os.path.join(OUT, f"{name}_gait.npz"), **gait)
        res = get_residual(name)
        ref = trajectory_from_gait(res.model, gait)
        H = ref.H
        r0 = np.zeros((H, res.idx.nz))
        rz0 = np.zeros((H, res.idx.nz, res.idx.nz))
        rth0 = np.zeros((H, res.idx.nz, res.idx.ntheta))
        for t in range(H):
            r0[t], rz0[t], rth0[t] = linearized_step(res, ref.z[t], ref.theta[t], kappa)
        np.savez_compressed(os.path.join(OUT, f"{name}_lin.npz"), z0=ref.z, th0=ref.theta, r0=r0, rz0=rz0,
                            rth0=rth0, kappa=kappa)
        print(name, "H =", H, "max|r0| =", float(np.abs(r0).max()))


if __name__ == "__main__":
    main(sys.argv[1:])
