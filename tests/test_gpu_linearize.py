"""Device-side linearization (`cimpc_linearize`: generated r!, rz!, rθ! evaluated on the GPU at every reference
knot) against the committed golden `LinearizedStep` data, which the oracle's independent
sympy restatement produced from the reference's gait files (oracle/make_golden.py).  Run with `-m gpu`.

Reference analogue: `ImplicitTrajectory(ref_traj, s; κ)` builds `lin[t] = LinearizedStep(s, z_t, θ_t, κ)`
(src/controller/implicit_dynamics.jl:56) and test/controller/linearized_solver.jl:38-67 checks the linearized
hooks against `rz0 \\ r0`, `rz0 \\ rθ0`."""
import numpy as np
import pytest

from common import SIZES, load_gait, load_lin, make_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("robot", ["quadruped", "flamingo", "hopper_2D", "centroidal_quadruped"])
def test_device_linearization_matches_golden(cuda_device, robot):
    import cimpc_b200 as cb
    lin = load_lin(robot)
    kappa = float(lin["kappa"])
    im = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], kappa=kappa, mode="configuration")
    r0, rz0, rth0 = im.get_linearization()
    # fp64 straight-line code on both sides (different CSE): agreement to round-off of the largest entries
    for name, a, b in (("r0", r0, lin["r0"]), ("rz0", rz0, lin["rz0"]), ("rth0", rth0, lin["rth0"])):
        err = np.abs(a - b).max() / max(1.0, np.abs(b).max())
        assert err < 1e-10, (robot, name, err)
    # structural zeros stay exactly zero
    assert np.all(rz0[lin["rz0"] == 0.0] == 0.0)


@pytest.mark.parametrize("robot,mode", [("quadruped", "configuration"), ("flamingo", "configurationforce")])
def test_solves_after_device_linearization_match_uploaded(cuda_device, robot, mode):
    """The IP solves on a device-linearized trajectory equal those on the uploaded golden linearization."""
    import cimpc_b200 as cb
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-8, max_iter=100, max_ls=0, diff_sol=True)
    im_up = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], mode=mode, opts=opts)
    im_dev = cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], kappa=float(lin["kappa"]), mode=mode, opts=opts)
    knot, theta, q2 = make_batch(robot, lin, gait, 500, seed=3)
    za, dza, sa, ia = im_up.solve_host(knot, theta, q2)
    zb, dzb, sb, ib = im_dev.solve_host(knot, theta, q2)
    assert np.array_equal(sa, sb)
    # the two linearizations agree to 1e-13; at κ_tol = 1e-8 a convergence test can sit on the tolerance, so a handful of
    # problems may need one iteration more or less (flamingo, cond(rz) ≈ 5e5: 2 of 500; quadruped: none)
    same = ia == ib
    assert same.mean() >= 0.99, (robot, np.flatnonzero(~same), ia[~same], ib[~same])
    # the solutions (solved to 1e-8) amplify the 1e-13 by the conditioning of the contact problem (observed 1e-7
    # absolute on z of magnitude 28)
    assert np.abs(za - zb)[same].max() < 1e-7 * max(1.0, np.abs(za).max())
    assert np.abs(dza - dzb)[same].max() < 1e-6 * max(1.0, np.abs(dza).max())
    # re-linearizing in place (`update!`) at shifted knots changes the constants, and back restores them
    z1 = lin["z0"].copy(); z1[:, :im_dev.nq] += 0.01
    im_dev.linearize(z1, lin["th0"], float(lin["kappa"]))
    zc, _, _, _ = im_dev.solve_host(knot, theta, q2)
    assert np.abs(zc - zb).max() > 1e-6
    im_dev.linearize(lin["z0"], lin["th0"], float(lin["kappa"]))
    zd, _, sd, idd = im_dev.solve_host(knot, theta, q2)
    assert np.array_equal(zd, zb) and np.array_equal(idd, ib)


@pytest.mark.parametrize("robot", ["hopper_2D", "flamingo", "quadruped"])
def test_device_linearization_on_piecewise_terrain_matches_oracle(cuda_device, robot):
    """`cimpc_linearize` on a `<robot>_piecewise` context (LinearizedStep of an `approx = true` simulation,
    src/simulation/residual_approx.jl:14-99) against the oracle's terrain residual: r, rz, rθ at reference knots moved
    over the terrain."""
    import cimpc_b200 as cb
    from oracle.residual import get_residual
    lin = load_lin(robot)
    res = get_residual(robot + "_piecewise")
    nq = SIZES[robot][0]
    H = lin["z0"].shape[0]
    z0, th0 = lin["z0"].copy(), lin["th0"].copy()
    x_new = np.linspace(0.1, 2.6, H)
    lift = np.array([res.terrain.height(x) for x in x_new])
    dx = x_new - z0[:, 0]
    for a in (z0[:, 0:nq], th0[:, 0:nq], th0[:, nq:2 * nq]):  # q2, q0, q1 move together
        a[:, 0] += dx
        a[:, 1] += lift
    kappa = float(lin["kappa"])
    im = cb.ImplicitTrajectory(*SIZES[robot], z0, th0, kappa=kappa, mode="configuration", model=robot + "_piecewise")
    r0, rz0, rth0 = im.get_linearization()
    worst = 0.0
    for t in range(0, H, max(1, H // 12)):
        for name, a, b in (("r0", r0[t], res.r(z0[t], th0[t], kappa)), ("rz0", rz0[t], res.rz(z0[t], th0[t])),
                           ("rth0", rth0[t], res.rth(z0[t], th0[t]))):
            err = np.abs(a - b).max() / max(1.0, np.abs(b).max())
            worst = max(worst, err)
            assert err < 1e-9, (robot, t, name, err)
    # and it differs from the flat linearization where the ground is not flat
    im_flat = cb.ImplicitTrajectory(*SIZES[robot], z0, th0, kappa=kappa, mode="configuration")
    rf, rzf, _ = im_flat.get_linearization()
    assert np.abs(rf - r0).max() > 1e-3 and np.abs(rzf - rz0).max() > 1e-2
