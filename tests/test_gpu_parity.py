"""CUDA path vs the CPU oracle, through the C ABI (run with `-m gpu` on a B200).

Tolerances: ‖z − z*‖∞ ≤ 1e-7·max(1, ‖z*‖∞), δz to 1e-6 relative, identical status,
identical iteration counts (mismatches are reported and must stay below 1 %)."""
import numpy as np
import pytest

from common import (SIZES, independent_violation, load_gait, load_lin, make_batch, oracle_solve_batch)

pytestmark = pytest.mark.gpu

Z_TOL = 1e-7
DZ_TOL = 1e-6


def _ctx(robot, lin, mode, opts):
    import cimpc_b200 as cb
    nq, nu, nw, nc, nb = SIZES[robot]
    return cb.ImplicitTrajectory(nq, nu, nw, nc, nb, lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                 mode=mode, opts=opts)


def _opts(cb, **kw):
    return cb.InteriorPointOptions(**kw)


def _oracle_opts(o):
    from oracle.ip import IPOptions
    return IPOptions(r_tol=o.r_tol, kappa_tol=o.kappa_tol, max_iter=o.max_iter, max_ls=o.max_ls,
                     ls_scale=o.ls_scale, diff_sol=o.diff_sol, eps_min=o.eps_min, kappa_reg=o.kappa_reg,
                     gamma_reg=o.gamma_reg, undercut=o.undercut)


def _compare(z, dz, st, it, zo, dzo, sto, ito):
    assert np.array_equal(st, sto), f"status mismatch at {np.nonzero(st != sto)[0][:10]}"
    same_it = it == ito
    assert same_it.mean() >= 0.99, f"iteration-count mismatch rate {1 - same_it.mean():.3%}"
    sel = same_it & sto
    scale = np.maximum(1.0, np.abs(zo).max(axis=1, keepdims=True))
    err_z = (np.abs(z - zo) / scale)[sel].max()
    assert err_z <= Z_TOL, f"z error {err_z:.3e}"
    if dzo is not None:
        dscale = np.maximum(1.0, np.abs(dzo).max(axis=(1, 2), keepdims=True))
        err_dz = (np.abs(dz - dzo) / dscale)[sel].max()
        assert err_dz <= DZ_TOL, f"dz error {err_dz:.3e}"
    return err_z


@pytest.mark.parametrize("robot,mode,opt_kw", [
    # examples/quadruped/monte_carlo.jl:51-58 (the BASELINE metric's options)
    ("quadruped", "configuration", dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True)),
    # test/controller/implicit_dynamics.jl:11-15
    ("quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True)),
    ("quadruped", "configurationforce", dict(r_tol=1e-8, kappa_tol=1e-4, diff_sol=True)),
    # examples/flamingo/piecewise.jl (default mode :configurationforce, κ = 2e-4)
    ("flamingo", "configurationforce", dict(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True)),
    ("flamingo", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True)),
    ("centroidal_quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True)),
])
def test_parity_vs_oracle(cuda_device, robot, mode, opt_kw):
    import cimpc_b200 as cb
    lin, gait = load_lin(robot), load_gait(robot)
    opts = _opts(cb, **opt_kw)
    im = _ctx(robot, lin, mode, opts)
    n = 3 * lin["z0"].shape[0] + 7  # every knot three times + a ragged tail
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    zo, dzo, sto, ito = oracle_solve_batch(robot, lin, knot, theta, q2, _oracle_opts(opts), mode=mode)
    assert sto.mean() > 0.9
    _compare(z, dz, st, it, zo, dzo, sto, ito)
    rv, kv = independent_violation(robot, lin, knot, theta, z)
    assert (rv[st] < opts.r_tol * (1 + 1e-6) + 1e-12).all() and (kv[st] < opts.kappa_tol * (1 + 1e-6)).all()
