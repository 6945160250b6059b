"""CUDA path vs the CPU oracle, through the C ABI (run with `-m gpu` on a B200).

Parity protocol.  The reference's iteration has two sources of irreproducible
round-off that no re-implementation (including two faithful CPU restatements of the reference, see
tests/test_oracle.py::test_c_restatement_matches_numpy_oracle) can match bit for bit:
  (1) its MGS-QR solve loses ≈ cond(S)²·eps  (1e-9 quadruped … 1e-5 flamingo);
  (2) the residual back-tracking test `r_cand <= r_vio` compares two numbers that are both at
      round-off level after the first full Newton step, i.e. it is a coin flip on ≈1 % of problems.
So:
  * STRICT parity is asserted against the oracle with accurate LU solves and back-tracking off
    (max_ls = 0): EVERY problem, z to 1e-9, δz to 1e-8 relative, identical status and iteration count.
  * At the REFERENCE settings (MGS semantics, max_ls = 3) the CUDA result must be as close to the
    faithful MGS oracle as two faithful MGS oracles are to each other (quantile test), agree with the
    LU oracle to 1e-9 on ≥ 97 % of problems, match status everywhere, and every converged point must
    satisfy the convergence criteria when the residual is recomputed independently.
Tolerances are fp64; SURVEY.md §8c asked for 1e-7 / 1e-6 — the strict numbers here are tighter.
"""
import numpy as np
import pytest

from common import (SIZES, independent_violation, load_gait, load_lin, make_batch, oracle_solve_batch)

pytestmark = pytest.mark.gpu

CONFIGS = [
    # examples/quadruped/monte_carlo.jl:51-58 (the BASELINE metric's options)
    ("quadruped", "configuration", dict(r_tol=1e-4, kappa_tol=1e-4, max_iter=100)),
    # test/controller/implicit_dynamics.jl:11-15
    ("quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("quadruped", "configurationforce", dict(r_tol=1e-8, kappa_tol=1e-4)),
    # examples/flamingo/piecewise.jl:24-48 (default mode :configurationforce, κ = 2e-4)
    ("flamingo", "configurationforce", dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("flamingo", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),
    # examples/centroidal_quadruped/flat_trot.jl:31-62
    ("centroidal_quadruped", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("centroidal_quadruped", "configurationforce", dict(r_tol=1e-8, kappa_tol=2e-4)),
    # examples/hopper/flat.jl:24-48 (4 lanes per subproblem, 8 subproblems per warp)
    ("hopper_2D", "configuration", dict(r_tol=1e-8, kappa_tol=2e-4)),
    ("hopper_2D", "configurationforce", dict(r_tol=1e-8, kappa_tol=2e-4)),
]


def _ctx(robot, lin, mode, opts):
    import cimpc_b200 as cb
    return cb.ImplicitTrajectory(*SIZES[robot], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"],
                                 mode=mode, opts=opts)


def _oracle_opts(o):
    from oracle.ip import IPOptions
    return IPOptions(r_tol=o.r_tol, kappa_tol=o.kappa_tol, max_iter=o.max_iter, max_ls=o.max_ls,
                     ls_scale=o.ls_scale, diff_sol=o.diff_sol, eps_min=o.eps_min, kappa_reg=o.kappa_reg,
                     gamma_reg=o.gamma_reg, undercut=o.undercut)


def _c_oracle(robot, lin, mode, solver):
    from oracle.c_oracle import COracle
    return COracle(*SIZES[robot], lin, mode=mode, solver=solver)


def _errs(z, zo, dz, dzo):
    ez = np.abs(z - zo).max(axis=1) / np.maximum(1.0, np.abs(zo).max(axis=1))
    edz = np.abs(dz - dzo).max(axis=(1, 2)) / np.maximum(1.0, np.abs(dzo).max(axis=(1, 2)))
    return ez, edz


@pytest.mark.parametrize("robot,mode,kw", CONFIGS)
def test_strict_parity_every_problem(cuda_device, robot, mode, kw):
    import cimpc_b200 as cb
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(diff_sol=True, max_ls=0, **kw)
    im = _ctx(robot, lin, mode, opts)
    n = 20 * lin["z0"].shape[0] + 7  # every knot twenty times + a ragged tail
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    zo, dzo, sto, ito = _c_oracle(robot, lin, mode, "lu").solve(knot, theta, q2, _oracle_opts(opts))
    assert sto.mean() >= 0.995, "oracle did not converge on the seeded batch"
    assert np.array_equal(st, sto)
    assert np.array_equal(it[sto], ito[sto]), f"iteration counts differ on {np.mean(it != ito):.3%}"
    ez, edz = _errs(z, zo, dz, dzo)
    assert ez[sto].max() <= 1e-9, f"z error {ez[sto].max():.3e}"
    assert edz[sto].max() <= 1e-8, f"dz error {edz[sto].max():.3e}"
    # the python (numpy) oracle says the same on a sub-sample
    sub = np.arange(0, n, 17)
    zn, dzn, stn, itn = oracle_solve_batch(robot, lin, knot[sub], theta[sub], q2[sub], _oracle_opts(opts),
                                           mode=mode, solver="lu")
    ezn, edzn = _errs(z[sub], zn, dz[sub], dzn)
    ok = stn
    assert np.array_equal(it[sub][ok], itn[ok]) and ezn[ok].max() <= 1e-9 and edzn[ok].max() <= 1e-8


@pytest.mark.parametrize("robot,mode,kw", CONFIGS[:4] + CONFIGS[5:6] + CONFIGS[7:8])
def test_reference_settings_parity(cuda_device, robot, mode, kw):
    import cimpc_b200 as cb
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(diff_sol=True, **kw)  # max_ls = 3 (RoboDojo default)
    im = _ctx(robot, lin, mode, opts)
    n = 5 * lin["z0"].shape[0] + 3
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=101)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    oo = _oracle_opts(opts)
    z_lu, dz_lu, st_lu, it_lu = _c_oracle(robot, lin, mode, "lu").solve(knot, theta, q2, oo)
    z_cm, dz_cm, st_cm, it_cm = _c_oracle(robot, lin, mode, "mgs").solve(knot, theta, q2, oo)
    z_nm, dz_nm, st_nm, it_nm = oracle_solve_batch(robot, lin, knot, theta, q2, oo, mode=mode, solver="mgs")
    assert np.array_equal(st, st_lu) and np.array_equal(st, st_cm) and st.mean() >= 0.995
    # (a) accurate oracle: everything except the ≈1 % back-tracking coin flips agrees to 1e-9
    ez, edz = _errs(z, z_lu, dz, dz_lu)
    good = (ez <= 1e-9) & (edz <= 1e-8) & (it == it_lu)
    assert good.mean() >= 0.97, f"only {good.mean():.3%} of problems agree with the LU oracle"
    # (b) reference-faithful MGS oracle: CUDA is within the reference's own round-off floor
    floor, _ = _errs(z_nm, z_cm, dz_nm, dz_cm)   # two faithful restatements against each other
    ours, _ = _errs(z, z_nm, dz, dz_nm)
    for q in (0.5, 0.9):
        assert np.quantile(ours, q) <= 10.0 * np.quantile(floor, q) + 1e-11, (q, np.quantile(ours, q),
                                                                               np.quantile(floor, q))
    # (c) every returned point satisfies the stopping criteria, recomputed from the raw fixture
    rv, kv = independent_violation(robot, lin, knot, theta, z)
    assert (rv[st] < opts.r_tol + 1e-11).all() and (kv[st] < opts.kappa_tol * (1 + 1e-9)).all()


def test_mismatch_rate_at_bench_settings_full_batch(cuda_device):
    """SURVEY §8c asks for the mismatch RATE, measured and recorded, at the bench's own settings: quadruped, r_tol =
    κ_tol = 1e-4, max_ls = 3 (the reference's back-tracking, whose acceptance test compares round-off-level numbers),
    the full 655 360-subproblem sweep of BASELINE config 4 against the C oracle with LU-accurate solves.  Recorded in
    gpurun_out/r02_mismatch_rate.json (copied to profiles/).  Also exercises the output mask: only z and the status
    cross the bus."""
    import json
    import os
    import cimpc_b200 as cb
    robot, mode = "quadruped", "configuration"
    lin, gait = load_lin(robot), load_gait(robot)
    R, H = 65536, 10
    n = R * H
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=100)
    nq = SIZES[robot][0]
    stage = (np.arange(n) // R).astype(np.int32)
    theta = np.ascontiguousarray(theta - lin["th0"][knot] + lin["th0"][stage])
    q2 = np.ascontiguousarray(q2 - lin["z0"][knot, :nq] + lin["z0"][stage, :nq])
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=False)  # max_ls = 3
    im = _ctx(robot, lin, mode, opts)
    z = np.zeros((n, im.nz)); st = np.zeros(n, np.uint8); it = np.zeros(n, np.int32)
    im.solve_host_into(stage, theta, q2, None, z, None, st, it, out_mask=im.OUT_Z | im.OUT_STATUS)
    zo, _, sto, ito = _c_oracle(robot, lin, mode, "lu").solve(stage, theta, q2, _oracle_opts(opts))
    st = st.astype(bool)
    ez = np.abs(z - zo).max(axis=1) / np.maximum(1.0, np.abs(zo).max(axis=1))
    same_it = it == ito
    agree = same_it & (ez <= 1e-9) & (st == sto)
    rec = {"workload": "BASELINE config 4 sweep: quadruped, 655360 subproblems, r_tol = kappa_tol = 1e-4, max_ls = 3",
           "status_mismatches": int((st != sto).sum()), "converged_frac_gpu": float(st.mean()),
           "iteration_count_mismatch_rate": float(1.0 - same_it.mean()),
           "z_mismatch_rate_1e-9": float(1.0 - agree.mean()),
           "max_rel_z_error_where_iterations_agree": float(ez[same_it].max()),
           "median_rel_z_error": float(np.median(ez)),
           "mean_iterations_gpu": float(it.mean()), "mean_iterations_oracle": float(ito.mean())}
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_mismatch_rate.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(rec)
    assert rec["status_mismatches"] == 0 and st.all()
    assert agree.mean() >= 0.97, rec
    # every returned point satisfies the stopping criteria on the independently recomputed residual (a 1/64 sample)
    sel = np.arange(0, n, 64)
    rv, kv = independent_violation(robot, lin, stage[sel], theta[sel], z[sel])
    assert (rv < 1e-4 + 1e-11).all() and (kv < 1e-4 * (1 + 1e-9)).all()


def test_host_and_device_entry_points_agree(cuda_device):
    import torch
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, diff_sol=True)
    im = _ctx(robot, lin, "configuration", opts)
    n = 1000
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=3)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    zd, dzd, std, itd = im.solve_device(torch.from_numpy(knot).to(cuda_device), torch.from_numpy(theta).to(cuda_device),
                                        torch.from_numpy(q2).to(cuda_device))
    torch.cuda.synchronize()
    assert np.array_equal(z, zd.cpu().numpy())
    assert np.array_equal(dz, dzd.cpu().numpy().transpose(0, 2, 1))
    assert np.array_equal(st, std.cpu().numpy().astype(bool)) and np.array_equal(it, itd.cpu().numpy())
    # deterministic: a second launch returns the same bits
    z2, dz2, st2, it2 = im.solve_host(knot, theta, q2)
    assert np.array_equal(z, z2) and np.array_equal(dz, dz2) and np.array_equal(it, it2)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 31, 33, 127, 129])
def test_ragged_and_tiny_batches(cuda_device, n):
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True, max_ls=0)
    im = _ctx(robot, lin, "configuration", opts)
    knot, theta, q2 = make_batch(robot, lin, gait, max(n, 1), seed=11)
    knot, theta, q2 = knot[:n], theta[:n], q2[:n]
    z, dz, st, it = im.solve_host(knot, theta, q2)
    assert z.shape == (n, 43) and dz.shape == (n, 11, 30)
    if n:
        zo, dzo, sto, ito = _c_oracle(robot, lin, "configuration", "lu").solve(knot, theta, q2, _oracle_opts(opts))
        ez, edz = _errs(z, zo, dz, dzo)
        assert np.array_equal(st, sto) and np.array_equal(it, ito) and ez.max() <= 1e-9 and edz.max() <= 1e-8


def test_hopper_smallest_group(cuda_device):
    """hopper_2D sizes (nx = ny = 4 → 4 lanes per problem, 8 problems per warp) on a synthetic
    well-posed linearization: a random strictly monotone LCP-like problem with the block structure of the contact
    residual (simulation.jl:133-158: the fri row and the ψ column touch only the E / μ couplings), which the kernel's
    closed-form elimination of ψ1 relies on — an upload WITHOUT that structure is refused."""
    import cimpc_b200 as cb
    from oracle.c_oracle import COracle
    rng = np.random.default_rng(5)
    nq, nu, nw, nc, nb = SIZES["hopper_2D"]
    nz, nth, ny = nq + 4 * nc + 2 * nb, 2 * nq + nu + nw + 2, 2 * nc + nb
    nr = nc + nb
    H = 6
    z0 = np.abs(rng.standard_normal((H, nz))) + 0.5
    th0 = rng.standard_normal((H, nth))
    r0 = 0.1 * rng.standard_normal((H, nz))
    rz0 = np.zeros((H, nz, nz))
    rth0 = 0.3 * rng.standard_normal((H, nz, nth))
    for t in range(H):
        A = rng.standard_normal((nq, nq))
        rz0[t, :nq, :nq] = A @ A.T + nq * np.eye(nq)                 # Dx
        Jt = rng.standard_normal((nq, nr))
        rz0[t, :nq, nq:nq + nr] = -Jt                                  # Dy1[:, γb]; Dy1[:, ψ] = 0
        rz0[t, nq:nq + nr, :nq] = Jt.T                                 # Rx[imp ∪ mdp, :]; Rx[fri, :] = 0
        K = rng.standard_normal((nr, nr))
        rz0[t, nq:nq + nr, nq:nq + nr] = 0.1 * (K - K.T)               # Ry1[imp ∪ mdp, γb] (skew)
        rz0[t, nq + nc:nq + nr, nq + nr] = -1.0                        # −Eᵀ: mdp rows ← ψ
        rz0[t, nq + nr, nq:nq + nc] = -0.7                             # fri row: −μ γ
        rz0[t, nq + nr, nq + nc:nq + nr] = 1.0                         #          + E b
        rz0[t, nq:nq + ny, nq + ny:] = np.eye(ny)                      # Ry2 = 1
        rz0[t, nq + ny:, nq:nq + ny] = np.diag(z0[t, nq + ny:])
        rz0[t, nq + ny:, nq + ny:] = np.diag(z0[t, nq:nq + ny])
        rth0[t, nq + ny:, :] = 0.0
        rth0[t, nq + nr, :2 * nq + nu] = 0.0                           # the fri row depends on θ through μ only
    bad = rz0.copy()
    bad[:, nq + nr, nq + nr] = 0.3                                     # Ry1[fri, ψ] ≠ 0: not a contact residual
    with pytest.raises(cb.CimpcError) as ei:
        cb.ImplicitTrajectory(nq, nu, nw, nc, nb, z0, th0, r0, bad, rth0)
    assert ei.value.code == 2
    lin = dict(z0=z0, th0=th0, r0=r0, rz0=rz0, rth0=rth0)
    for mode in ("configuration", "configurationforce"):
        opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=1e-6, diff_sol=True, max_ls=0)
        im = cb.ImplicitTrajectory(nq, nu, nw, nc, nb, z0, th0, r0, rz0, rth0, mode=mode, opts=opts)
        n = 203
        knot = (np.arange(n) % H).astype(np.int32)
        theta = th0[knot] + 0.05 * rng.standard_normal((n, nth))
        q2 = z0[knot, :nq] + 0.05 * rng.standard_normal((n, nq))
        z, dz, st, it = im.solve_host(knot, theta, q2)
        zo, dzo, sto, ito = COracle(nq, nu, nw, nc, nb, lin, mode=mode, solver="lu").solve(knot, theta, q2,
                                                                                          _oracle_opts(opts))
        assert sto.mean() > 0.8 and np.array_equal(st, sto)  # random LCP-like data: not every draw is solvable
        ez, edz = _errs(z, zo, dz, dzo)
        assert (it == ito)[sto].mean() > 0.99
        sel = sto & (it == ito)
        # random draws include ill-conditioned Schur complements: δz of the worst draw sits at cond·eps
        # (1.2e-8 observed), so the bulk is held to 1e-9 and the maximum to SURVEY §8c's 1e-6
        assert ez[sel].max() <= 1e-9 and np.quantile(edz[sel], 0.99) <= 1e-9 and edz[sel].max() <= 1e-6


def test_altitude_offsets(cuda_device):
    """`set_altitude!` (implicit_dynamics.jl:141-154): per-problem alt enters the impact rows (rlin! :370)."""
    import cimpc_b200 as cb
    robot = "flamingo"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True, max_ls=0)
    im = _ctx(robot, lin, "configurationforce", opts)
    n = 300
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=12)
    rng = np.random.default_rng(13)
    alt = 0.02 * rng.random((n, 4))  # piecewise terrain heights under the four contacts
    z, dz, st, it = im.solve_host(knot, theta, q2, alt=alt)
    zo, dzo, sto, ito = _c_oracle(robot, lin, "configurationforce", "lu").solve(knot, theta, q2, _oracle_opts(opts),
                                                                              alt=alt)
    ez, edz = _errs(z, zo, dz, dzo)
    assert np.array_equal(st, sto) and np.array_equal(it, ito) and ez.max() <= 1e-9 and edz.max() <= 1e-8
    z_flat, _, _, _ = im.solve_host(knot, theta, q2)
    assert np.abs(z - z_flat).max() > 1e-4  # the offsets matter
    rv, kv = independent_violation(robot, lin, knot, theta, z, alt=alt)
    assert (rv[st] < 1e-8 + 1e-11).all()


def test_non_convergence_is_reported_not_hidden(cuda_device):
    """status = 0 + last iterate when the iteration cap is hit (the reference @warns and continues)."""
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True, max_ls=0, max_iter=2)
    im = _ctx(robot, lin, "configuration", opts)
    knot, theta, q2 = make_batch(robot, lin, gait, 64, seed=14)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    zo, dzo, sto, ito = _c_oracle(robot, lin, "configuration", "lu").solve(knot, theta, q2, _oracle_opts(opts))
    assert not st.any() and np.array_equal(st, sto) and (it == 2).all() and np.array_equal(it, ito)
    ez, edz = _errs(z, zo, dz, dzo)
    assert ez.max() <= 1e-9 and edz.max() <= 1e-8
    # diff_sol = 0: no sensitivity output requested / produced
    o2 = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=False)
    z2, dz2, st2, it2 = im.solve_host(knot, theta, q2, opts=o2)
    assert dz2 is None and st2.all()


def test_error_codes(cuda_device):
    import ctypes as C
    import cimpc_b200 as cb
    from contactimplicitmpc_jl_b200 import capi
    lib = cb.load_library()
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    im = _ctx(robot, lin, "configuration", cb.InteriorPointOptions(diff_sol=True))
    knot, theta, q2 = make_batch(robot, lin, gait, 8, seed=15)
    bad = knot.copy()
    bad[3] = lin["z0"].shape[0]  # one past the last knot
    with pytest.raises(cb.CimpcError) as e:
        im.solve_host(bad, theta, q2)
    assert e.value.code == 1
    # solve before upload → NOT_INITIALIZED
    ctx = C.c_void_p()
    desc = capi.ModelDesc(11, 8, 2, 4, 8, 0)
    assert lib.cimpc_create(C.byref(ctx), 0, C.byref(desc)) == 0
    o = cb.InteriorPointOptions().to_c()
    z = np.zeros((8, 43)); st = np.zeros(8, np.uint8); it = np.zeros(8, np.int32)
    rc = lib.cimpc_ip_solve_batch_host(ctx, 8, knot.ctypes.data, theta.ctypes.data, q2.ctypes.data, None,
                                       C.byref(o), z.ctypes.data, None, st.ctypes.data, it.ctypes.data)
    assert rc == 3
    lib.cimpc_destroy(ctx)


def test_relinearization_update(cuda_device):
    """`update!` (linearized_solver.jl:497-565): uploading a new linearization replaces the old one."""
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-8, kappa_tol=2e-4, diff_sol=True, max_ls=0)
    im = _ctx(robot, lin, "configuration", opts)
    H = lin["z0"].shape[0]
    roll = {k: (np.roll(v, 7, axis=0) if k != "kappa" else v) for k, v in lin.items()}
    knot, theta, q2 = make_batch(robot, roll, gait, 2 * H, seed=16)
    im.update(roll["z0"], roll["th0"], roll["r0"], roll["rz0"], roll["rth0"])
    z, dz, st, it = im.solve_host(knot, theta, q2)
    zo, dzo, sto, ito = _c_oracle(robot, roll, "configuration", "lu").solve(knot, theta, q2, _oracle_opts(opts))
    ez, edz = _errs(z, zo, dz, dzo)
    assert np.array_equal(st, sto) and np.array_equal(it, ito) and ez.max() <= 1e-9 and edz.max() <= 1e-8


def test_full_size_batch_properties(cuda_device):
    """BASELINE config 2 size: quadruped, H = 10, 4096 rollouts → 40 960 subproblems in one launch.
    Full-size parity against the C oracle (LU variant) plus the size-independent properties:
    every problem converges, and the stopping criteria hold on the independently recomputed residual."""
    import cimpc_b200 as cb
    robot = "quadruped"
    lin, gait = load_lin(robot), load_gait(robot)
    opts = cb.InteriorPointOptions(r_tol=1e-4, kappa_tol=1e-4, max_iter=100, diff_sol=True, max_ls=0)
    im = _ctx(robot, lin, "configuration", opts)
    n = 40960
    knot, theta, q2 = make_batch(robot, lin, gait, n, seed=17)
    z, dz, st, it = im.solve_host(knot, theta, q2)
    assert st.all()
    zo, dzo, sto, ito = _c_oracle(robot, lin, "configuration", "lu").solve(knot, theta, q2, _oracle_opts(opts))
    assert np.array_equal(st, sto)
    same = it == ito
    assert same.mean() >= 0.999
    ez, edz = _errs(z, zo, dz, dzo)
    assert ez[same].max() <= 1e-9 and edz[same].max() <= 1e-8
    for lo in range(0, n, 8192):
        sl = slice(lo, lo + 8192)
        rv, kv = independent_violation(robot, lin, knot[sl], theta[sl], z[sl])
        assert (rv < 1e-4).all() and (kv < 1e-4).all()
