"""Product-side reference-trajectory handling (`contactimplicitmpc.jl_b200/trajectory.py`):
the JLD2 gait reader, `ContactTraj` and `tracking_error`, against the committed golden fixtures (made by the
ORACLE's independent reader / restatement) and the oracle's `tracking_error`."""
import os

import numpy as np
import pytest

from common import SIZES, load_gait, load_lin

REF = "/root/reference/src/dynamics"
GAITS = {"quadruped": ("quadruped/gaits/gait2.jld2", "split_traj_alt"),
         "flamingo": ("flamingo/gaits/gait_forward_36_4.jld2", "split_traj_alt"),
         "centroidal_quadruped": ("centroidal_quadruped/gaits/inplace_trot_v4.jld2", "split_traj_alt"),
         "hopper_2D": ("hopper_2D/gaits/gait_forward.jld2", "joint_traj")}


@pytest.mark.parametrize("robot", list(GAITS))
def test_gait_reader_and_contact_traj(robot):
    """`get_trajectory` on the reference's own gait files (container only: they are not shipped) reproduces the golden
    (z, θ) the linearization fixtures were built at."""
    import cimpc_b200 as cb
    path, load_type = GAITS[robot]
    path = os.path.join(REF, path)
    if not os.path.exists(path):
        pytest.skip("reference gait files are only present in the build container")
    gait = cb.load_gait(path, load_type)
    gold, lin = load_gait(robot), load_lin(robot)
    for k in ("q", "u", "gamma", "b"):
        assert np.array_equal(gait[k], gold[k]), k
    assert float(gait["h"]) == float(gold["h"])
    tr = cb.ContactTraj.from_gait(robot, gait)
    assert (tr.nq, tr.nu, tr.nw, tr.nc, tr.nb) == SIZES[robot]
    assert np.abs(tr.z - lin["z0"]).max() < 1e-14
    assert np.abs(tr.theta - lin["th0"]).max() == 0.0


def test_contact_traj_from_golden_gait_and_npz_round_trip(tmp_path):
    """Same construction from the committed fixture (runs anywhere), plus save/load."""
    import cimpc_b200 as cb
    gold, lin = load_gait("quadruped"), load_lin("quadruped")
    tr = cb.ContactTraj.from_gait("quadruped", gold, kappa=1e-4)
    assert np.abs(tr.z - lin["z0"]).max() < 1e-14 and np.abs(tr.theta - lin["th0"]).max() == 0.0
    cb.save_traj(str(tmp_path / "t.npz"), tr)
    t2 = cb.load_traj(str(tmp_path / "t.npz"))
    for k in ("q", "u", "w", "gamma", "b", "z", "theta"):
        assert np.array_equal(getattr(tr, k), getattr(t2, k))
    assert t2.kappa == 1e-4 and t2.h == tr.h
    # update_z! / update_θ! are consistent with the stored arrays
    z, th = tr.z.copy(), tr.theta.copy()
    tr.update_z(); tr.update_theta()
    assert np.array_equal(z, tr.z) and np.array_equal(th, tr.theta)


def test_tracking_error_matches_oracle():
    """trajectory.jl:188-217 restated twice (product / oracle) agree on a perturbed, strided simulated trajectory."""
    import cimpc_b200 as cb
    from oracle.residual import get_residual
    from oracle.simulator import tracking_error as oracle_te
    from oracle.trajectory import trajectory_from_gait
    gold = load_gait("quadruped")
    res = get_residual("quadruped")
    ref_o = trajectory_from_gait(res.model, gold)
    ref_p = cb.ContactTraj.from_gait("quadruped", gold)
    rng = np.random.default_rng(3)
    N_sample, H_sim = 5, 430  # more than one gait period at the MPC rate (60 knots × 5 = 300 steps)
    nq, nu, nw, nc, nb = SIZES["quadruped"]
    sim_q = rng.standard_normal((H_sim + 2, nq)); sim_u = rng.standard_normal((H_sim, nu))
    sim_g = rng.standard_normal((H_sim, nc)); sim_b = rng.standard_normal((H_sim, nb))
    a = cb.tracking_error(ref_p, sim_q, sim_u, sim_g, sim_b, N_sample)
    b = oracle_te(ref_o, res.model, sim_q, sim_u, sim_g, sim_b, N_sample)
    assert np.allclose(a, b, rtol=1e-13, atol=0)


def test_reference_window_matches_rot_n_stride():
    """The product's policy glue (`rollout.py::ReferenceWindow`: `rot_n_stride!`, mpc_utils.jl:1-101, and
    `update_window!`, policy.jl:162-171, on the shared reference incl. the contact forces used in :configurationforce
    mode) against the oracle's restatement, over more than one gait period."""
    import cimpc_b200 as cb
    from oracle.models import get_model
    from oracle.newton import get_stride, rot_n_stride, update_window
    from oracle.trajectory import trajectory_from_gait
    robot, H_mpc = "flamingo", 15
    gold = load_gait(robot)
    m = get_model(robot)
    ref = trajectory_from_gait(m, gold)
    pw = cb.ReferenceWindow(ref.q, ref.u, H_mpc, ref.gamma, ref.b)
    ptraj, stride, window = ref.copy(), get_stride(m, ref), list(range(H_mpc + 2))
    assert np.array_equal(pw.stride, stride)
    for step in range(ref.H + 7):
        assert list(pw.window) == window
        assert np.array_equal(pw.q, ptraj.q) and np.array_equal(pw.u, ptraj.u)
        assert np.array_equal(pw.gamma, ptraj.gamma) and np.array_equal(pw.b, ptraj.b)
        pw.advance()
        rot_n_stride(ptraj, stride)
        window = update_window(window, ref.H)
    pw.reset()
    assert np.array_equal(pw.q, ref.q) and list(pw.window) == list(range(H_mpc + 2))
