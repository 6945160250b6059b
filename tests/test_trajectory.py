"""Product-side reference-trajectory handling (`contactimplicitmpc.jl_b200/trajectory.py`):
the JLD2 gait reader, `ContactTraj` and `tracking_error`, against the committed golden fixtures (made by the
ORACLE's independent reader / restatement) and the oracle's `tracking_error`."""
import os

import numpy as np
import pytest

from common import ROOT, SIZES, load_gait, load_lin

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


# ---- JLD2 writer ----------------------------------------------------------------------------------
# SHA-256 of the reference's own gait files (written by JLD2.jl under the Julia version named in their text header); the
# committed golden .npz hold exactly their contents (oracle/make_golden.py).
_GAIT_DIGESTS = {
    "quadruped": ("1.6.0", 53305, "98af32a60bbd486e15319fde92568e70b392d5a71b25b11ee526c9a6ca1a7d12"),        # quadruped/gaits/gait2.jld2
    "flamingo": ("1.6.0", 58853, "8e7b01ce970e8f8335a71073e6b5dd88f0629a369a20ddf32efef7fb0193e7bc"),         # flamingo/gaits/gait_forward_36_4.jld2
    "centroidal_quadruped": ("1.7.2", 70441, "8164961526e6990a662a13643b0d0fd46671a3b32f8b6886c4b045e84183d580"),  # …/inplace_trot_v4.jld2
}


@pytest.mark.parametrize("robot", sorted(_GAIT_DIGESTS))
def test_jld2_writer_reproduces_the_reference_gait_files_byte_for_byte(robot):
    """`save_split_traj_alt` (the reference's `@save path qm um γm bm ψm ηm μm hm`, read back by
    `get_trajectory(...; load_type = :split_traj_alt)`, src/controller/trajectory.jl:168-179): re-encoding the contents of a
    reference gait gives the very file JLD2.jl wrote — same length, same SHA-256 (object-header and superblock checksums
    included).  No Julia here, so JLD2.jl's own output is the oracle for the writer."""
    import hashlib
    from cimpc_b200 import package
    jw = __import__("importlib").import_module(package.__name__ + ".jld2_writer")
    g = np.load(os.path.join(ROOT, "tests", "golden", f"{robot}_gait.npz"))
    julia, size, digest = _GAIT_DIGESTS[robot]
    blob = jw.encode_split_traj_alt(g["q"], g["u"], g["gamma"], g["b"], g["psi"], g["eta"], float(g["mu"]), float(g["h"]), julia=julia)
    assert len(blob) == size and hashlib.sha256(blob).hexdigest() == digest


def test_jld2_writer_round_trip_and_reference_tree(tmp_path):
    """A gait with other sizes goes through writer → reader unchanged; and, where the reference tree is present (this
    container, not the GPU box), EVERY `:split_traj_alt` gait file of the reference is reproduced byte for byte — except
    files in which several knots alias one Julia vector (JLD2 stores such a vector once; values alone cannot tell)."""
    from cimpc_b200 import package
    jw = __import__("importlib").import_module(package.__name__ + ".jld2_writer")
    tr = package.trajectory
    rng = np.random.default_rng(3)
    H, nq, nu, nc, nb = 37, 7, 3, 2, 4
    g = dict(q=rng.standard_normal((H + 2, nq)), u=rng.standard_normal((H, nu)), gamma=rng.random((H, nc)),
             b=rng.random((H, nb)), psi=rng.random((H, nc)), eta=rng.random((H, nb)), mu=0.7, h=0.0125)
    for julia in ("1.6.0", "1.8.2"):
        path = str(tmp_path / f"gait_{julia}.jld2")
        jw.save_split_traj_alt(path, **g, julia=julia)
        back = tr.load_gait(path, "split_traj_alt")
        for k, v in g.items():
            assert np.array_equal(np.asarray(back[k]), np.asarray(v)), k
    # a long horizon: more than 255 bytes of references switches the header to a two-byte size field
    g2 = dict(g, q=rng.standard_normal((302, nq)), **{k: rng.random((300, g[k].shape[1])) for k in ("u", "gamma", "b", "psi", "eta")})
    path = str(tmp_path / "long.jld2")
    jw.save_split_traj_alt(path, **g2)
    assert np.array_equal(tr.load_gait(path)["eta"], g2["eta"])
    ref_root = "/root/reference/src/dynamics"
    if not os.path.isdir(ref_root):
        return
    import glob
    same = aliased = 0
    for p in sorted(glob.glob(os.path.join(ref_root, "*", "gaits", "*.jld2"))):
        orig = open(p, "rb").read()
        try:
            gg = tr.load_gait(p, "split_traj_alt")
        except KeyError:
            continue  # a :joint_traj file (serialized ContactTraj)
        blob = jw.encode_split_traj_alt(gg["q"], gg["u"], gg["gamma"], gg["b"], gg["psi"], gg["eta"], gg["mu"], gg["h"],
                                        julia=orig[52:57].decode())
        if blob == orig:
            same += 1
        else:
            assert len(blob) > len(orig), p  # aliased vectors were written once by JLD2
            aliased += 1
    assert same >= 30 and aliased <= 2, (same, aliased)
