"""Every config robot pinned to REFERENCE DATA, independently for the oracle's restatement and for the product's
generated code (so the two transcriptions being textual siblings does not matter).

The pin is the reference's own identity (test/simulator/quadruped.jl:16-19, test/simulator/flamingo.jl): a shipped gait
`(q, u, γ, b, ψ, η)` satisfies `residual(z_t, θ_t, 0) ≈ 0` knot by knot.  The dynamics rows involve every mass, inertia,
link length, centre of mass, the input map B, the contact Jacobian and the contact-force stack, so a wrong parameter or
sign in a transcription shows up at the 1e-2 .. 1 level.  Fixtures: tests/golden/*_gait.npz, produced from the reference's
`.jld2` files by oracle/make_golden.py.

Centroidal quadruped (the robot of BASELINE config 5).  `inplace_trot_v4.jld2` — the gait of
examples/centroidal_quadruped/flat_trot.jl — does NOT satisfy the identity as a whole, and the cause is the gait, not a
parameter: knots 32-63 (all feet in stance) satisfy the translational dynamics rows to 2e-13 (orientation rows 4e-5), knots 0-31 (swing of feet 2 and 3) are
off by 1e-2 in the body rows and by 0.21 / 0.135 at the touch-down knots 30 / 31, where the swing feet stop at z = 0
from −0.675 m/s with no impact impulse in γ (m_foot·Δv = 0.2·0.675 = 0.135; plus the unbalanced −u_z − h·m·g): the
first half is a kinematically drawn swing, not a solution of the variational integrator.  No (mass_foot, damping,
mass_body) fit removes it (least squares leaves 0.17).  What pins the model instead: the stance half of v4 (1e-12),
`stand_euler_v0.jld2` (dynamics rows < 1e-4 with the shipped, damped model) and `inplace_trot_v0.jld2` (generated with
`centroidal_quadruped_undamped`: translational rows 1e-13 with joint friction off)."""
import numpy as np
import pytest

from common import SIZES, GeneratedResidual, load_gait, load_lin


def _layout(robot):
    nq, nu, nw, nc, nb = SIZES[robot]
    o = np.cumsum([0, nq, nc, nb, nc, nc, nb, nc])
    return nq, nu, nw, nc, nb, o


def _gait_points(gait, robot, mu=None):
    """(z_t, θ_t) without s1 (filled by the caller from the model under test)."""
    nq, nu, nw, nc, nb, o = _layout(robot)
    mu = gait["mu"] if mu is None else mu
    H = gait["u"].shape[0]
    nf = nb // nc
    for t in range(H):
        q2 = gait["q"][t + 2]
        s2 = mu * gait["gamma"][t] - gait["b"][t].reshape(nc, nf).sum(1)
        z = np.concatenate([q2, gait["gamma"][t], gait["b"][t], gait["psi"][t], np.zeros(nc), gait["eta"][t], s2])
        w = gait["w"][t] if "w" in gait else np.zeros(nw)
        th = np.concatenate([gait["q"][t], gait["q"][t + 1], gait["u"][t], w, [mu], [gait["h"]]])
        yield t, z, th


def _residuals(rfun, gait, robot):
    """Residual of every knot with s1 = ϕ(q2) taken from the model under test (the `imp` rows read s1 − ϕ(q2))."""
    nq, nu, nw, nc, nb, o = _layout(robot)
    out = []
    for t, z, th in _gait_points(gait, robot):
        z[o[4]:o[5]] = -rfun(z, th, 0.0)[o[1]:o[2]]  # s1 = ϕ(q2)
        out.append(rfun(z, th, 0.0))
    return np.array(out)


@pytest.fixture(scope="module")
def generated(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("gen"))
    cache = {}

    def get(robot):
        if robot not in cache:
            cache[robot] = GeneratedResidual(robot, d)
        return cache[robot]
    return get


def _oracle_r(robot):
    from oracle.residual import get_residual
    return get_residual(robot).r


@pytest.mark.parametrize("side", ["product", "oracle"])
@pytest.mark.parametrize("robot,fixture,tol", [
    ("quadruped", "quadruped", 1e-4),                      # test/simulator/quadruped.jl:16-19 (gait2)
    ("flamingo", "flamingo", 1e-4),                        # test/simulator/flamingo.jl (gait_forward_36_4)
])
def test_gait_residual_identity(generated, side, robot, fixture, tol):
    gait = load_gait(fixture)
    rfun = generated(robot).r if side == "product" else _oracle_r(robot)
    r = _residuals(rfun, gait, robot)
    assert np.linalg.norm(r, axis=1).max() < tol


@pytest.mark.parametrize("side", ["product", "oracle"])
def test_hopper_gait_identity(generated, side):
    """examples/hopper/flat.jl's `gait_forward.jld2` (a serialized ContactTraj, z and θ stored in the file).  The
    DYNAMICS rows — masses, inertias, gravity, B, the contact Jacobian, (q, u, γ, b) of the gait — vanish (< 5e-5 on the
    body rows everywhere; on the leg-length row < 1e-4 on all but a handful of knots around the touch-down, worst 4.2e-3).  The slack
    entries (s1, ψ, η, s2) of the stored z are not those of `pack_z`, so the contact rows are not part of this pin."""
    lin = load_lin("hopper_2D")
    rfun = generated("hopper_2D").r if side == "product" else _oracle_r("hopper_2D")
    R = np.array([rfun(lin["z0"][t], lin["th0"][t], 0.0) for t in range(lin["z0"].shape[0])])
    assert np.abs(R[:, :3]).max() < 5e-5
    leg = np.sort(np.abs(R[:, 3]))
    assert leg[-1] < 5e-3 and leg[-2] < 1e-3 and np.median(leg) < 1e-6 and (leg > 1e-4).sum() <= 12


@pytest.mark.parametrize("side", ["product", "oracle"])
def test_centroidal_pins(generated, side):
    robot = "centroidal_quadruped"
    nq = SIZES[robot][0]
    rfun = generated(robot).r if side == "product" else _oracle_r(robot)
    # (1) a stand pins mass, gravity, B_func (the body is held up through u) and the contact stack
    r = _residuals(rfun, load_gait("centroidal_quadruped_stand_euler_v0"), robot)
    assert np.abs(r[:, :nq]).max() < 1e-4, np.abs(r[:, :nq]).max()
    assert np.abs(r[:, :3]).max() < 1e-7     # translational rows: m_body·g against Σ u
    # (2) the gait of flat_trot.jl: stance half exact, swing half inconsistent AT THE GAIT LEVEL (module docstring)
    r = _residuals(rfun, load_gait(robot), robot)
    dyn = np.abs(r[:, :nq]).max(axis=1)
    lin_rows = np.r_[0:3, 6:18]  # body translation and the four feet; rows 3:6 are the body orientation
    assert np.abs(r[32:][:, lin_rows]).max() < 1e-11
    assert np.abs(r[32:, 3:6]).max() < 1e-4
    assert 0.20 < dyn[30] < 0.22 and 0.13 < dyn[31] < 0.14   # m_foot·Δv of the un-modelled touch-down
    assert dyn[:30].max() < 0.03
    # the misfit cannot be explained by the inertial parameters: best (Δmass_foot, Δdamping) fit of the foot rows
    g = load_gait(robot)
    h = g["h"]
    A, y = [], []
    for t in range(g["u"].shape[0]):
        v1, v2 = (g["q"][t + 1] - g["q"][t]) / h, (g["q"][t + 2] - g["q"][t + 1]) / h
        ez = np.tile([0.0, 0.0, 1.0], 4)
        A.append(np.stack([(v1 - v2)[6:18] - h * 9.81 * ez, -h * v2[6:18]], 1))
        y.append(r[t, 6:18])
    A, y = np.concatenate(A), np.concatenate(y)
    x = np.linalg.lstsq(A, -y, rcond=None)[0]
    assert np.abs(y + A @ x).max() > 0.15


def test_centroidal_undamped_pin():
    """`inplace_trot_v0.jld2` was generated with `centroidal_quadruped_undamped` (model.jl:220-230, joint friction 0):
    its translational dynamics rows vanish to round-off with the damping switched off, and do not with it on — which
    pins the joint-friction term of the variational dynamics (src/dynamics/model.jl:40) as well."""
    import oracle.residual as R
    from oracle.models import CentroidalQuadruped

    class Undamped(CentroidalQuadruped):
        def __init__(self):
            super().__init__()
            self.joint_friction = [0.0] * 18

    orig = R.get_model
    try:
        R.get_model = lambda name: Undamped() if name == "cq_undamped" else orig(name)
        res_u = R.Residual("cq_undamped")
    finally:
        R.get_model = orig
    gait = load_gait("centroidal_quadruped_inplace_trot_v0")
    ru = _residuals(res_u.r, gait, "centroidal_quadruped")
    rd = _residuals(_oracle_r("centroidal_quadruped"), gait, "centroidal_quadruped")
    assert np.abs(ru[:, :3]).max() < 1e-12
    assert np.abs(rd[:, :3]).max() > 5e-4
    assert np.abs(ru[:, 6:18]).max() < 3e-3 < np.abs(rd[:, 6:18]).max()


def test_planar_kinematics_identities():
    """src/dynamics/quadruped/test.jl:6-28 and src/dynamics/flamingo/test.jl: end-effector / centre-of-mass positions of
    every link written out as explicit sums of l·sin, l·cos, for BOTH transcriptions (oracle/models.py and the product's
    modelgen/robots.py)."""
    import cimpc_b200 as cb
    from oracle import models as om
    pm = cb.package.modelgen.robots if hasattr(cb.package, "modelgen") else None
    if pm is None:
        import importlib
        pm = importlib.import_module(cb.package.__name__ + ".modelgen.robots")
    rng = np.random.default_rng(0)
    for mod in (om, pm):
        m = mod.Quadruped()
        q = rng.random(11)
        s, c = np.sin, np.cos
        lt, lth, ll = 0.366, 0.2, 0.2
        feet = np.array([float(v) for v in m.kinematics(list(q))]).reshape(4, 2)
        want = np.array([
            [q[0] + lth * s(q[3]) + ll * s(q[4]), q[1] - lth * c(q[3]) - ll * c(q[4])],                                  # calf_1 ee
            [q[0] + lth * s(q[5]) + ll * s(q[6]), q[1] - lth * c(q[5]) - ll * c(q[6])],                                  # calf_2 ee
            [q[0] + lt * s(q[2]) + lth * s(q[7]) + ll * s(q[8]), q[1] - lt * c(q[2]) - lth * c(q[7]) - ll * c(q[8])],    # calf_3 ee
            [q[0] + lt * s(q[2]) + lth * s(q[9]) + ll * s(q[10]), q[1] - lt * c(q[2]) - lth * c(q[9]) - ll * c(q[10])],  # calf_4 ee
        ])
        assert np.abs(feet - want).max() < 1e-14
        # centres of mass of the nine links (d_torso = l/2 + 0.0127, d_thigh = 0.1 − 0.00323, d_calf = 0.1 − 0.006435)
        dt, dth, dl = 0.5 * lt + 0.0127, 0.5 * lth - 0.00323, 0.5 * ll - 0.006435
        com = [m._pos(list(q), ch) for ch, _, _ in m.bodies()]
        com = np.array([[float(a), float(b)] for a, b in com])
        want = np.array([
            [q[0] + dt * s(q[2]), q[1] - dt * c(q[2])],
            [q[0] + dth * s(q[3]), q[1] - dth * c(q[3])],
            [q[0] + lth * s(q[3]) + dl * s(q[4]), q[1] - lth * c(q[3]) - dl * c(q[4])],
            [q[0] + dth * s(q[5]), q[1] - dth * c(q[5])],
            [q[0] + lth * s(q[5]) + dl * s(q[6]), q[1] - lth * c(q[5]) - dl * c(q[6])],
            [q[0] + lt * s(q[2]) + dth * s(q[7]), q[1] - lt * c(q[2]) - dth * c(q[7])],
            [q[0] + lt * s(q[2]) + lth * s(q[7]) + dl * s(q[8]), q[1] - lt * c(q[2]) - lth * c(q[7]) - dl * c(q[8])],
            [q[0] + lt * s(q[2]) + dth * s(q[9]), q[1] - lt * c(q[2]) - dth * c(q[9])],
            [q[0] + lt * s(q[2]) + lth * s(q[9]) + dl * s(q[10]), q[1] - lt * c(q[2]) - lth * c(q[9]) - dl * c(q[10])],
        ])
        assert np.abs(com - want).max() < 1e-14
        # contact Jacobian = derivative of the kinematics (test.jl:30-60 uses ForwardDiff; here central differences)
        J = np.array([[float(v) for v in row] for row in m.J_func(list(q))])
        eps = 1e-6
        for j in range(11):
            qp, qm = q.copy(), q.copy()
            qp[j] += eps; qm[j] -= eps
            fd = (np.array([float(v) for v in m.kinematics(list(qp))]) - np.array([float(v) for v in m.kinematics(list(qm))])) / (2 * eps)
            assert np.abs(fd - J[:, j]).max() < 1e-8
        # flamingo: toe / heel of both feet (src/dynamics/flamingo/model.jl kinematics :355-362)
        f = mod.Flamingo()
        q = rng.random(9)
        lth, lc, lf, df = 0.42, 0.45, 0.1725, 0.0525
        pts = np.array([float(v) for v in f.kinematics(list(q))]).reshape(4, 2)
        want = np.array([
            [q[0] + lth * s(q[3]) + lc * s(q[4]) + lf * s(q[7]), q[1] - lth * c(q[3]) - lc * c(q[4]) - lf * c(q[7])],
            [q[0] + lth * s(q[3]) + lc * s(q[4]) - df * s(q[7]), q[1] - lth * c(q[3]) - lc * c(q[4]) + df * c(q[7])],
            [q[0] + lth * s(q[5]) + lc * s(q[6]) + lf * s(q[8]), q[1] - lth * c(q[5]) - lc * c(q[6]) - lf * c(q[8])],
            [q[0] + lth * s(q[5]) + lc * s(q[6]) - df * s(q[8]), q[1] - lth * c(q[5]) - lc * c(q[6]) + df * c(q[8])],
        ])
        assert np.abs(pts - want).max() < 1e-14
