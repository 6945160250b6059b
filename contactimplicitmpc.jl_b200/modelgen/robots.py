"""Robot models of the reference as sympy expressions — input of the product's CUDA code generator
(`modelgen/codegen.py`), the analogue of the reference's Symbolics pipeline
(`src/dynamics/code_gen_dynamics.jl`, `src/simulation/code_gen_simulation.jl`).

This file is deliberately self-contained (the product never imports `oracle/`; the oracle keeps its own
restatement, and tests/test_codegen.py checks the two against each other numerically).

Each class restates one `src/dynamics/<robot>/model.jl` of the reference
(dojo-sim/ContactImplicitMPC.jl @ 989c8e6): kinematics, `M_func`, `C_func` (or the
Lagrangian it is derived from, `src/dynamics/code_gen_dynamics.jl:38-51`), `B_func`,
`A_func`, `J_func`, `ϕ_func`, plus the per-robot contact-force / tangential-velocity
stacks (`src/simulation/contact_methods.jl:27-48` and the per-robot overrides).  Environments: the flat ones
`flat_2D_lc` / `flat_3D_lc` (`src/simulation/environments/flat.jl`: `surf = 0`, surface rotation = identity) and, for the
planar robots, the piecewise terrains of `src/simulation/environments/piecewise.jl` (`<robot>_piecewise` models,
`terrain.py`): the contact stacks then take the surface rotation of every contact (`src/simulator/environment.jl:68-103`)
as the pair (cos, sin) of its angle, supplied as symbols by `symbolic.py`.

All functions accept sympy symbols or plain floats (everything is built from
`sympy.sin/cos`, +, *), so the same code yields the symbolic residual (for Jacobians
and code generation) and numeric spot values.
"""
from __future__ import annotations

import math

import sympy as sp


def _zeros(r, c):
    return [[sp.Integer(0) for _ in range(c)] for _ in range(r)]


class Model:
    name = ""
    world = 2  # dim(env): 2 for R2, 3 for R3  (environment.jl:123-124)
    nq = nu = nw = nc = 0
    mu_world = 1.0
    joint_friction: list = []
    C_analytical = True  # generate_dynamics.jl: which robots derive C from the Lagrangian
    terrain = None       # name of a non-flat 2-D environment (terrain.py), None = flat

    # ---- sizes (src/simulation/index.jl:371-390, environment.jl:126-127) ----
    @property
    def nf(self):
        return 2 if self.world == 2 else 4

    @property
    def nb(self):
        return self.nc * self.nf

    @property
    def nz(self):
        return self.nq + 4 * self.nc + 2 * self.nb

    @property
    def ntheta(self):
        return 2 * self.nq + self.nu + self.nw + 2

    @property
    def ny(self):
        return 2 * self.nc + self.nb

    # ---- to be provided per robot ----
    def lagrangian(self, q, qd):
        return sp.Integer(0)

    def M_func(self, q):
        raise NotImplementedError

    def C_func(self, q, qd):
        raise NotImplementedError

    def B_func(self, q):
        raise NotImplementedError

    def A_func(self, q):
        raise NotImplementedError

    def J_func(self, q):
        raise NotImplementedError

    def phi_func(self, q):
        raise NotImplementedError

    # ---- friction mapping (environment.jl:105-113) ----
    def friction_mapping(self):
        if self.world == 2:
            return [[1, -1]]
        return [[1, 0, -1, 0], [0, 1, 0, -1]]

    # contact_forces (contact_methods.jl:27-33; per-robot forms e.g. quadruped/model.jl:472-479):
    # λ_i = rotation(env, k_i)ᵀ [m b_i; γ_i].  `rot[i] = (c, s)`: the world→surface rotation of contact i is
    # [[c, s], [−s, c]] with c = 1/√(1+s'²), s = s'/√(1+s'²), s' the surface slope under the contact
    # (environment.jl:81-96: ang = −atan(s')); None = flat terrain, rotation = I.
    def contact_forces(self, gamma, b, rot=None):
        m = self.friction_mapping()
        nf, ne = self.nf, self.world
        out = []
        for i in range(self.nc):
            bi = b[i * nf:(i + 1) * nf]
            tb = [sum(row[j] * bi[j] for j in range(nf)) for row in m]
            if rot is None:
                out += tb + [gamma[i]]
            else:
                assert ne == 2, "surface rotation: planar environments only"
                c, sn = rot[i]
                out += [c * tb[0] - sn * gamma[i], sn * tb[0] + c * gamma[i]]
        assert len(out) == self.nc * ne
        return out

    # velocity_stack (contact_methods.jl:42-48; quadruped/model.jl:481-493): tangential part of rotation(env, k_i) v_i
    def velocity_stack(self, q1, q2, h, rot=None):
        J = self.J_func(q2)
        ne = self.world
        dq = [(q2[j] - q1[j]) / h for j in range(self.nq)]
        v = [sum(J[i][j] * dq[j] for j in range(self.nq)) for i in range(self.nc * ne)]
        m = self.friction_mapping()
        out = []
        for i in range(self.nc):
            if rot is None:
                vt = v[i * ne:i * ne + ne - 1]
            else:
                c, sn = rot[i]
                vt = [c * v[2 * i] + sn * v[2 * i + 1]]
            # transpose(friction_mapping) * v_T  ==  [v_T; -v_T]
            for j in range(self.nf):
                out.append(sum(m[k][j] * vt[k] for k in range(ne - 1)))
        return out


# --------------------------------------------------------------------------------------
class Hopper2D(Model):
    """src/dynamics/hopper_2D/model.jl"""
    name = "hopper_2D"
    world = 2
    nq, nu, nw, nc = 4, 2, 2, 1
    mb, ml, Jb, Jl = 3.0, 0.3, 0.75, 0.075  # :86-89
    mu_world, g = 0.8, 9.81                 # :80-82
    joint_friction = [0.0] * 4              # :102

    def kinematics(self, q):  # :36-39
        return [q[0] + q[3] * sp.sin(q[2]), q[1] - q[3] * sp.cos(q[2])]

    def M_func(self, q):  # :42-47
        d = [self.mb + self.ml, self.mb + self.ml, self.Jb + self.Jl, self.ml]
        M = _zeros(4, 4)
        for i in range(4):
            M[i][i] = d[i]
        return M

    def C_func(self, q, qd):  # :49-54
        return [0, (self.mb + self.ml) * self.g, 0, 0]

    def phi_func(self, q):  # :56-59 (flat: surf = 0)
        return [self.kinematics(q)[1]]

    def J_func(self, q):  # :61-64
        return [[1, 0, q[3] * sp.cos(q[2]), sp.sin(q[2])],
                [0, 1, q[3] * sp.sin(q[2]), -sp.cos(q[2])]]

    def B_func(self, q):  # :66-69
        return [[0, 0, 1, 0], [-sp.sin(q[2]), sp.cos(q[2]), 0, 1]]

    def A_func(self, q):  # :71-74
        return [[1, 0, 0, 0], [0, 1, 0, 0]]


# --------------------------------------------------------------------------------------
class _PlanarLegged(Model):
    """Shared machinery for the planar articulated robots (quadruped, flamingo):
    bodies are described as chains  [(angle index, length) ...] hanging from (x, z)."""
    world = 2
    nw = 2
    C_analytical = False
    g = 9.81

    # list of bodies: (chain of (qi, length, sign) up to the parent end, (qi, d_com, sign), mass, inertia)
    def bodies(self):
        raise NotImplementedError

    def contact_chains(self):
        raise NotImplementedError

    @staticmethod
    def _pos(q, chain):
        x, z = q[0], q[1]
        for (i, r, s) in chain:
            x = x + s * r * sp.sin(q[i])
            z = z - s * r * sp.cos(q[i])
        return [x, z]

    def _jac(self, q, chain):
        J = _zeros(2, self.nq)
        J[0][0] = sp.Integer(1)
        J[1][1] = sp.Integer(1)
        for (i, r, s) in chain:
            J[0][i] = J[0][i] + s * r * sp.cos(q[i])
            J[1][i] = J[1][i] + s * r * sp.sin(q[i])
        return J

    def lagrangian(self, q, qd):
        L = sp.Integer(0)
        for chain, m, Jin in self.bodies():
            p = self._pos(q, chain)
            Jc = self._jac(q, chain)
            v = [sum(Jc[a][j] * qd[j] for j in range(self.nq)) for a in range(2)]
            L = L + 0.5 * m * (v[0] * v[0] + v[1] * v[1])
            L = L + 0.5 * Jin * qd[chain[-1][0]] ** 2
            L = L - m * self.g * p[1]
        return L

    def M_func(self, q):
        M = _zeros(self.nq, self.nq)
        for chain, m, Jin in self.bodies():
            i = chain[-1][0]
            M[i][i] = M[i][i] + Jin
            Jc = self._jac(q, chain)
            for a in range(self.nq):
                for b in range(self.nq):
                    M[a][b] = M[a][b] + m * (Jc[0][a] * Jc[0][b] + Jc[1][a] * Jc[1][b])
        return M

    def kinematics(self, q):
        out = []
        for chain in self.contact_chains():
            out += self._pos(q, chain)
        return out

    def phi_func(self, q):
        return [self._pos(q, chain)[1] for chain in self.contact_chains()]

    def J_func(self, q):
        J = []
        for chain in self.contact_chains():
            J += self._jac(q, chain)
        return J

    def A_func(self, q):
        A = _zeros(2, self.nq)
        A[0][0] = 1
        A[1][1] = 1
        return A


class Quadruped(_PlanarLegged):
    """src/dynamics/quadruped/model.jl (planar Unitree-A1-like, :510-528).
    q = (x, z, torso, thigh1, calf1, thigh2, calf2, thigh3, calf3, thigh4, calf4)."""
    name = "quadruped"
    nq, nu, nc = 11, 8, 4
    mu_world = 1.0
    mu_joint = 0.1
    m_torso = 4.713 + 4 * 0.696
    m_thigh, m_leg = 1.013, 0.166
    J_torso = 0.01683 + 4 * 0.696 * 0.183 ** 2.0
    J_thigh, J_leg = 0.00552, 0.00299
    l_torso, l_thigh, l_leg = 0.183 * 2, 0.2, 0.2
    d_torso = 0.5 * l_torso + 0.0127
    d_thigh = 0.5 * l_thigh - 0.00323
    d_leg = 0.5 * l_leg - 0.006435

    def __init__(self, payload=False):
        self.joint_friction = [0.0] * 3 + [self.mu_joint] * 8  # :574
        if payload:  # quadruped_payload :576-590
            self.m_torso = self.m_torso + 3.0
            self.J_torso = self.J_torso + 0.03
            self.name = "quadruped_payload"

    def bodies(self):
        lt, lth = self.l_torso, self.l_thigh
        T, TH, LG = (self.m_torso, self.J_torso), (self.m_thigh, self.J_thigh), (self.m_leg, self.J_leg)
        # index (0-based): torso 2, thigh1 3, calf1 4, thigh2 5, calf2 6, thigh3 7, calf3 8, thigh4 9, calf4 10
        return [
            ([(2, self.d_torso, 1)], *T),                                  # torso (kinematics_1 :79-107)
            ([(3, self.d_thigh, 1)], *TH),                                 # thigh 1
            ([(3, lth, 1), (4, self.d_leg, 1)], *LG),                      # calf 1 (kinematics_2 :140-185)
            ([(5, self.d_thigh, 1)], *TH),                                 # thigh 2
            ([(5, lth, 1), (6, self.d_leg, 1)], *LG),                      # calf 2
            ([(2, lt, 1), (7, self.d_thigh, 1)], *TH),                     # thigh 3 (hangs from torso end)
            ([(2, lt, 1), (7, lth, 1), (8, self.d_leg, 1)], *LG),          # calf 3 (kinematics_3 :241-270)
            ([(2, lt, 1), (9, self.d_thigh, 1)], *TH),                     # thigh 4
            ([(2, lt, 1), (9, lth, 1), (10, self.d_leg, 1)], *LG),         # calf 4
        ]

    def contact_chains(self):  # kinematics :389-396
        lt, lth, ll = self.l_torso, self.l_thigh, self.l_leg
        return [
            [(3, lth, 1), (4, ll, 1)],
            [(5, lth, 1), (6, ll, 1)],
            [(2, lt, 1), (7, lth, 1), (8, ll, 1)],
            [(2, lt, 1), (9, lth, 1), (10, ll, 1)],
        ]

    def B_func(self, q):  # :451-460
        B = _zeros(8, 11)
        rows = [(2, 3), (3, 4), (2, 5), (5, 6), (2, 7), (7, 8), (2, 9), (9, 10)]
        for r, (neg, pos) in enumerate(rows):
            B[r][neg] = -1
            B[r][pos] = 1
        return B


class Flamingo(_PlanarLegged):
    """src/dynamics/flamingo/model.jl.
    q = (x, z, torso, thigh1, calf1, thigh2, calf2, foot1, foot2); contacts toe1, heel1, toe2, heel2."""
    name = "flamingo"
    nq, nu, nc = 9, 6, 4
    mu_world = 0.9
    m_torso, m_thigh, m_calf, m_foot = 12.0, 0.4598, 0.306, 0.3466  # :457-460
    l_torso, l_thigh, l_calf, l_foot = 0.385, 0.42, 0.45, 0.1725    # :462-465
    d_torso, d_thigh, d_calf, d_foot = 0.20, 0.42 / 2, 0.45 / 2, 0.0525
    J_torso, J_thigh, J_calf, J_foot = 0.10, 0.01256, 0.00952, 0.0015

    def __init__(self):
        self.joint_friction = [0.0] * 9  # :487

    def bodies(self):
        lth, lc = self.l_thigh, self.l_calf
        cb = 0.5 * (self.l_foot - self.d_foot)  # kinematics_3 :194,203
        return [
            ([(2, self.d_torso, -1)], self.m_torso, self.J_torso),  # torso points up (:66-73)
            ([(3, self.d_thigh, 1)], self.m_thigh, self.J_thigh),
            ([(3, lth, 1), (4, self.d_calf, 1)], self.m_calf, self.J_calf),
            ([(3, lth, 1), (4, lc, 1), (7, cb, 1)], self.m_foot, self.J_foot),
            ([(5, self.d_thigh, 1)], self.m_thigh, self.J_thigh),
            ([(5, lth, 1), (6, self.d_calf, 1)], self.m_calf, self.J_calf),
            ([(5, lth, 1), (6, lc, 1), (8, cb, 1)], self.m_foot, self.J_foot),
        ]

    def contact_chains(self):  # kinematics :355-362
        lth, lc = self.l_thigh, self.l_calf
        return [
            [(3, lth, 1), (4, lc, 1), (7, self.l_foot, 1)],    # toe 1
            [(3, lth, 1), (4, lc, 1), (7, self.d_foot, -1)],   # heel 1
            [(5, lth, 1), (6, lc, 1), (8, self.l_foot, 1)],    # toe 2
            [(5, lth, 1), (6, lc, 1), (8, self.d_foot, -1)],   # heel 2
        ]

    def B_func(self, q):  # :412-419
        B = _zeros(6, 9)
        rows = [(2, 3), (3, 4), (2, 5), (5, 6), (4, 7), (6, 8)]
        for r, (neg, pos) in enumerate(rows):
            B[r][neg] = -1
            B[r][pos] = 1
        return B


# --------------------------------------------------------------------------------------
def _skew(x):
    return [[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]]


def euler_rotation_matrix(t):  # src/dynamics/euler.jl:3-11
    a, b, c = t
    ca, sa, cb, sb, cc, sc = sp.cos(a), sp.sin(a), sp.cos(b), sp.sin(b), sp.cos(c), sp.sin(c)
    return [[ca * cb, ca * sb * sc - sa * cc, ca * sb * cc + sa * sc],
            [sa * cb, sa * sb * sc + ca * cc, sa * sb * cc - ca * sc],
            [-sb, cb * sc, cb * cc]]


class CentroidalQuadruped(Model):
    """src/dynamics/centroidal_quadruped/model.jl.  q = (p, euler, f1, f2, f3, f4)."""
    name = "centroidal_quadruped"
    world = 3
    nq, nu, nw, nc = 18, 12, 3, 4
    mu_world, g = 0.3, 9.81
    mass_body, mass_foot = 13.5, 0.2
    inertia = [0.0178533 * 10.0, 0.0377999 * 10.0, 0.0456542 * 10.0]  # :196-203

    def __init__(self, mass_body=None, inertia=None):
        # "payload" variants of config 5 perturb these two; defaults = reference
        if mass_body is not None:
            self.mass_body = mass_body
        if inertia is not None:
            self.inertia = list(inertia)
        self.joint_friction = [10.0] * 3 + [30.0] * 3 + [10.0] * 12  # :213 (μ_joint = 1)

    def kinematics(self, q):
        return list(q[6:18])

    def M_func(self, q):  # :66-73
        d = [self.mass_body] * 3 + list(self.inertia) + [self.mass_foot] * 12
        M = _zeros(18, 18)
        for i in range(18):
            M[i][i] = d[i]
        return M

    def C_func(self, q, qd):  # :75-84
        w = qd[3:6]
        Iw = [self.inertia[i] * w[i] for i in range(3)]
        S = _skew(w)
        tau = [sum(S[i][j] * Iw[j] for j in range(3)) for i in range(3)]
        out = [0, 0, self.mass_body * self.g] + tau
        for _ in range(4):
            out += [0, 0, self.mass_foot * self.g]
        return out

    def phi_func(self, q):  # :86-94
        return [q[8], q[11], q[14], q[17]]

    def B_func(self, q):  # :96-117  (returns nu x nq = transpose of the 18x12 block matrix)
        R = euler_rotation_matrix(q[3:6])
        Bt = _zeros(18, 12)
        for f in range(4):
            r = [q[6 + 3 * f + i] - q[i] for i in range(3)]
            S = _skew(r)
            for i in range(3):
                Bt[i][3 * f + i] = 1
                Bt[6 + 3 * f + i][3 * f + i] = -1
                for j in range(3):
                    # transpose(R) * skew(r)
                    Bt[3 + i][3 * f + j] = sum(R[k][i] * S[k][j] for k in range(3))
        return [[Bt[i][j] for i in range(18)] for j in range(12)]

    def A_func(self, q):  # :119-123
        A = _zeros(3, 18)
        for i in range(3):
            A[i][i] = 1
        return A

    def J_func(self, q):  # :125-134
        J = _zeros(12, 18)
        for i in range(12):
            J[i][6 + i] = 1
        return J


def get_model(name: str) -> Model:
    if name.endswith("_piecewise"):
        # get_simulation(<robot>, "piecewise1_2D_lc", "piecewise", approx = true)  (examples/*/piecewise*.jl:11-14)
        m = get_model(name[:-len("_piecewise")])
        assert m.world == 2, "piecewise terrains are planar"
        m.name = name
        m.terrain = "piecewise1_2D_lc"
        return m
    if name == "hopper_2D":
        return Hopper2D()
    if name == "quadruped":
        return Quadruped()
    if name == "quadruped_payload":
        return Quadruped(payload=True)
    if name == "flamingo":
        return Flamingo()
    if name == "centroidal_quadruped":
        return CentroidalQuadruped()
    if name == "centroidal_quadruped_payload":
        # BASELINE config 5 ("centroidal_quadruped + payload").  The reference ships no centroidal payload model; this
        # one carries the payload of `quadruped_payload` (src/dynamics/quadruped/model.jl:530-531: 3 kg, 0.03 kg m²) on
        # the body: mass_body + 3.0, inertia_body + 0.03·I.
        m = CentroidalQuadruped(mass_body=13.5 + 3.0, inertia=[0.0178533 * 10.0 + 0.03, 0.0377999 * 10.0 + 0.03,
                                                                  0.0456542 * 10.0 + 0.03])
        m.name = "centroidal_quadruped_payload"
        return m
    raise KeyError(name)
