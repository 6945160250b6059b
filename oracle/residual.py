"""Nonlinear contact residual r(z, θ, κ) and its Jacobians rz, rθ (CPU oracle, sympy → numpy).

TEST INFRASTRUCTURE ONLY.  Restates
  * the variational-midpoint dynamics `dynamics(model, h, q0, q1, u1, w1, Λ1, q2)`
    (`src/dynamics/model.jl:18-41`) with `C` either analytical or derived from the Lagrangian
    (`src/dynamics/code_gen_dynamics.jl:38-51`),
  * the LinearizedCone residual (`src/simulation/simulation.jl:133-158`),
  * the z / θ / r layouts of `src/simulation/index.jl`,
  * the code generation step `generate_residual_expressions`
    (`src/simulation/code_gen_simulation.jl:114-168`): rz = ∂r/∂z, rθ = ∂r/∂θ, dense.

z = [q2; γ1; b1; ψ1; s1; η1; s2]   θ = [q0; q1; u1; w1; μ; h]
r = [dyn; imp; mdp; fri; bimp; bmdp; bfri]
"""
from __future__ import annotations

import functools

import numpy as np
import sympy as sp

from .models import Model, get_model


class Index:
    """0-based index maps of `src/simulation/index.jl` (z :13-107, θ :117-178, r :187-269)."""

    def __init__(self, m: Model):
        nq, nu, nw, nc, nb = m.nq, m.nu, m.nw, m.nc, m.nb
        o = 0
        self.q2 = np.arange(o, o + nq); o += nq
        self.g1 = np.arange(o, o + nc); o += nc
        self.b1 = np.arange(o, o + nb); o += nb
        self.psi1 = np.arange(o, o + nc); o += nc
        self.s1 = np.arange(o, o + nc); o += nc
        self.eta1 = np.arange(o, o + nb); o += nb
        self.s2 = np.arange(o, o + nc); o += nc
        self.nz = o
        o = 0
        self.q0 = np.arange(o, o + nq); o += nq
        self.q1 = np.arange(o, o + nq); o += nq
        self.u1 = np.arange(o, o + nu); o += nu
        self.w1 = np.arange(o, o + nw); o += nw
        self.mu = np.arange(o, o + 1); o += 1
        self.h = np.arange(o, o + 1); o += 1
        self.ntheta = o
        # residual rows share the z offsets (index.jl:187-269)
        self.dyn, self.imp, self.mdp, self.fri = self.q2, self.g1, self.b1, self.psi1
        self.bimp, self.bmdp, self.bfri = self.s1, self.eta1, self.s2
        # linearization groups (index.jl:289-327)
        self.x = self.q2
        self.y1 = np.concatenate([self.g1, self.b1, self.psi1])
        self.y2 = np.concatenate([self.s1, self.eta1, self.s2])
        self.rst = np.concatenate([self.imp, self.mdp, self.fri])
        self.bil = np.concatenate([self.bimp, self.bmdp, self.bfri])
        self.alt = self.imp


def _matvec_T(Mt, v, n):
    # transpose(M) * v with M given as rows
    return [sum(Mt[k][i] * v[k] for k in range(len(v))) for i in range(n)]


def symbolic_C(m: Model, q, qd):
    """C(q, q̇): analytical, or  (∂²L/∂q̇∂q) q̇ − ∂L/∂q  (code_gen_dynamics.jl:44-50)."""
    if m.C_analytical:
        return [sp.sympify(c) for c in m.C_func(q, qd)]
    L = m.lagrangian(q, qd)
    dLq = [sp.diff(L, qi) for qi in q]
    dLqd = [sp.diff(L, qdi) for qdi in qd]
    C = []
    for i in range(m.nq):
        c = sum(sp.diff(dLqd[i], q[j]) * qd[j] for j in range(m.nq)) - dLq[i]
        C.append(c)
    return C


def symbolic_residual(m: Model, surface=None):
    """`surface`: None on flat ground, else three lists of nc symbols (f, c, s): height of the surface under every
    contact point and (cos, sin) of its world→surface rotation [[c, s], [−s, c]] (planar environments,
    src/simulator/environment.jl:81-96).  They enter as in the per-robot contact stacks of the reference
    (e.g. src/dynamics/quadruped/model.jl:472-493, hopper_2D/model.jl:74-86): λ_i = Rᵀ[m b_i; γ_i], tangential
    velocity = (R v_i)[1], ϕ_i = p_z − surf(p_x) (flamingo/model.jl:391-401)."""
    nq = m.nq
    idx = Index(m)
    z = sp.symbols(f"z0:{idx.nz}", real=True)
    th = sp.symbols(f"t0:{idx.ntheta}", real=True)
    kappa = sp.Symbol("kappa", real=True)

    q2 = [z[i] for i in idx.q2]
    g1 = [z[i] for i in idx.g1]
    b1 = [z[i] for i in idx.b1]
    psi1 = [z[i] for i in idx.psi1]
    s1 = [z[i] for i in idx.s1]
    eta1 = [z[i] for i in idx.eta1]
    s2 = [z[i] for i in idx.s2]
    q0 = [th[i] for i in idx.q0]
    q1 = [th[i] for i in idx.q1]
    u1 = [th[i] for i in idx.u1]
    w1 = [th[i] for i in idx.w1]
    mu = th[idx.mu[0]]
    h = th[idx.h[0]]

    # --- dynamics (model.jl:18-41) ---
    qs = sp.symbols(f"q_0:{nq}", real=True)
    qds = sp.symbols(f"qd_0:{nq}", real=True)
    Csym = symbolic_C(m, qs, qds)
    Msym = m.M_func(qs)

    def lagr_derivs(qm, vm):  # model.jl:12-16
        sub = dict(zip(qs, qm))
        sub.update(dict(zip(qds, vm)))
        D1L = [-sp.sympify(c).xreplace(sub) for c in Csym]
        D2L = [sum(sp.sympify(Msym[i][j]).xreplace(sub) * vm[j] for j in range(nq)) for i in range(nq)]
        return D1L, D2L

    qm1 = [0.5 * (q0[i] + q1[i]) for i in range(nq)]
    vm1 = [(q1[i] - q0[i]) / h for i in range(nq)]
    qm2 = [0.5 * (q1[i] + q2[i]) for i in range(nq)]
    vm2 = [(q2[i] - q1[i]) / h for i in range(nq)]
    D1L1, D2L1 = lagr_derivs(qm1, vm1)
    D1L2, D2L2 = lagr_derivs(qm2, vm2)
    lam = m.contact_forces(g1, b1)
    if surface is not None:
        assert m.world == 2
        _, cs, sn = surface
        lam = [w for i in range(m.nc) for w in (cs[i] * lam[2 * i] - sn[i] * lam[2 * i + 1],
                                                 sn[i] * lam[2 * i] + cs[i] * lam[2 * i + 1])]
    Lam = _matvec_T(m.J_func(q2), lam, nq)
    Bu = _matvec_T(m.B_func(qm2), u1, nq)
    Aw = _matvec_T(m.A_func(qm2), w1, nq)
    d = [0.5 * h * D1L1[i] + D2L1[i] + 0.5 * h * D1L2[i] - D2L2[i] + Bu[i] + Aw[i] + Lam[i]
         - h * m.joint_friction[i] * vm2[i] for i in range(nq)]

    # --- contact (simulation.jl:133-158) ---
    phi = m.phi_func(q2)
    vT = m.velocity_stack(q1, q2, h)
    nc, nf = m.nc, m.nf
    if surface is not None:
        f, cs, sn = surface
        phi = [phi[i] - f[i] for i in range(nc)]
        Jc = m.J_func(q2)
        v = [sum(Jc[a][j] * (q2[j] - q1[j]) / h for j in range(nq)) for a in range(2 * nc)]
        vt = [cs[i] * v[2 * i] + sn[i] * v[2 * i + 1] for i in range(nc)]
        vT = [w for i in range(nc) for w in (vt[i], -vt[i])]
    r = list(d)
    r += [s1[i] - phi[i] for i in range(nc)]
    r += [eta1[i] - vT[i] - psi1[i // nf] for i in range(m.nb)]          # η1 − vT − Eᵀψ1
    r += [s2[i] - (mu * g1[i] - sum(b1[i * nf:(i + 1) * nf])) for i in range(nc)]  # s2 − (μγ1 − E b1)
    r += [g1[i] * s1[i] - kappa for i in range(nc)]
    r += [b1[i] * eta1[i] - kappa for i in range(m.nb)]
    r += [psi1[i] * s2[i] - kappa for i in range(nc)]
    assert len(r) == idx.nz
    return z, th, kappa, [sp.sympify(e) for e in r]


class Residual:
    """Numeric r, rz, rθ for one robot (mirrors `ResidualMethods` r!, rz!, rθ!)."""

    def __init__(self, name: str):
        self.model = get_model(name)
        self.idx = Index(self.model)
        z, th, kappa, r = symbolic_residual(self.model)
        self.sym = (z, th, kappa, r)
        rvec = sp.Matrix(r)
        rz = rvec.jacobian(sp.Matrix(z))
        rth = rvec.jacobian(sp.Matrix(th))
        self._r = sp.lambdify([z, th, kappa], rvec, modules="numpy", cse=True)
        self._rz = sp.lambdify([z, th], rz, modules="numpy", cse=True)
        self._rth = sp.lambdify([z, th], rth, modules="numpy", cse=True)
        self.sym_rz, self.sym_rth = rz, rth

    def r(self, z, th, kappa):
        return np.asarray(self._r(z, th, kappa), dtype=np.float64).reshape(-1)

    def rz(self, z, th):
        return np.asarray(self._rz(z, th), dtype=np.float64)

    def rth(self, z, th):
        return np.asarray(self._rth(z, th), dtype=np.float64)


class TerrainResidual:
    """r, rz, rθ of a planar robot on a piecewise terrain, as the reference's `approx = true` simulations define them
    (`get_simulation(robot, "piecewise1_2D_lc", "piecewise", approx = true)`; src/simulation/residual_approx.jl:14-99):
    the Jacobians hold the surface ROTATION under every contact fixed (`dcf`, `vsq2`, `vsq1h` take the kinematics `k` as
    an undifferentiated argument) and differentiate everything else, ϕ included (`rcz`, `rcθ`).

    The Jacobians here are complex-step derivatives of the residual with the rotation frozen at the evaluation point
    and the surface height continued along its tangent (oracle/terrain.py) — exact to round-off, and a different route
    from the product's symbolic one (modelgen/symbolic.py), so the two check each other."""

    def __init__(self, name: str, terrain: str = "piecewise1_2D_lc"):
        from .terrain import get_terrain
        assert name.endswith("_piecewise")
        self.model = get_model(name[:-len("_piecewise")])
        self.idx = Index(self.model)
        self.terrain = get_terrain(terrain)
        nc = self.model.nc
        surface = (sp.symbols(f"sf0:{nc}", real=True), sp.symbols(f"sc0:{nc}", real=True), sp.symbols(f"ss0:{nc}", real=True))
        z, th, kappa, r = symbolic_residual(self.model, surface)
        flat = [w for grp in surface for w in grp]
        self._r = sp.lambdify([z, th, kappa, flat], sp.Matrix(r), modules="numpy", cse=True)
        q2 = [z[i] for i in self.idx.q2]
        kin = self.model.kinematics(q2)
        self._px = sp.lambdify([q2], [kin[2 * i] for i in range(nc)], modules="numpy")

    def _surface(self, z, rot_at=None):
        """[f; c; s] at z; the rotation is evaluated at `rot_at` (default: z itself)."""
        px = np.asarray(self._px(np.asarray(z)[self.idx.q2]))
        pr = px if rot_at is None else np.asarray(self._px(np.asarray(rot_at)[self.idx.q2]))
        f = [self.terrain.height(x) for x in px]
        cs = [self.terrain.rotation(x) for x in pr]
        return np.array(f + [c for c, _ in cs] + [s_ for _, s_ in cs])

    def r(self, z, th, kappa):
        return np.asarray(self._r(z, th, kappa, self._surface(z)), dtype=np.float64).reshape(-1)

    def _complex_step(self, z, th, wrt_z: bool):
        z = np.asarray(z, dtype=np.float64)
        th = np.asarray(th, dtype=np.float64)
        n = z.size if wrt_z else th.size
        J = np.zeros((z.size, n))
        eps = 1e-30
        for j in range(n):
            zc, tc = z.astype(np.complex128), th.astype(np.complex128)
            (zc if wrt_z else tc)[j] += 1j * eps
            rc = np.asarray(self._r(zc, tc, 0.0, self._surface(zc, rot_at=z)), dtype=np.complex128).reshape(-1)
            J[:, j] = rc.imag / eps
        return J

    def rz(self, z, th):
        return self._complex_step(z, th, True)

    def rth(self, z, th):
        return self._complex_step(z, th, False)


@functools.lru_cache(maxsize=None)
def get_residual(name: str):
    if name.endswith("_piecewise"):
        return TerrainResidual(name)
    return Residual(name)
