"""Symbolic nonlinear contact residual r(z, θ, κ) and its z-Jacobian for one robot — what the reference
generates with Symbolics in `generate_residual_expressions` (src/simulation/code_gen_simulation.jl:114-168)
from `residual` (src/simulation/simulation.jl:133-158) and `dynamics` (src/dynamics/model.jl:18-41).

z = [q2; γ1; b1; ψ1; s1; η1; s2]   θ = [q0; q1; u1; w1; μ; h]   r = [dyn; imp; mdp; fri; bimp; bmdp; bfri]
(src/simulation/index.jl:13-107, 117-178, 187-269)."""
from __future__ import annotations

import sympy as sp

from .robots import Model


def _matvec_T(Mt, v, n):
    return [sum(Mt[k][i] * v[k] for k in range(len(v))) for i in range(n)]


def symbolic_C(m: Model, q, qd):
    """C(q, q̇): analytical, or (∂²L/∂q̇∂q) q̇ − ∂L/∂q  (src/dynamics/code_gen_dynamics.jl:38-51)."""
    if m.C_analytical:
        return [sp.sympify(c) for c in m.C_func(q, qd)]
    L = m.lagrangian(q, qd)
    dLq = [sp.diff(L, qi) for qi in q]
    dLqd = [sp.diff(L, qdi) for qdi in qd]
    return [sum(sp.diff(dLqd[i], q[j]) * qd[j] for j in range(m.nq)) - dLq[i] for i in range(m.nq)]


def symbolic_residual(m: Model):
    """Returns (z symbols, θ symbols, κ symbol, list of nz residual expressions)."""
    return symbolic_residual_env(m)[:4]


def symbolic_residual_env(m: Model):
    """Returns (z, θ, κ, r, r_diff, terr).

    Flat environment: r_diff is r, terr is None.  On a terrain (`m.terrain`), contact i sees the surface through four
    run-time values, symbols here: height `tf{i}` and slope `tg{i}` under the contact point, and (`tc{i}`, `ts{i}`) =
    (cos, sin) of the surface rotation.  `terr["px"][i]` is the x coordinate (an expression of q2) they are evaluated
    at.  The Jacobians are those of the reference's `approx = true` simulations (`rz_approx!`, `rθ_approx!`,
    src/simulation/residual_approx.jl:14-99): the rotation is held constant (`dcf`, `vsq2` take the kinematics `k` as a
    separate, undifferentiated argument) and ϕ is differentiated exactly (`rcz`, ∂ϕ_i/∂q2 = ∂p_z/∂q2 − s'(p_x) ∂p_x/∂q2).
    `r_diff` is the residual to differentiate: it carries the surface as its tangent line `tf + tg (p_x(q2) − x0_i)`,
    whose z-derivative is that expression whatever the expansion points `x0_i`."""
    nq, nu, nw, nc, nb, nf = m.nq, m.nu, m.nw, m.nc, m.nb, m.nf
    z = sp.symbols(f"z0:{m.nz}", real=True)
    th = sp.symbols(f"t0:{m.ntheta}", real=True)
    kappa = sp.Symbol("kappa", real=True)
    o = 0
    q2 = z[o:o + nq]; o += nq
    g1 = z[o:o + nc]; o += nc
    b1 = z[o:o + nb]; o += nb
    psi1 = z[o:o + nc]; o += nc
    s1 = z[o:o + nc]; o += nc
    eta1 = z[o:o + nb]; o += nb
    s2 = z[o:o + nc]; o += nc
    o = 0
    q0 = th[o:o + nq]; o += nq
    q1 = th[o:o + nq]; o += nq
    u1 = th[o:o + nu]; o += nu
    w1 = th[o:o + nw]; o += nw
    mu = th[o]; h = th[o + 1]

    qs = sp.symbols(f"q_0:{nq}", real=True)
    qds = sp.symbols(f"qd_0:{nq}", real=True)
    Csym = symbolic_C(m, qs, qds)
    Msym = m.M_func(qs)

    def lagr(qm, vm):  # lagrangian_derivatives, model.jl:12-16
        sub = dict(zip(qs, qm))
        sub.update(dict(zip(qds, vm)))
        D1L = [-sp.sympify(c).xreplace(sub) for c in Csym]
        D2L = [sum(sp.sympify(Msym[i][j]).xreplace(sub) * vm[j] for j in range(nq)) for i in range(nq)]
        return D1L, D2L

    qm1 = [0.5 * (q0[i] + q1[i]) for i in range(nq)]
    vm1 = [(q1[i] - q0[i]) / h for i in range(nq)]
    qm2 = [0.5 * (q1[i] + q2[i]) for i in range(nq)]
    vm2 = [(q2[i] - q1[i]) / h for i in range(nq)]
    D1L1, D2L1 = lagr(qm1, vm1)
    D1L2, D2L2 = lagr(qm2, vm2)
    terr = None
    rot = None
    if m.terrain is not None:
        assert m.world == 2
        k = m.kinematics(q2)
        terr = {"name": m.terrain, "px": [sp.sympify(k[2 * i]) for i in range(nc)],
                "tf": sp.symbols(f"tf0:{nc}", real=True), "tg": sp.symbols(f"tg0:{nc}", real=True),
                "tc": sp.symbols(f"tc0:{nc}", real=True), "ts": sp.symbols(f"ts0:{nc}", real=True),
                "x0": sp.symbols(f"tx0:{nc}", real=True)}
        rot = list(zip(terr["tc"], terr["ts"]))
    lam = m.contact_forces(g1, b1, rot)
    Lam = _matvec_T(m.J_func(q2), lam, nq)
    Bu = _matvec_T(m.B_func(qm2), u1, nq)
    Aw = _matvec_T(m.A_func(qm2), w1, nq)
    r = [0.5 * h * D1L1[i] + D2L1[i] + 0.5 * h * D1L2[i] - D2L2[i] + Bu[i] + Aw[i] + Lam[i]
         - h * m.joint_friction[i] * vm2[i] for i in range(nq)]
    phi = m.phi_func(q2)  # height of the contact points over z = 0
    vT = m.velocity_stack(q1, q2, h, rot)
    n_dyn = len(r)
    r += [s1[i] - phi[i] + (terr["tf"][i] if terr else 0) for i in range(nc)]
    r += [eta1[i] - vT[i] - psi1[i // nf] for i in range(nb)]
    r += [s2[i] - (mu * g1[i] - sum(b1[i * nf:(i + 1) * nf])) for i in range(nc)]
    r += [g1[i] * s1[i] - kappa for i in range(nc)]
    r += [b1[i] * eta1[i] - kappa for i in range(nb)]
    r += [psi1[i] * s2[i] - kappa for i in range(nc)]
    assert len(r) == m.nz
    r = [sp.sympify(e) for e in r]
    r_diff = r
    if terr:
        r_diff = list(r)
        for i in range(nc):
            r_diff[n_dyn + i] = r[n_dyn + i] + terr["tg"][i] * (terr["px"][i] - terr["x0"][i])
    return z, th, kappa, r, r_diff, terr
