"""sympy → CUDA/C++ code generator for the nonlinear residual `r!` and its Jacobian `rz!` of one robot
(the product's replacement for the reference's `residual.jld2` / `jacobians.jld2` expression files,
src/simulation/code_gen_simulation.jl:114-197).

    python -m contactimplicitmpc_jl_b200.modelgen.codegen            (via cimpc_b200.package)
    python contactimplicitmpc.jl_b200/modelgen/codegen.py [robot ...]

writes `csrc/gen/residual_<robot>.h`: common-subexpression-eliminated straight-line code,
    eval_r (z, θ, κ, r)      all nz residual rows
    eval_rz(z, θ, J)         the structural non-zeros of ∂r/∂z, in the order of RZ_ROW / RZ_COL
with accessor functors for inputs/outputs, so the same text serves the thread-per-rollout CUDA kernels
(strided global memory) and the host-side parity shim (plain arrays).

Both functions are emitted as NS independent *slices* (`eval_r_<s>`, `eval_rz_<s>`): the outputs are dealt
to the slices by expression size (largest first onto the lightest slice) and every slice is CSE'd on its
own.  A CTA of the simulator kernel runs slice s on warp s (same 32 rollouts, a different part of the
outputs), which cuts the latency of the generated phases NS-fold and keeps each slice's temporaries inside
the register budget (the single-function form needed > 255 registers and spilled).  `eval_r` / `eval_rz`
call all slices in turn.

Trigonometry is hoisted: every distinct sin/cos argument of the residual and its Jacobian (they are
linear in z and θ: link angles at q2 and at the two midpoints) becomes a *trig atom*; `trig_arg(k, z, θ)`
returns the k-th argument and the slices read sin / cos of the atoms through the accessor `tr(2k)`, `tr(2k+1)`.
Atoms 0..NTRIG_VAR-1 depend on z, the rest only on θ (constant over an interior-point solve).  The kernels
fill the table with ONE sincos call site in a rolled loop — inlining fp64 sin/cos at ≈ 400 call sites made
a 73 k-instruction kernel that was bound by instruction fetch."""
from __future__ import annotations

import os
import sys

import sympy as sp
from sympy.printing.c import C99CodePrinter

HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(HERE))
    from modelgen.robots import get_model  # type: ignore
    from modelgen.symbolic import symbolic_residual_env  # type: ignore
    from modelgen.terrain import get_terrain  # type: ignore
else:
    from .robots import get_model
    from .symbolic import symbolic_residual_env
    from .terrain import get_terrain

GEN_DIR = os.path.join(os.path.dirname(HERE), "csrc", "gen")
ROBOTS = {"hopper_2D": "hopper2d", "quadruped": "quadruped", "flamingo": "flamingo",
          "centroidal_quadruped": "centroidal",
          # payload variants: same sizes and kernels, other inertial parameters (examples/quadruped/payload.jl simulates
          # `quadruped_payload` under a policy built on the nominal model; BASELINE config 5 asks for the centroidal analogue)
          "quadruped_payload": "quadruped_payload", "centroidal_quadruped_payload": "centroidal_payload",
          # planar robots on the piecewise terrain (get_simulation(robot, "piecewise1_2D_lc", "piecewise", approx = true),
          # examples/hopper/piecewise.jl:11, examples/flamingo/piecewise.jl:11, examples/quadruped/piecewise.jl:11)
          "hopper_2D_piecewise": "hopper2d_piecewise", "flamingo_piecewise": "flamingo_piecewise",
          "quadruped_piecewise": "quadruped_piecewise"}


class _Printer(C99CodePrinter):
    def _print_Pow(self, e):
        b, p = e.base, e.exp
        if p.is_Integer and 2 <= int(p) <= 4:
            s = self._print(b)
            return "(" + "*".join([f"({s})"] * int(p)) + ")"
        if p == -1:
            return f"(1.0/({self._print(b)}))"
        if p.is_Integer and -4 <= int(p) <= -2:
            s = self._print(b)
            return "(1.0/(" + "*".join([f"({s})"] * (-int(p))) + "))"
        return super()._print_Pow(e)


NS = 8  # slices per generated function = warps per simulator CTA


def _deal(costs, ns=NS):
    """Greedy balanced partition: largest expression first onto the lightest slice."""
    bins = [[] for _ in range(ns)]
    load = [0] * ns
    for k in sorted(range(len(costs)), key=lambda k_: (-costs[k_], k_)):
        b = min(range(ns), key=lambda b_: (load[b_], b_))
        bins[b].append(k)
        load[b] += costs[k]
    return [sorted(b) for b in bins]


def _hoist_trig(exprs, z, n_terr=0):
    """Replace sin(a)/cos(a) by symbols sn<k>/cs<k>; returns (new exprs, [arguments], #z-dependent).
    `n_terr` terrain contacts reserve 2·n_terr atoms (argument 0) right after the z-dependent trig atoms: atom
    T0 + 2i holds (height, slope) under contact i, atom T0 + 2i + 1 (cos, sin) of its surface rotation."""
    args = set()
    for e in exprs:
        for a in e.atoms(sp.sin, sp.cos):
            args.add(a.args[0])
    zs = set(z)
    var = sorted((a for a in args if a.free_symbols & zs), key=sp.default_sort_key)
    const = sorted((a for a in args if not (a.free_symbols & zs)), key=sp.default_sort_key)
    order = var + [sp.Integer(0)] * (2 * n_terr) + const
    sub = {}
    for k, a in enumerate(order):
        if len(var) <= k < len(var) + 2 * n_terr:
            continue
        sub[sp.sin(a)] = sp.Symbol(f"sn{k}", real=True)
        sub[sp.cos(a)] = sp.Symbol(f"cs{k}", real=True)
    out = [e.xreplace(sub) for e in exprs]
    for e in out:
        assert not e.atoms(sp.sin, sp.cos), "unhoisted trigonometric atom"
    return out, order, len(var) + 2 * n_terr


def _emit(exprs, idx, out_fmt, z, th, pr, ntrig=0, terr0=0, n_terr=0):
    """Straight-line code for the outputs `idx` of `exprs` (own CSE)."""
    if not idx:
        return [], 0
    repl, red = sp.cse([exprs[k] for k in idx], symbols=sp.numbered_symbols("x"), optimizations="basic")
    lines = []
    used = set()
    for _, e in repl:
        used |= e.free_symbols
    for e in red:
        used |= e.free_symbols
    for i, s in enumerate(z):
        if s in used:
            lines.append(f"  const double z{i} = z({i});")
    for i, s in enumerate(th):
        if s in used:
            lines.append(f"  const double t{i} = th({i});")
    names = {str(s) for s in used}
    for k in range(ntrig):
        if f"sn{k}" in names:
            lines.append(f"  const double sn{k} = tr({2 * k});")
        if f"cs{k}" in names:
            lines.append(f"  const double cs{k} = tr({2 * k + 1});")
    for i in range(n_terr):
        for nm, slot in ((f"tf{i}", 2 * (terr0 + 2 * i)), (f"tg{i}", 2 * (terr0 + 2 * i) + 1),
                         (f"tc{i}", 2 * (terr0 + 2 * i + 1)), (f"ts{i}", 2 * (terr0 + 2 * i + 1) + 1)):
            if nm in names:
                lines.append(f"  const double {nm} = tr({slot});")
    for s, e in repl:
        lines.append(f"  const double {s} = {pr.doprint(e)};")
    for k, e in zip(idx, red):
        lines.append("  " + out_fmt.format(i=k, e=pr.doprint(e)))
    return lines, len(repl)


def generate(robot: str) -> str:
    m = get_model(robot)
    z, th, kappa, r, r_diff, terr = symbolic_residual_env(m)
    n_terr = m.nc if terr else 0
    pr = _Printer()
    rvec = sp.Matrix(r_diff)  # = r on flat ground; on a terrain the surface enters as its tangent line (symbolic.py)
    J = rvec.jacobian(sp.Matrix(z))
    nnz = [(i, j) for j in range(m.nz) for i in range(m.nz) if J[i, j] != 0]  # column-major order
    Jt = rvec.jacobian(sp.Matrix(th))
    nnzt = [(i, j) for j in range(m.ntheta) for i in range(m.nz) if Jt[i, j] != 0]
    px = list(terr["px"]) if terr else []
    hoisted, trig_args, ntv = _hoist_trig(list(r) + [J[i, j] for i, j in nnz] + [Jt[i, j] for i, j in nnzt] + px, z, n_terr)
    nt = len(trig_args)
    terr0 = ntv - 2 * n_terr
    if terr:
        px_h = hoisted[len(hoisted) - n_terr:]
        hoisted = hoisted[:len(hoisted) - n_terr]
        for e in hoisted:
            assert not (e.free_symbols & set(terr["x0"])), "expansion point left in an output"
    r_exprs = hoisted[:m.nz]
    j_exprs = hoisted[m.nz:m.nz + len(nnz)]
    t_exprs = hoisted[m.nz + len(nnz):]
    r_bins = _deal([sp.count_ops(e) + 1 for e in r_exprs])
    j_bins = _deal([sp.count_ops(e) + 1 for e in j_exprs])
    r_sl = [_emit(r_exprs, b, "r({i}, {e});", z, th, pr, nt, terr0, n_terr) for b in r_bins]
    j_sl = [_emit(j_exprs, b, "J({i}, {e});", z, th, pr, nt, terr0, n_terr) for b in j_bins]
    t_bins = _deal([sp.count_ops(e) + 1 for e in t_exprs])
    t_sl = [_emit(t_exprs, b, "J({i}, {e});", z, th, pr, nt, terr0, n_terr) for b in t_bins]
    tag = ROBOTS[robot]
    out = []
    out.append(f"// GENERATED by contactimplicitmpc.jl_b200/modelgen/codegen.py — do not edit.")
    out.append(f"// Nonlinear contact residual of `{robot}` (src/simulation/simulation.jl:133-158 with")
    out.append(f"// src/dynamics/{robot}/model.jl and src/dynamics/model.jl:18-41), CSE'd straight-line code in")
    out.append(f"// {NS} independent slices.  r temporaries per slice: {[t for _, t in r_sl]};")
    out.append(f"// rz: {len(nnz)} structural non-zeros of {m.nz}x{m.nz}, temporaries per slice: {[t for _, t in j_sl]};")
    out.append(f"// rθ: {len(nnzt)} structural non-zeros of {m.nz}x{m.ntheta}, temporaries per slice: {[t for _, t in t_sl]}.")
    out.append("#pragma once")
    out.append("#include <math.h>")
    out.append("#ifdef __CUDACC__\n#define CIMPC_GEN_HD __host__ __device__ __forceinline__\n#else\n#define CIMPC_GEN_HD inline\n#endif")
    out.append(f"namespace cimpc {{ namespace gen_{tag} {{")
    out.append(f"constexpr int NQ = {m.nq}, NU = {m.nu}, NW = {m.nw}, NC = {m.nc}, NB = {m.nb};")
    out.append(f"constexpr int NZ = {m.nz}, NTH = {m.ntheta}, NNZ = {len(nnz)}, NNZT = {len(nnzt)}, NS = {NS};")
    out.append(f"// trig atoms: sin/cos arguments; the first NTRIG_VAR depend on z, the others only on θ")
    out.append(f"constexpr int NTRIG = {nt}, NTRIG_VAR = {ntv};")
    out.append(f"// terrain atoms (table slots of NTRIG_VAR): contact i has (height, slope) in atom TERR0 + 2i and (cos, sin) of the")
    out.append(f"// surface rotation in atom TERR0 + 2i + 1; NTERR = 0 on flat ground")
    out.append(f"constexpr int NTERR = {n_terr}, TERR0 = {terr0};")
    zsub = {s_: sp.Symbol(f"z({i})") for i, s_ in enumerate(z)}
    zsub.update({s_: sp.Symbol(f"th({i})") for i, s_ in enumerate(th)})
    out.append("template <class ZA, class TA>\nCIMPC_GEN_HD double trig_arg(int k, ZA z, TA th) {")
    out.append("  switch (k) {")
    for k, a in enumerate(trig_args):
        out.append(f"    case {k}: return {pr.doprint(a.xreplace(zsub))};")
    out.append("    default: return 0.0;\n  }\n}")
    if terr:
        out.append(get_terrain(terr["name"]).c_source())
        out.append("// x coordinate of contact i (the point the surface is evaluated at), from z and the trig table")
        out.append("template <class ZA, class TA, class TR>\nCIMPC_GEN_HD double terr_x(int i, ZA z, TA th, TR tr) {")
        out.append("  switch (i) {")
        tsub = dict(zsub)
        for k in range(nt):
            tsub[sp.Symbol(f"sn{k}", real=True)] = sp.Symbol(f"tr({2 * k})")
            tsub[sp.Symbol(f"cs{k}", real=True)] = sp.Symbol(f"tr({2 * k + 1})")
        for i, e in enumerate(px_h):
            out.append(f"    case {i}: return {pr.doprint(e.xreplace(tsub))};")
        out.append("    default: return 0.0;\n  }\n}")
    else:
        out.append("CIMPC_GEN_HD void surf(const double, double& s, double& ds) { s = 0.0; ds = 0.0; }")
        out.append("template <class ZA, class TA, class TR>\nCIMPC_GEN_HD double terr_x(int, ZA, TA, TR) { return 0.0; }")
    out.append("// the two atoms of one contact from its x coordinate: (height, slope), (cos, sin) of the surface rotation")
    out.append("// (src/simulator/environment.jl:81-96: the world→surface angle is −atan(slope))")
    out.append("CIMPC_GEN_HD void terr_atoms(const double x, double& f, double& g, double& c, double& s) {")
    out.append("  surf(x, f, g);\n  c = 1.0 / sqrt(1.0 + g * g);\n  s = g * c;\n}")
    out.append("// fills tab[2k] = sin(arg_k), tab[2k+1] = cos(arg_k) for k0 <= k < NTRIG, then the terrain atoms (host-side convenience)")
    out.append("template <class ZA, class TA>\nCIMPC_GEN_HD void eval_trig(ZA z, TA th, double* tab, int k0 = 0) {")
    out.append("  for (int k = k0; k < NTRIG; ++k) { const double a = trig_arg(k, z, th); tab[2 * k] = sin(a); tab[2 * k + 1] = cos(a); }")
    out.append("  for (int i = 0; i < NTERR; ++i) {")
    out.append("    const double x = terr_x(i, z, th, [&](int j) { return tab[j]; });")
    out.append("    terr_atoms(x, tab[2 * (TERR0 + 2 * i)], tab[2 * (TERR0 + 2 * i) + 1], tab[2 * (TERR0 + 2 * i + 1)], tab[2 * (TERR0 + 2 * i + 1) + 1]);")
    out.append("  }")
    out.append("}")
    out.append("// structural non-zeros of rz (0-based row / column), column-major order")
    out.append("constexpr short RZ_ROW[NNZ] = {" + ", ".join(str(i) for i, _ in nnz) + "};")
    out.append("constexpr short RZ_COL[NNZ] = {" + ", ".join(str(j) for _, j in nnz) + "};")
    out.append("// structural non-zeros of rθ (0-based row / column), column-major order")
    out.append("constexpr short RTH_ROW[NNZT] = {" + ", ".join(str(i) for i, _ in nnzt) + "};")
    out.append("constexpr short RTH_COL[NNZT] = {" + ", ".join(str(j) for _, j in nnzt) + "};")
    out.append("// z(i), th(i): input accessors; r(i, value): output sink")
    for s_, (lines, _) in enumerate(r_sl):
        out.append(f"template <class ZA, class TA, class TR, class RA>\nCIMPC_GEN_HD void eval_r_{s_}(ZA z, TA th, TR tr, const double kappa, RA r) {{")
        out += lines
        out.append("}")
    out.append("template <class ZA, class TA, class RA>\nCIMPC_GEN_HD void eval_r(ZA z, TA th, const double kappa, RA r) {")
    out.append("  double tab[2 * NTRIG + 2];\n  eval_trig(z, th, tab);\n  auto tr = [&](int i) { return tab[i]; };")
    out += [f"  eval_r_{s_}(z, th, tr, kappa, r);" for s_ in range(NS)]
    out.append("}")
    out.append("// J(k, value): k-th structural non-zero (RZ_ROW[k], RZ_COL[k])")
    for s_, (lines, _) in enumerate(j_sl):
        out.append(f"template <class ZA, class TA, class TR, class JA>\nCIMPC_GEN_HD void eval_rz_{s_}(ZA z, TA th, TR tr, JA J) {{")
        out += lines
        out.append("}")
    out.append("template <class ZA, class TA, class JA>\nCIMPC_GEN_HD void eval_rz(ZA z, TA th, JA J) {")
    out.append("  double tab[2 * NTRIG + 2];\n  eval_trig(z, th, tab);\n  auto tr = [&](int i) { return tab[i]; };")
    out += [f"  eval_rz_{s_}(z, th, tr, J);" for s_ in range(NS)]
    out.append("}")
    out.append("// J(k, value): k-th structural non-zero of rθ (RTH_ROW[k], RTH_COL[k])")
    for s_, (lines, _) in enumerate(t_sl):
        out.append(f"template <class ZA, class TA, class TR, class JA>\nCIMPC_GEN_HD void eval_rth_{s_}(ZA z, TA th, TR tr, JA J) {{")
        out += lines
        out.append("}")
    out.append("template <class ZA, class TA, class JA>\nCIMPC_GEN_HD void eval_rth(ZA z, TA th, JA J) {")
    out.append("  double tab[2 * NTRIG + 2];\n  eval_trig(z, th, tab);\n  auto tr = [&](int i) { return tab[i]; };")
    out += [f"  eval_rth_{s_}(z, th, tr, J);" for s_ in range(NS)]
    out.append("}")
    out.append("#ifdef __CUDACC__")
    out.append("static __device__ const short RTH_ROW_D[NNZT] = {" + ", ".join(str(i) for i, _ in nnzt) + "};")
    out.append("static __device__ const short RTH_COL_D[NNZT] = {" + ", ".join(str(j) for _, j in nnzt) + "};")
    out.append("static __device__ const short RZ_ROW_D[NNZ] = {" + ", ".join(str(i) for i, _ in nnz) + "};")
    out.append("static __device__ const short RZ_COL_D[NNZ] = {" + ", ".join(str(j) for _, j in nnz) + "};")
    out.append("#endif")
    out.append("// traits bundle consumed by the simulator kernels (csrc/sim_kernel.cuh)")
    out.append("struct Gen {")
    out.append("  static constexpr int NQ = gen_%s::NQ, NU = gen_%s::NU, NW = gen_%s::NW, NC = gen_%s::NC, NB = gen_%s::NB;" % ((tag,) * 5))
    out.append("  static constexpr int NZ = gen_%s::NZ, NTH = gen_%s::NTH, NNZ = gen_%s::NNZ, NNZT = gen_%s::NNZT, NS = gen_%s::NS;" % ((tag,) * 5))
    out.append("  static constexpr int NTRIG = gen_%s::NTRIG, NTRIG_VAR = gen_%s::NTRIG_VAR;" % ((tag,) * 2))
    out.append("  static constexpr int NTERR = gen_%s::NTERR, TERR0 = gen_%s::TERR0;" % ((tag,) * 2))
    out.append("  static constexpr CIMPC_GEN_HD bool is_terr(int k) { return NTERR > 0 && k >= TERR0 && k < TERR0 + 2 * NTERR; }")
    out.append("  template <class ZA, class TA, class TR> static CIMPC_GEN_HD double terr_x(int i, ZA z, TA th, TR tr) { return gen_%s::terr_x(i, z, th, tr); }" % tag)
    out.append("  static CIMPC_GEN_HD void terr_atoms(double x, double& f, double& g, double& c, double& s) { gen_%s::terr_atoms(x, f, g, c, s); }" % tag)
    out.append("  template <class ZA, class TA> static CIMPC_GEN_HD double trig_arg(int k, ZA z, TA th) { return gen_%s::trig_arg(k, z, th); }" % tag)
    out.append("  template <class ZA, class TA, class RA> static CIMPC_GEN_HD void r(ZA z, TA th, double kappa, RA out) { eval_r(z, th, kappa, out); }")
    out.append("  template <class ZA, class TA, class JA> static CIMPC_GEN_HD void rz(ZA z, TA th, JA out) { eval_rz(z, th, out); }")
    out.append("  // slice s of the outputs (s warp-uniform in the kernels)")
    out.append("  template <class ZA, class TA, class TR, class RA> static CIMPC_GEN_HD void r_slice(int s, ZA z, TA th, TR tr, double kappa, RA out) {")
    out.append("    switch (s) {")
    out += [f"      case {s_}: eval_r_{s_}(z, th, tr, kappa, out); break;" for s_ in range(NS)]
    out.append("      default: break;\n    }\n  }")
    out.append("  template <class ZA, class TA, class TR, class JA> static CIMPC_GEN_HD void rz_slice(int s, ZA z, TA th, TR tr, JA out) {")
    out.append("    switch (s) {")
    out += [f"      case {s_}: eval_rz_{s_}(z, th, tr, out); break;" for s_ in range(NS)]
    out.append("      default: break;\n    }\n  }")
    out.append("  template <class ZA, class TA, class TR, class JA> static CIMPC_GEN_HD void rth_slice(int s, ZA z, TA th, TR tr, JA out) {")
    out.append("    switch (s) {")
    out += [f"      case {s_}: eval_rth_{s_}(z, th, tr, out); break;" for s_ in range(NS)]
    out.append("      default: break;\n    }\n  }")
    out.append("#ifdef __CUDACC__")
    out.append("  static __device__ __forceinline__ int trow(int k) { return RTH_ROW_D[k]; }")
    out.append("  static __device__ __forceinline__ int tcol(int k) { return RTH_COL_D[k]; }")
    out.append("  static __device__ __forceinline__ int row(int k) { return RZ_ROW_D[k]; }")
    out.append("  static __device__ __forceinline__ int col(int k) { return RZ_COL_D[k]; }")
    out.append("#endif")
    out.append("};")
    out.append("}}  // namespace")
    return "\n".join(out) + "\n"


def main(robots=None):
    os.makedirs(GEN_DIR, exist_ok=True)
    for robot in (robots or list(ROBOTS)):
        txt = generate(robot)
        path = os.path.join(GEN_DIR, f"residual_{ROBOTS[robot]}.h")
        with open(path, "w") as f:
            f.write(txt)
        print(path, len(txt), "bytes")


if __name__ == "__main__":
    main(sys.argv[1:])
