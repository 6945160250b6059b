"""Shared helpers for the tests, `__graft_entry__.smoke()` and `bench.py`: committed fixtures,
seeded problem batches, and the oracle evaluated from those fixtures (no sympy, no /root/reference)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# robot -> (nq, nu, nw, nc, nb)   (SURVEY.md §2 dimension table)
SIZES = {
    "hopper_2D": (4, 2, 2, 1, 2),
    "quadruped": (11, 8, 2, 4, 8),
    "flamingo": (9, 6, 2, 4, 8),
    "centroidal_quadruped": (18, 12, 3, 4, 16),
}


def load_lin(robot: str) -> dict:
    with np.load(os.path.join(GOLDEN, f"{robot}_lin.npz")) as f:
        return {k: f[k] for k in f.files}


def load_gait(robot: str) -> dict:
    with np.load(os.path.join(GOLDEN, f"{robot}_gait.npz")) as f:
        return {k: (f[k] if f[k].ndim else float(f[k])) for k in f.files}


class FixtureIndex:
    """index.jl layout from sizes only (so the oracle IP can run without building the sympy model)."""

    def __init__(self, nq, nu, nw, nc, nb):
        o = 0
        self.q2 = np.arange(o, o + nq); o += nq
        self.g1 = np.arange(o, o + nc); o += nc
        self.b1 = np.arange(o, o + nb); o += nb
        self.psi1 = np.arange(o, o + nc); o += nc
        self.s1 = np.arange(o, o + nc); o += nc
        self.eta1 = np.arange(o, o + nb); o += nb
        self.s2 = np.arange(o, o + nc); o += nc
        self.nz = o
        self.ntheta = 2 * nq + nu + nw + 2
        self.q0 = np.arange(0, nq)
        self.q1 = np.arange(nq, 2 * nq)
        self.u1 = np.arange(2 * nq, 2 * nq + nu)
        self.dyn, self.imp, self.mdp, self.fri = self.q2, self.g1, self.b1, self.psi1
        self.bimp, self.bmdp, self.bfri = self.s1, self.eta1, self.s2
        self.x = self.q2
        self.y1 = np.concatenate([self.g1, self.b1, self.psi1])
        self.y2 = np.concatenate([self.s1, self.eta1, self.s2])
        self.rst = np.concatenate([self.imp, self.mdp, self.fri])
        self.bil = np.concatenate([self.bimp, self.bmdp, self.bfri])
        self.alt = self.imp


def oracle_problems(robot: str, lin: dict, solver: str = "mgs"):
    """One oracle `LinProblem` per knot, built from the committed linearization fixture.
    solver = "mgs" (the reference's QR) or "lu" (accurate variant, see oracle/linearized.py)."""
    from oracle.linearized import LinProblem, lin_blocks
    nq, nu, nw, nc, nb = SIZES[robot]
    idx = FixtureIndex(nq, nu, nw, nc, nb)
    probs = []
    for t in range(lin["z0"].shape[0]):
        blk = lin_blocks(idx, nc, lin["z0"][t], lin["th0"][t], lin["r0"][t], lin["rz0"][t], lin["rth0"][t])
        probs.append(LinProblem(blk, idx, solver=solver))
    return idx, probs


def make_batch(robot: str, lin: dict, gait: dict, n: int, seed: int = 100, sigma_q: float = 0.02,
               sigma_u: float = 0.1):
    """Seeded (knot, θ, q2_init) batch: reference-gait θ_t plus N(0, σ²) on q0, q1 and
    0.1·|u|·N(0,1) on u1; cold-start q2 = reference q_{t+2} plus the same σ.
    Problems are ordered stage-major: knot = i mod H_ref."""
    nq, nu, nw, nc, nb = SIZES[robot]
    rng = np.random.Generator(np.random.Philox(seed))
    H = lin["z0"].shape[0]
    knot = (np.arange(n) % H).astype(np.int32)
    theta = lin["th0"][knot].copy()
    theta[:, :2 * nq] += sigma_q * rng.standard_normal((n, 2 * nq))
    theta[:, 2 * nq:2 * nq + nu] += sigma_u * np.abs(theta[:, 2 * nq:2 * nq + nu]) * rng.standard_normal((n, nu))
    q2 = lin["z0"][knot, :nq] + sigma_q * rng.standard_normal((n, nq))
    return knot, theta, q2


def oracle_solve_batch(robot, lin, knot, theta, q2, opts, alt=None, mode="configuration", solver="mgs"):
    """Run oracle/ip.py on every problem of the batch (slow: small n only)."""
    from oracle.ip import interior_point_solve
    nq, nu, nw, nc, nb = SIZES[robot]
    idx, probs = oracle_problems(robot, lin, solver=solver)
    n = len(knot)
    nd = nq if mode == "configuration" else nq + nc + nb
    ncol = 2 * nq + nu
    z = np.zeros((n, idx.nz))
    dz = np.zeros((n, nd, ncol))
    status = np.zeros(n, dtype=bool)
    iters = np.zeros(n, dtype=np.int32)
    for i in range(n):
        p = probs[knot[i]]
        p.b.alt = np.zeros(nc) if alt is None else np.asarray(alt[i], dtype=np.float64)
        z0 = np.ones(idx.nz)
        z0[idx.q2] = q2[i]
        st, zs, dzs, it = interior_point_solve(p, z0, theta[i], opts)
        z[i], status[i], iters[i] = zs, st, it
        if dzs is not None:
            dz[i] = dzs[:nd, :ncol]
    return z, dz, status, iters


def independent_violation(robot, lin, knot, theta, z, alt=None):
    """r_vio, κ_vio of the linearized residual recomputed with plain numpy from the raw fixture
    (r0 + rz0 (z − z0) + rθ0 (θ − θ0) on the linear rows, y1∘y2 on the bilinear rows)."""
    nq, nu, nw, nc, nb = SIZES[robot]
    idx = FixtureIndex(nq, nu, nw, nc, nb)
    lin_rows = np.concatenate([idx.dyn, idx.rst])
    r_vio = np.zeros(len(knot))
    k_vio = np.zeros(len(knot))
    for i, t in enumerate(knot):
        r = lin["r0"][t] + lin["rz0"][t] @ (z[i] - lin["z0"][t]) + lin["rth0"][t] @ (theta[i] - lin["th0"][t])
        if alt is not None:
            r[idx.alt] += alt[i]
        r_vio[i] = np.abs(r[lin_rows]).max()
        k_vio[i] = np.abs(z[i][idx.y1] * z[i][idx.y2]).max()
    return r_vio, k_vio


# ---- the PRODUCT's generated residual code evaluated on the host (no oracle involved) ---------------------------
GEN_DIR = os.path.join(ROOT, "contactimplicitmpc.jl_b200", "csrc", "gen")
GEN_TAG = {"hopper_2D": "hopper2d", "quadruped": "quadruped", "flamingo": "flamingo", "centroidal_quadruped": "centroidal",
           "hopper_2D_piecewise": "hopper2d_piecewise", "flamingo_piecewise": "flamingo_piecewise", "quadruped_piecewise": "quadruped_piecewise",
           "quadruped_payload": "quadruped_payload", "centroidal_quadruped_payload": "centroidal_payload"}
_GEN_SHIM = r'''
#include "residual_%(tag)s.h"
using namespace cimpc::gen_%(tag)s;
extern "C" {
int nnz() { return NNZ; }
void pattern(int* row, int* col) { for (int k = 0; k < NNZ; ++k) { row[k] = RZ_ROW[k]; col[k] = RZ_COL[k]; } }
void r_eval(const double* z, const double* th, double kappa, double* r) {
  eval_r([&](int i) { return z[i]; }, [&](int i) { return th[i]; }, kappa, [&](int i, double v) { r[i] = v; });
}
void rz_eval(const double* z, const double* th, double* J) {
  eval_rz([&](int i) { return z[i]; }, [&](int i) { return th[i]; }, [&](int k, double v) { J[k] = v; });
}
int nnzt() { return NNZT; }
void pattern_t(int* row, int* col) { for (int k = 0; k < NNZT; ++k) { row[k] = RTH_ROW[k]; col[k] = RTH_COL[k]; } }
void rth_eval(const double* z, const double* th, double* J) {
  eval_rth([&](int i) { return z[i]; }, [&](int i) { return th[i]; }, [&](int k, double v) { J[k] = v; });
}
}
'''


class GeneratedResidual:
    """`csrc/gen/residual_<tag>.h` (what the CUDA kernels compile) built for the host with g++ and called via ctypes."""

    def __init__(self, robot: str, workdir: str):
        import ctypes as C
        import subprocess
        tag = GEN_TAG[robot]
        nq, nu, nw, nc, nb = SIZES[robot.replace("_payload", "").replace("_piecewise", "")]
        self.nz, self.nth = nq + 4 * nc + 2 * nb, 2 * nq + nu + nw + 2
        src = os.path.join(workdir, f"shim_{tag}.cpp")
        so = os.path.join(workdir, f"shim_{tag}.so")
        with open(src, "w") as f:
            f.write(_GEN_SHIM % {"tag": tag})
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{GEN_DIR}", src, "-o", so], check=True)
        self.lib = lib = C.CDLL(so)
        lib.r_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        self.nnz, self.nnzt = lib.nnz(), lib.nnzt()
        self.row = np.zeros(self.nnz, np.int32); self.col = np.zeros(self.nnz, np.int32)
        lib.pattern(self.row.ctypes.data_as(C.c_void_p), self.col.ctypes.data_as(C.c_void_p))
        self.trow = np.zeros(self.nnzt, np.int32); self.tcol = np.zeros(self.nnzt, np.int32)
        lib.pattern_t(self.trow.ctypes.data_as(C.c_void_p), self.tcol.ctypes.data_as(C.c_void_p))

    def r(self, z, th, kappa):
        z, th = np.ascontiguousarray(z, dtype=np.float64), np.ascontiguousarray(th, dtype=np.float64)
        out = np.zeros(self.nz)
        self.lib.r_eval(z.ctypes.data, th.ctypes.data, float(kappa), out.ctypes.data)
        return out

    def rz(self, z, th):
        z, th = np.ascontiguousarray(z, dtype=np.float64), np.ascontiguousarray(th, dtype=np.float64)
        J = np.zeros(self.nnz)
        self.lib.rz_eval(z.ctypes.data, th.ctypes.data, J.ctypes.data)
        d = np.zeros((self.nz, self.nz)); d[self.row, self.col] = J
        return d

    def rth(self, z, th):
        z, th = np.ascontiguousarray(z, dtype=np.float64), np.ascontiguousarray(th, dtype=np.float64)
        J = np.zeros(self.nnzt)
        self.lib.rth_eval(z.ctypes.data, th.ctypes.data, J.ctypes.data)
        d = np.zeros((self.nz, self.nth)); d[self.trow, self.tcol] = J
        return d
