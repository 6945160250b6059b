"""Linearized contact subproblem: blocks, residual, Schur + MGS-QR solves (CPU oracle, numpy fp64).

TEST INFRASTRUCTURE ONLY.  Restates, function by function,
  * `RLin` / `RZLin` / `RθLin` block slicing     src/controller/linearized_solver.jl:67-161, 224-304, 325-359
  * `rlin!`                                      src/controller/linearized_solver.jl:364-373
  * `rzlin!`                                     src/controller/linearized_solver.jl:378-399
  * `linear_solve!(Δ, rz, r)`                    src/controller/linearized_solver.jl:424-444
  * `linear_solve!(δz, rz, rθ)`                  src/controller/linearized_solver.jl:451-479
  * `residual_violation` / `bilinear_violation`  src/controller/linearized_solver.jl:401-409
  * `general_correction_term!`                   src/controller/linearized_solver.jl:411-418
  * `Schur`, `schur_factorize!`, `schur_solve!`  src/solver/schur.jl:33-49, 80-88, 93-110
  * `SDMGSSolver` `factorize!`, `qr_solve!`      src/solver/qr.jl:113-137, 142-158
  * `LinearizedStep`                             src/controller/linearized_step.jl:10-29

Plain loops on purpose: the goal is to be readable against the Julia, not fast
(the timed CPU baseline is the C restatement in `oracle/c/`).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .residual import Index, Residual


# ------------------------------------------------------------------ MGS QR (qr.jl)
class MGS:
    """Static dense modified Gram-Schmidt (qr.jl:65-158).  qs[j] = j-th orthonormal column,
    rs = packed upper triangle, column-major: rs[triu_perm(k, j)] = R[k, j]  (qr.jl:18-22)."""

    def __init__(self, n: int):
        self.n = n
        self.qs = np.zeros((n, n))  # qs[j] is row j here == column vector q_j
        self.rs = np.zeros(n * (n + 1) // 2)

    @staticmethod
    def triu_perm(k: int, j: int) -> int:  # 1-based (k, j) → 0-based offset
        return (j - 1) * j // 2 + k - 1

    def factorize(self, A: np.ndarray):  # qr.jl:113-137
        n = self.n
        off = 0
        for j in range(n):
            q = A[:, j].copy()
            for k in range(j):
                rkj = float(np.dot(q, self.qs[k]))
                self.rs[off] = rkj
                q = q - self.qs[k] * rkj
                off += 1
            nrm = float(np.sqrt(np.dot(q, q)))
            self.rs[off] = nrm
            self.qs[j] = q / nrm
            off += 1

    def solve(self, b: np.ndarray) -> np.ndarray:  # qr.jl:142-158
        n = self.n
        x = np.array([float(np.dot(self.qs[j], b)) for j in range(n)])
        for j in range(n, 0, -1):
            for k in range(j + 1, n + 1):
                x[j - 1] -= self.rs[self.triu_perm(j, k)] * x[k - 1]
            x[j - 1] /= self.rs[self.triu_perm(j, j)]
        return x


class DenseLU:
    """NOT in the reference: LAPACK LU (partial pivoting) with the MGS interface.  Used by the tests as
    the *accurate* variant of the oracle: the reference's MGS solve R⁻¹Qᵀb loses ≈ cond(S)²·eps, which on
    ill-conditioned knots is far above fp64 round-off, so two faithful restatements of the reference
    that differ only in summation order already disagree there (see tests/test_oracle.py)."""

    def __init__(self, n: int):
        self.n = n

    def factorize(self, A: np.ndarray):
        import scipy.linalg
        self.lu = scipy.linalg.lu_factor(A)

    def solve(self, b: np.ndarray) -> np.ndarray:
        import scipy.linalg
        return scipy.linalg.lu_solve(self.lu, b)


# ------------------------------------------------------------------ Schur (schur.jl)
class Schur:
    """[A B; C D] with constant A, B, C: caches A⁻¹, C A⁻¹, C A⁻¹ B (schur.jl:33-49)."""

    def __init__(self, A, B, C, D, solver: str = "mgs"):
        self.A, self.B, self.C = A, B, C
        self.Ai = np.linalg.inv(A)           # schur.jl:39  `inv(A)`
        self.CAi = C @ self.Ai
        self.CAiB = C @ self.Ai @ B
        self.gs = MGS(D.shape[0]) if solver == "mgs" else DenseLU(D.shape[0])
        self.factorize(D)

    def factorize(self, D):  # schur_factorize!  schur.jl:80-88
        self.gs.factorize(D - self.CAiB)

    def solve(self, u, v):  # schur_solve!  schur.jl:93-110
        temp = self.gs.solve(self.CAi @ u - v)
        x = self.Ai @ (u + self.B @ temp)
        y = -temp
        return x, y


# ------------------------------------------------------------------ linearization data
@dataclass
class LinBlocks:
    """Everything `RLin` + `RZLin` + `RθLin` hold for one reference knot (a2 in SURVEY §8)."""
    nx: int
    ny: int
    nth: int
    nc: int
    Dx: np.ndarray
    Dy1: np.ndarray
    Rx: np.ndarray
    Ry1: np.ndarray
    Ry2: np.ndarray        # diag(rz0[rst, y2])
    rth_dyn: np.ndarray
    rth_rst: np.ndarray
    rth_bil: np.ndarray
    rdyn0: np.ndarray
    rrst0: np.ndarray
    rbil0: np.ndarray
    x0: np.ndarray
    y10: np.ndarray
    y20: np.ndarray
    th0: np.ndarray
    alt: np.ndarray = field(default=None)

    def __post_init__(self):
        if self.alt is None:
            self.alt = np.zeros(self.nc)


def linearized_step(res: Residual, z0, th0, kappa):
    """LinearizedStep(s, z, θ, κ)  linearized_step.jl:10-29 → (r0, rz0, rθ0)."""
    z0 = np.asarray(z0, dtype=np.float64)
    th0 = np.asarray(th0, dtype=np.float64)
    return res.r(z0, th0, kappa), res.rz(z0, th0), res.rth(z0, th0)


def lin_blocks(idx: Index, nc: int, z0, th0, r0, rz0, rth0) -> LinBlocks:
    """RLin / RZLin / RθLin constructors: block slicing (linearized_solver.jl:92-120, 245-251, 346-348)."""
    ix, iy1, iy2 = idx.x, idx.y1, idx.y2
    idyn, irst, ibil = idx.dyn, idx.rst, idx.bil
    return LinBlocks(
        nx=len(ix), ny=len(iy1), nth=idx.ntheta, nc=nc,
        Dx=rz0[np.ix_(idyn, ix)].copy(), Dy1=rz0[np.ix_(idyn, iy1)].copy(),
        Rx=rz0[np.ix_(irst, ix)].copy(), Ry1=rz0[np.ix_(irst, iy1)].copy(),
        Ry2=np.diag(rz0[np.ix_(irst, iy2)]).copy(),
        rth_dyn=rth0[idyn, :].copy(), rth_rst=rth0[irst, :].copy(), rth_bil=rth0[ibil, :].copy(),
        rdyn0=r0[idyn].copy(), rrst0=r0[irst].copy(), rbil0=r0[ibil].copy(),
        x0=np.asarray(z0)[ix].copy(), y10=np.asarray(z0)[iy1].copy(), y20=np.asarray(z0)[iy2].copy(),
        th0=np.asarray(th0).copy())


class LinProblem:
    """One `InteriorPoint` object of `ImplicitTrajectory.ip[t]` with r = RLin, rz = RZLin, rθ = RθLin
    (implicit_dynamics.jl:58-68): the hooks the IP loop calls."""

    def __init__(self, blk: LinBlocks, idx: Index, solver: str = "mgs"):
        self.b = blk
        self.idx = idx
        self.nz = idx.nz
        # RZLin ctor (linearized_solver.jl:253-258) builds the Schur object once; its initial
        # factorization (at the reference y1, y2) is always overwritten by rzlin! before any
        # solve, so the constant parts (A⁻¹, C A⁻¹, C A⁻¹ B) are all that matter here.
        self.S = Schur(blk.Dx, blk.Dy1, blk.Rx, blk.Ry1 - np.eye(blk.ny), solver=solver)
        self.y1 = blk.y10.copy()   # diag(rz0[bil, y2]) == y1 at z0  (:250)
        self.y2 = blk.y20.copy()   # diag(rz0[bil, y1]) == y2 at z0  (:251)
        self.rdyn = np.zeros(blk.nx)
        self.rrst = np.zeros(blk.ny)
        self.rbil = np.zeros(blk.ny)

    # rlin!  linearized_solver.jl:364-373
    def r(self, z, th, kappa):
        b, i = self.b, self.idx
        x, y1, y2 = z[i.x], z[i.y1], z[i.y2]
        alt_full = np.concatenate([b.alt, np.zeros(b.ny - b.nc)])
        self.rdyn = b.rdyn0 + b.Dx @ (x - b.x0) + b.Dy1 @ (y1 - b.y10) + b.rth_dyn @ (th - b.th0)
        self.rrst = (b.rrst0 + b.Rx @ (x - b.x0) + b.Ry1 @ (y1 - b.y10) + b.Ry2 * (y2 - b.y20)
                     + b.rth_rst @ (th - b.th0) + alt_full)
        self.rbil = y1 * y2 - kappa

    # rzlin!  linearized_solver.jl:378-399
    def rz(self, z, reg=0.0):
        b, i = self.b, self.idx
        self.y1 = z[i.y1].copy()
        self.y2 = z[i.y2].copy()
        y1r = np.maximum(self.y1, reg)
        y2r = np.maximum(self.y2, reg)
        D = b.Ry1 - np.diag(b.Ry2 * y2r / y1r)
        self.S.factorize(D)

    # residual_violation / bilinear_violation  :401-409
    def r_vio(self):
        return max(np.abs(self.rdyn).max(), np.abs(self.rrst).max())

    def k_vio(self):
        return np.abs(self.rbil).max()

    # general_correction_term!  :411-418
    def correction(self, Delta):
        i = self.idx
        self.rbil = self.rbil + Delta[i.y1] * Delta[i.y2]

    # linear_solve!(Δ, rz, r; reg)  :424-444
    def solve(self, reg=0.0):
        b, i = self.b, self.idx
        y1r = np.maximum(reg, self.y1)
        y2r = np.maximum(reg, self.y2)
        u = self.rdyn
        v = self.rrst - b.Ry2 * self.rbil / y1r
        dx, dy1 = self.S.solve(u, v)
        Delta = np.zeros(self.nz)
        Delta[i.x] = dx
        Delta[i.y1] = dy1
        Delta[i.y2] = (self.rbil - y2r * dy1) / y1r
        return Delta

    # linear_solve!(δz, rz, rθ; reg)  :451-479
    def solve_sensitivity(self, reg=0.0):
        b, i = self.b, self.idx
        y1r = np.maximum(reg, self.y1)
        y2r = np.maximum(reg, self.y2)
        dz = np.zeros((self.nz, b.nth))
        for c in range(b.nth):
            u = b.rth_dyn[:, c]
            v = b.rth_rst[:, c]
            dx, dy1 = self.S.solve(u, v)
            dz[i.x, c] = dx
            dz[i.y1, c] = dy1
            dz[i.y2, c] = (b.rth_bil[:, c] - y2r * dy1) / y1r
        return dz
