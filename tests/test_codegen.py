"""The product's generated residual code (`csrc/gen/residual_<robot>.h`, sympy → C++/CUDA by
`modelgen/codegen.py`) against the oracle's independent restatement: same r, rz and rθ at random points.
(The reference's analogue: test/dynamics/quadruped.jl compares code-generated against hand-written functions.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, SIZES

GEN = os.path.join(ROOT, "contactimplicitmpc.jl_b200", "csrc", "gen")
SHIM = r'''
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


@pytest.mark.parametrize("robot,tag", [("hopper_2D", "hopper2d"), ("quadruped", "quadruped"), ("flamingo", "flamingo"),
                                       ("centroidal_quadruped", "centroidal"), ("quadruped_payload", "quadruped_payload"),
                                       ("centroidal_quadruped_payload", "centroidal_payload")])
def test_generated_residual_matches_oracle(tmp_path, robot, tag):
    from oracle.residual import get_residual
    hdr = os.path.join(GEN, f"residual_{tag}.h")
    assert os.path.exists(hdr), "run contactimplicitmpc.jl_b200/modelgen/codegen.py"
    src = tmp_path / "shim.cpp"
    src.write_text(SHIM % {"tag": tag})
    so = tmp_path / "shim.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{GEN}", str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    res = get_residual(robot)
    nz, nth = res.idx.nz, res.idx.ntheta
    base = robot.replace("_payload", "")
    assert (nz, nth) == (SIZES[base][0] + 4 * SIZES[base][3] + 2 * SIZES[base][4], nth)
    nnz = lib.nnz()
    row = np.zeros(nnz, np.int32); col = np.zeros(nnz, np.int32)
    lib.pattern(row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p))
    nnzt = lib.nnzt()
    trow = np.zeros(nnzt, np.int32); tcol = np.zeros(nnzt, np.int32)
    lib.pattern_t(trow.ctypes.data_as(C.c_void_p), tcol.ctypes.data_as(C.c_void_p))
    rng = np.random.default_rng(0)
    for trial in range(5):
        z = rng.random(nz) + 0.1
        th = rng.random(nth) * 0.5 + 0.1
        th[-1] = 0.01 + 0.01 * rng.random()
        kappa = 1e-4 * trial
        r = np.zeros(nz); J = np.zeros(nnz)
        lib.r_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        lib.r_eval(z.ctypes.data, th.ctypes.data, kappa, r.ctypes.data)
        lib.rz_eval(z.ctypes.data, th.ctypes.data, J.ctypes.data)
        ro, Jo = res.r(z, th, kappa), res.rz(z, th)
        scale = max(1.0, np.abs(ro).max())
        assert np.abs(r - ro).max() <= 1e-10 * scale
        dense = np.zeros((nz, nz)); dense[row, col] = J
        assert np.abs(dense - Jo).max() <= 1e-9 * max(1.0, np.abs(Jo).max())
        # the pattern is exactly the structural non-zero set (bilinear diagonals included)
        assert set(zip(*np.nonzero(Jo))) <= set(zip(row, col))
        # rθ (the third generated function of code_gen_simulation.jl:150-168)
        Jt = np.zeros(nnzt)
        lib.rth_eval(z.ctypes.data, th.ctypes.data, Jt.ctypes.data)
        To = res.rth(z, th)
        dense_t = np.zeros((nz, nth)); dense_t[trow, tcol] = Jt
        assert np.abs(dense_t - To).max() <= 1e-9 * max(1.0, np.abs(To).max())
        assert set(zip(*np.nonzero(To))) <= set(zip(trow, tcol))


@pytest.mark.parametrize("robot,tag", [("hopper_2D_piecewise", "hopper2d_piecewise"), ("flamingo_piecewise", "flamingo_piecewise"),
                                       ("quadruped_piecewise", "quadruped_piecewise")])
def test_generated_terrain_residual_matches_oracle(tmp_path, robot, tag):
    """Planar robots on `piecewise1_2D_lc` (get_simulation(robot, "piecewise1_2D_lc", "piecewise", approx = true)): the
    generated r and the `approx` Jacobians (surface rotation held fixed, src/simulation/residual_approx.jl:14-99) against
    the oracle, whose Jacobians are complex-step derivatives — on the flat part, both ramps and inside both smoothed kinks."""
    from oracle.residual import get_residual
    src = tmp_path / "shim.cpp"
    src.write_text(SHIM % {"tag": tag})
    so = tmp_path / "shim.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{GEN}", str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.r_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    res = get_residual(robot)
    nz, nth = res.idx.nz, res.idx.ntheta
    nnz, nnzt = lib.nnz(), lib.nnzt()
    row = np.zeros(nnz, np.int32); col = np.zeros(nnz, np.int32)
    lib.pattern(row.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p))
    trow = np.zeros(nnzt, np.int32); tcol = np.zeros(nnzt, np.int32)
    lib.pattern_t(trow.ctypes.data_as(C.c_void_p), tcol.ctypes.data_as(C.c_void_p))
    rng = np.random.default_rng(1)
    sections = set()
    for x_body in (-1.0, 0.2, 0.35, 0.45, 0.5, 0.55, 0.65, 1.0, 1.6, 1.85, 1.95, 2.0, 2.05, 2.15, 2.6, 4.0):
        z = rng.random(nz) + 0.1
        z[res.idx.q2[2:]] = 0.6 * rng.standard_normal(res.model.nq - 2)
        z[0] = x_body
        th = rng.random(nth) * 0.5 + 0.1
        th[-1] = 0.01 + 0.01 * rng.random()
        for x in res._px(z[res.idx.q2]):
            sections.add(int(np.searchsorted([0.4, 0.6, 1.9, 2.1], x)))
        r = np.zeros(nz); J = np.zeros(nnz); Jt = np.zeros(nnzt)
        lib.r_eval(z.ctypes.data, th.ctypes.data, 1e-4, r.ctypes.data)
        lib.rz_eval(z.ctypes.data, th.ctypes.data, J.ctypes.data)
        lib.rth_eval(z.ctypes.data, th.ctypes.data, Jt.ctypes.data)
        ro, Jo, To = res.r(z, th, 1e-4), res.rz(z, th), res.rth(z, th)
        assert np.abs(r - ro).max() <= 1e-10 * max(1.0, np.abs(ro).max())
        dense = np.zeros((nz, nz)); dense[row, col] = J
        assert np.abs(dense - Jo).max() <= 1e-9 * max(1.0, np.abs(Jo).max())
        dense_t = np.zeros((nz, nth)); dense_t[trow, tcol] = Jt
        assert np.abs(dense_t - To).max() <= 1e-9 * max(1.0, np.abs(To).max())
    assert sections == {0, 1, 2, 3, 4}, sections  # every piece of the terrain was under some contact
