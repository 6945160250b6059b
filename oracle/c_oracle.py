"""ctypes wrapper of the C restatement `oracle/c/liboracle_ip.so` (TEST INFRASTRUCTURE ONLY).

Used by tests (fast checker at sizes the numpy oracle cannot reach) and by bench.py's
`cpu_baseline` / `--impl reference` legs.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liboracle_ip.so")


class COpts(C.Structure):
    _fields_ = [("r_tol", C.c_double), ("kappa_tol", C.c_double), ("eps_min", C.c_double),
                ("kappa_reg", C.c_double), ("gamma_reg", C.c_double), ("undercut", C.c_double),
                ("ls_scale", C.c_double), ("max_iter", C.c_int32), ("max_ls", C.c_int32),
                ("diff_sol", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load(build_if_missing: bool = True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        if not build_if_missing:
            raise FileNotFoundError(LIB)
        subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.oracle_create.argtypes = [C.c_int] * 7 + [vp] * 5
    lib.oracle_create.restype = vp
    lib.oracle_destroy.argtypes = [vp]
    lib.oracle_set_solver.argtypes = [vp, C.c_int]
    lib.oracle_ip_solve_batch.argtypes = [vp, C.c_int64, vp, vp, vp, vp, C.POINTER(COpts), vp, vp, vp, vp, C.c_int]
    lib.oracle_ip_solve_batch.restype = C.c_int
    lib.oracle_max_threads.restype = C.c_int
    _lib = lib
    return lib


class COracle:
    def __init__(self, nq, nu, nw, nc, nb, lin: dict, mode="configuration", solver="mgs"):
        """solver = "mgs": the reference's QR (CPU baseline); "lu": accurate test variant."""
        self.lib = load()
        self.nq, self.nu, self.nw, self.nc, self.nb = nq, nu, nw, nc, nb
        self.mode = mode
        self.nz = nq + 4 * nc + 2 * nb
        self.nth = 2 * nq + nu + nw + 2
        self.nd = nq if mode == "configuration" else nq + nc + nb
        self.ncol = 2 * nq + nu
        H = lin["z0"].shape[0]
        z0 = np.ascontiguousarray(lin["z0"], dtype=np.float64)
        th0 = np.ascontiguousarray(lin["th0"], dtype=np.float64)
        r0 = np.ascontiguousarray(lin["r0"], dtype=np.float64)
        rz0 = np.ascontiguousarray(np.transpose(lin["rz0"], (0, 2, 1)), dtype=np.float64)
        rth0 = np.ascontiguousarray(np.transpose(lin["rth0"], (0, 2, 1)), dtype=np.float64)
        self.h = self.lib.oracle_create(nq, nu, nw, nc, nb, 0 if mode == "configuration" else 1, H,
                                        z0.ctypes.data, th0.ctypes.data, r0.ctypes.data, rz0.ctypes.data,
                                        rth0.ctypes.data)
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.lib.oracle_set_solver(self.h, 1 if solver == "lu" else 0)

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def solve(self, knot, theta, q2, opts, alt=None, threads=0):
        """opts: any object with the IPOptions field names.  Returns z, dz (n, nd, ncol), status, iters."""
        knot = np.ascontiguousarray(knot, dtype=np.int32)
        n = knot.shape[0]
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        q2 = np.ascontiguousarray(q2, dtype=np.float64)
        altp = None
        if alt is not None:
            alt = np.ascontiguousarray(alt, dtype=np.float64)
            altp = alt.ctypes.data
        co = COpts(opts.r_tol, opts.kappa_tol, opts.eps_min, opts.kappa_reg, opts.gamma_reg, opts.undercut,
                   opts.ls_scale, opts.max_iter, opts.max_ls, int(opts.diff_sol), 0)
        z = np.empty((n, self.nz))
        dz = np.empty((n, self.ncol, self.nd)) if opts.diff_sol else None
        st = np.zeros(n, dtype=np.uint8)
        it = np.zeros(n, dtype=np.int32)
        rc = self.lib.oracle_ip_solve_batch(self.h, n, knot.ctypes.data, theta.ctypes.data, q2.ctypes.data, altp,
                                            C.byref(co), z.ctypes.data, dz.ctypes.data if dz is not None else None,
                                            st.ctypes.data, it.ctypes.data, threads)
        if rc != 0:
            raise RuntimeError("oracle_ip_solve_batch failed")
        if dz is not None:
            dz = np.transpose(dz, (0, 2, 1))
        return z, dz, st.astype(bool), it
