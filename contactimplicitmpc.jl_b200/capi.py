"""ctypes binding of `lib/libcimpc_b200.so` (the C ABI declared in `include/cimpc_b200.h`).

This is the host-side stub a reference maintainer would write as `ccall`s in Julia
(see INTEGRATION.md); in this repository it is Python because the image has no Julia.
There is NO fallback: if the CUDA library is missing or no sm_100 GPU is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CIMPC_B200_LIB") or os.path.join(HERE, "lib", "libcimpc_b200.so")

# every symbol include/cimpc_b200.h declares
SYMBOLS = [
    "cimpc_ip_opts_default", "cimpc_status_string", "cimpc_last_cuda_error", "cimpc_version",
    "cimpc_create", "cimpc_destroy", "cimpc_get_dims", "cimpc_upload_linearization",
    "cimpc_ip_solve_batch", "cimpc_ip_solve_batch_host", "cimpc_launch_count",
    "cimpc_newton_opts_default", "cimpc_newton_create", "cimpc_newton_solve_batch", "cimpc_newton_last_sweeps",
    "cimpc_sim_step_batch", "cimpc_linearize", "cimpc_get_linearization",
    "cimpc_newton_create_ex", "cimpc_newton_solve_batch_ex", "cimpc_newton_solve_batch_ex2",
    "cimpc_newton_solve_batch_host", "cimpc_sim_step_batch_ex", "cimpc_newton_create_dense", "cimpc_create_named", "cimpc_ip_solve_batch_host_ex",
    "cimpc_nccl_get_unique_id", "cimpc_comm_init", "cimpc_gather", "cimpc_sim_steps_batch",
]


class ModelDesc(C.Structure):
    _fields_ = [("nq", C.c_int32), ("nu", C.c_int32), ("nw", C.c_int32), ("nc", C.c_int32),
                ("nb", C.c_int32), ("mode", C.c_int32)]


class Dims(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("ntheta", C.c_int32),
                ("nd", C.c_int32), ("ncol", C.c_int32), ("group", C.c_int32)]


class IPOpts(C.Structure):
    _fields_ = [("r_tol", C.c_double), ("kappa_tol", C.c_double), ("eps_min", C.c_double),
                ("kappa_reg", C.c_double), ("gamma_reg", C.c_double), ("undercut", C.c_double),
                ("ls_scale", C.c_double), ("max_iter", C.c_int32), ("max_ls", C.c_int32),
                ("diff_sol", C.c_int32), ("reserved", C.c_int32)]


class NewtonOpts(C.Structure):
    _fields_ = [("r_tol", C.c_double), ("beta_init", C.c_double), ("max_iter", C.c_int32), ("reserved", C.c_int32)]


class CimpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cimpc status {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python contactimplicitmpc.jl_b200/build.py` "
            "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(path)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.cimpc_ip_opts_default.argtypes = [C.POINTER(IPOpts)]
    lib.cimpc_ip_opts_default.restype = None
    lib.cimpc_status_string.argtypes = [C.c_int]
    lib.cimpc_status_string.restype = C.c_char_p
    lib.cimpc_last_cuda_error.argtypes = [vp]
    lib.cimpc_last_cuda_error.restype = C.c_char_p
    lib.cimpc_version.argtypes = []
    lib.cimpc_version.restype = C.c_int
    lib.cimpc_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(ModelDesc)]
    lib.cimpc_create.restype = C.c_int
    lib.cimpc_create_named.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(ModelDesc), C.c_char_p]
    lib.cimpc_create_named.restype = C.c_int
    lib.cimpc_destroy.argtypes = [vp]
    lib.cimpc_destroy.restype = C.c_int
    lib.cimpc_get_dims.argtypes = [vp, C.POINTER(Dims)]
    lib.cimpc_get_dims.restype = C.c_int
    lib.cimpc_upload_linearization.argtypes = [vp, i32, dp, dp, dp, dp, dp, vp]
    lib.cimpc_upload_linearization.restype = C.c_int
    lib.cimpc_ip_solve_batch.argtypes = [vp, i64, dp, dp, dp, dp, C.POINTER(IPOpts), dp, dp, dp, dp, vp]
    lib.cimpc_ip_solve_batch.restype = C.c_int
    lib.cimpc_ip_solve_batch_host.argtypes = [vp, i64, dp, dp, dp, dp, C.POINTER(IPOpts), dp, dp, dp, dp]
    lib.cimpc_ip_solve_batch_host.restype = C.c_int
    lib.cimpc_ip_solve_batch_host_ex.argtypes = [vp, i64, dp, dp, dp, dp, C.POINTER(IPOpts), dp, dp, dp, dp, C.c_uint32]
    lib.cimpc_ip_solve_batch_host_ex.restype = C.c_int
    lib.cimpc_nccl_get_unique_id.argtypes = [dp]
    lib.cimpc_nccl_get_unique_id.restype = C.c_int
    lib.cimpc_comm_init.argtypes = [vp, i32, i32, dp]
    lib.cimpc_comm_init.restype = C.c_int
    lib.cimpc_gather.argtypes = [vp, vp, dp, i64, dp, dp, i32, vp]
    lib.cimpc_gather.restype = C.c_int
    lib.cimpc_launch_count.argtypes = [vp]
    lib.cimpc_launch_count.restype = C.c_int64
    lib.cimpc_newton_opts_default.argtypes = [C.POINTER(NewtonOpts)]
    lib.cimpc_newton_opts_default.restype = None
    lib.cimpc_newton_create.argtypes = [vp, i32, i64, dp, dp, C.c_double, C.POINTER(NewtonOpts), C.POINTER(IPOpts)]
    lib.cimpc_newton_create.restype = C.c_int
    lib.cimpc_newton_solve_batch.argtypes = [vp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, i32, dp, dp, dp, vp]
    lib.cimpc_newton_solve_batch.restype = C.c_int
    lib.cimpc_newton_last_sweeps.argtypes = [vp]
    lib.cimpc_newton_last_sweeps.restype = i32
    lib.cimpc_sim_step_batch.argtypes = [vp, i64, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.POINTER(IPOpts), dp, dp, dp,
                                         dp, dp, vp]
    lib.cimpc_sim_step_batch.restype = C.c_int
    lib.cimpc_sim_step_batch_ex.argtypes = [vp, i64, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.POINTER(IPOpts), dp, dp,
                                            dp, dp, dp, dp, vp]
    lib.cimpc_sim_step_batch_ex.restype = C.c_int
    lib.cimpc_sim_steps_batch.argtypes = [vp, i64, i32, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.POINTER(IPOpts), dp,
                                          dp, dp, dp, dp, dp, vp]
    lib.cimpc_sim_steps_batch.restype = C.c_int
    lib.cimpc_newton_create_ex.argtypes = [vp, i32, i64, dp, dp, dp, dp, dp, C.c_double, C.POINTER(NewtonOpts),
                                           C.POINTER(IPOpts)]
    lib.cimpc_newton_create_ex.restype = C.c_int
    lib.cimpc_newton_create_dense.argtypes = [vp, i32, i64, dp, dp, dp, dp, C.c_double, C.POINTER(NewtonOpts),
                                              C.POINTER(IPOpts)]
    lib.cimpc_newton_create_dense.restype = C.c_int
    lib.cimpc_newton_solve_batch_ex.argtypes = [vp, dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, i32, dp, dp, dp,
                                                dp, vp]
    lib.cimpc_newton_solve_batch_ex.restype = C.c_int
    lib.cimpc_newton_solve_batch_ex2.argtypes = [vp, dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, dp, i32, dp, dp,
                                                 dp, dp, vp]
    lib.cimpc_newton_solve_batch_ex2.restype = C.c_int
    lib.cimpc_newton_solve_batch_host.argtypes = [vp, dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, dp, i32, dp, dp]
    lib.cimpc_newton_solve_batch_host.restype = C.c_int
    lib.cimpc_linearize.argtypes = [vp, i32, dp, dp, C.c_double, vp]
    lib.cimpc_linearize.restype = C.c_int
    lib.cimpc_get_linearization.argtypes = [vp, dp, dp, dp]
    lib.cimpc_get_linearization.restype = C.c_int
    _lib = lib
    return lib


def check(ctx, code):
    if code != 0:
        lib = load_library()
        msg = lib.cimpc_status_string(code).decode()
        if ctx:
            extra = lib.cimpc_last_cuda_error(ctx).decode()
            if extra:
                msg += f" [{extra}]"
        raise CimpcError(code, msg)
