"""Device-side accuracy of `csrc/fastdiv.cuh` (profiles/tools/rcp_check.cu, built by `__graft_entry__.build()`): over 2^24
operands per operation the branch-free reciprocal and the Markstein-corrected quotient must equal `__drcp_rn` / `__ddiv_rn`
on EVERY sample — that is what lets `ip_solve_kernel` use them without changing its results."""
import json
import os
import subprocess

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


def test_fast_reciprocal_and_division_are_bit_exact(cuda_device):
    exe = os.path.join(ROOT, "profiles", "tools", "rcp_check")
    if not os.path.exists(exe):
        pytest.skip("profiles/tools/rcp_check not built (python __graft_entry__.py)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["samples"] == 1 << 24
    assert d["rcp"]["mismatches_vs_drcp_rn"] == 0
    assert d["rcp"]["seed_max_rel_error"] < 2e-6          # the bound tests/test_fastdiv_math.py assumes
    for k in ("div_wide", "div_ip_range", "over_24", "div_by_clamp"):
        assert d[k]["mismatches_vs_ddiv_rn"] == 0, k
    assert d["div_by_clamp"]["rcp_of_quotient_mismatches"] == 0
    assert d["div_by_clamp"]["rcp_of_clamp_exact"] == [1, 1, 1, 1]
