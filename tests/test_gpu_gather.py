"""`cimpc_gather` — the trajectory gather of the Monte-Carlo workload behind the C ABI (NCCL bound at run time, no
PyTorch on the data path; SURVEY.md §8e, examples/quadruped/monte_carlo.jl:83-90).  With two or more visible GPUs the
world-size-2 case runs under torchrun; the single-rank path (a device copy) is always checked."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, SIZES, load_lin

pytestmark = pytest.mark.gpu


def test_gather_single_rank(cuda_device):
    import torch
    import cimpc_b200 as cb
    lin = load_lin("quadruped")
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"])
    x = torch.arange(5 * 7 * 11, dtype=torch.float64, device=cuda_device).reshape(5, 7, 11)
    out = cb.gather_rollouts_capi(im, x, 5, 1, 0)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), x.cpu().numpy())


def test_gather_world_size_2(cuda_device):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "run_gather_nccl.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "GATHER_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
