"""World-size-N check of `cimpc_gather` (NCCL through the C ABI).  Launched by tests/test_gpu_gather.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/run_gather_nccl.py
torch.distributed (gloo) only hands the 128-byte NCCL unique id round; the data path is cimpc_comm_init + cimpc_gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cimpc_b200 as cb  # noqa: E402
from common import SIZES, load_lin  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lin = load_lin("quadruped")
    im = cb.ImplicitTrajectory(*SIZES["quadruped"], lin["z0"], lin["th0"], lin["r0"], lin["rz0"], lin["rth0"], device=local)
    cb.init_comm(im, world, rank)
    n_rollouts = 37  # ragged shards
    lo, hi = cb.shard_rollouts(n_rollouts, world, rank)
    full = np.arange(n_rollouts * 12 * 11, dtype=np.float64).reshape(n_rollouts, 12, 11)  # (rollout, recorded step, nq)
    mine = torch.from_numpy(full[lo:hi].copy()).to(dev)
    for root in range(min(world, 2)):
        out = cb.gather_rollouts_capi(im, mine, n_rollouts, world, rank, root=root)
        torch.cuda.synchronize()
        if rank == root:
            assert out is not None and np.array_equal(out.cpu().numpy(), full), "gathered trajectories differ"
        else:
            assert out is None
    dist.barrier()
    if rank == 0:
        print(f"GATHER_OK world={world}")
    im.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
