"""N > 1 path on CPU: world_size-2 `gloo` processes shard the rollout batch, solve their shard (here with
the C oracle standing in for the GPU — tests may use the oracle, the product never does), gather on rank 0
and must reproduce the single-process result bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import SIZES, load_gait, load_lin, make_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solve(lin, knot, theta, q2):
    from oracle.c_oracle import COracle
    from oracle.ip import IPOptions
    co = COracle(*SIZES["quadruped"], lin, mode="configuration", solver="lu")
    return co.solve(knot, theta, q2, IPOptions(r_tol=1e-4, kappa_tol=1e-4, diff_sol=False, max_ls=0), threads=1)


def _worker(rank, world, port, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cimpc_b200 as cb
    lin, gait = load_lin("quadruped"), load_gait("quadruped")
    knot, theta, q2 = make_batch("quadruped", lin, gait, n, seed=31)
    lo, hi = cb.shard_rollouts(n, world, rank)
    z, _, st, it = _solve(lin, knot[lo:hi], theta[lo:hi], q2[lo:hi])
    zg = cb.gather_rollout_results(torch.from_numpy(z), n)
    hist = cb.sum_statistics(torch.bincount(torch.from_numpy(it.astype(np.int64)), minlength=32))
    if rank == 0:
        np.savez(out_path, z=zg.numpy(), hist=hist.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    import cimpc_b200 as cb
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            spans = [cb.shard_rollouts(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_and_gather(tmp_path):
    n = 101  # odd: unequal shards
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    lin, gait = load_lin("quadruped"), load_gait("quadruped")
    knot, theta, q2 = make_batch("quadruped", lin, gait, n, seed=31)
    z, _, st, it = _solve(lin, knot, theta, q2)
    got = np.load(out)
    assert np.array_equal(got["z"], z)
    assert np.array_equal(got["hist"], np.bincount(it, minlength=32))
