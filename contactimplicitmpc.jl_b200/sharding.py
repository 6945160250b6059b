"""Multi-GPU plumbing of the Monte-Carlo batch: rollouts are independent, so they are
split statically across ranks (one process per GPU) and nothing is exchanged inside a solve; the only
collectives are the final gather of per-rollout results to rank 0 (`examples/quadruped/monte_carlo.jl:83-90`
collects the trajectories of all runs) and a sum of iteration statistics.  Backend-agnostic
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rollouts(n_rollouts: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one rollout."""
    base, rem = divmod(int(n_rollouts), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rollout_results(local: torch.Tensor, n_rollouts: int, group=None) -> torch.Tensor | None:
    """Gather the per-rollout rows of every rank (first dim = this rank's rollouts, shard_rollouts order)
    on rank 0; returns the (n_rollouts, ...) tensor there and None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_rollouts(n_rollouts, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def sum_statistics(local: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise sum over ranks (iteration histograms, failure counts)."""
    out = local.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
