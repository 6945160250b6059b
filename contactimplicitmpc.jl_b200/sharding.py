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


def init_comm(im_traj, world: int, rank: int, group=None) -> None:
    """Create the context's NCCL communicator (`cimpc_comm_init`).  The 128-byte unique id is produced by rank 0
    (`cimpc_nccl_get_unique_id`) and handed round with whatever the host has — here one `torch.distributed` broadcast
    (any backend, gloo in the tests); a Julia host would use MPI.jl / a file / NCCL.jl's `UniqueID`."""
    import ctypes as C
    from . import capi
    idb = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        capi.check(im_traj._ctx, im_traj.lib.cimpc_nccl_get_unique_id(C.cast(buf, C.c_void_p)))
        idb = torch.tensor(list(buf), dtype=torch.uint8)
    if world > 1:
        backend = dist.get_backend(group)
        t = idb.cuda() if backend == "nccl" else idb
        dist.broadcast(t, src=0, group=group)
        idb = t.cpu()
    arr = (C.c_uint8 * 128)(*idb.tolist())
    capi.check(im_traj._ctx, im_traj.lib.cimpc_comm_init(im_traj._ctx, int(world), int(rank), C.cast(arr, C.c_void_p)))


def gather_rollouts_capi(im_traj, local: torch.Tensor, n_rollouts: int, world: int, rank: int, root: int = 0,
                         stream=None) -> torch.Tensor | None:
    """`cimpc_gather`: the rows of every rank (first dim = this rank's rollouts, `shard_rollouts` order, fp64 CUDA,
    contiguous) land on `root` back to back — NCCL point-to-point over NVLink, driven through the C ABI (no
    torch.distributed on the data path).  Returns the (n_rollouts, ...) tensor on root, None elsewhere."""
    import ctypes as C
    from . import capi
    assert local.is_cuda and local.dtype == torch.float64 and local.is_contiguous()
    row = int(local[0].numel()) if local.shape[0] else int(torch.tensor(local.shape[1:]).prod().item())
    sizes = [shard_rollouts(n_rollouts, world, r) for r in range(world)]
    counts = (C.c_int64 * world)(*[(hi - lo) * row for lo, hi in sizes])
    out = torch.empty((n_rollouts,) + tuple(local.shape[1:]), dtype=torch.float64, device=local.device) if rank == root else None
    if stream is None:
        stream = torch.cuda.current_stream(local.device).cuda_stream
    capi.check(im_traj._ctx, im_traj.lib.cimpc_gather(
        im_traj._ctx, None, local.data_ptr(), int(local.numel()), out.data_ptr() if out is not None else None,
        C.cast(counts, C.c_void_p), int(root), C.c_void_p(stream)))
    return out


def sum_statistics(local: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise sum over ranks (iteration histograms, failure counts)."""
    out = local.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
