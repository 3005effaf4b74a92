"""Environment sharding across GPUs (one process per GPU) and synthetic workload generation.

Environments are independent (no cross-environment term anywhere on the path),
so the batch is split into contiguous blocks, one per rank, with NO per-step communication.
The only collectives are the end-of-run gather of per-environment returns
(``accum_obj_val``) and the reduction of step / evaluation counters.

Shard boundaries are multiples of ``BLOCK`` environments so that block-keyed synthetic inputs
(bench_workload.py) are identical however many ranks the batch is sharded over.
"""
from __future__ import annotations


BLOCK = 1024        # shard granularity (environments)

def shard_range(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [lo, hi) block of global environment indices owned by ``rank``: whole
    blocks of BLOCK environments, the first ``nblocks % world_size`` ranks get one block more."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if num_envs % BLOCK:
        raise ValueError(f"num_envs must be a multiple of {BLOCK}")
    nb = num_envs // BLOCK
    base, extra = divmod(nb, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * BLOCK, hi * BLOCK


def gather_returns(local_returns, local_counts, group=None):
    """End-of-run collectives: all-gather of the per-environment returns (rank order = global
    environment order) and sum-reduction of the integer counters.  ``local_returns`` is a 1-D
    tensor (CUDA with NCCL, CPU with gloo), ``local_counts`` a 1-D int64 tensor.
    Ranks may own different numbers of environments."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_returns.clone(), local_counts.clone()
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_returns.numel()], dtype=torch.int64, device=local_returns.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros((pad,), dtype=local_returns.dtype, device=local_returns.device)
    buf[: local_returns.numel()] = local_returns
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    counts = local_counts.clone()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)]), counts


def gather_trajectories(local_rows, local_count, group=None, dst=0):
    """End-of-run gather of the device-side trajectory rings (``ops.TrajectoryLog``): ``local_rows``
    [capacity, ncols, E_local] and ``local_count`` [E_local] of every rank -> on rank ``dst`` the concatenation over
    the environment axis in global environment order ([capacity, ncols, E], [E]); ``(None, None)`` elsewhere.
    Ranks may own different numbers of environments.  NCCL on GPUs, gloo on CPU tensors."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows, local_count
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local = torch.tensor([local_rows.shape[2]], dtype=torch.int64, device=local_rows.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    cap, ncols = local_rows.shape[0], local_rows.shape[1]
    buf = torch.zeros((cap, ncols, pad), dtype=local_rows.dtype, device=local_rows.device)
    buf[:, :, : local_rows.shape[2]] = local_rows
    cnt = torch.zeros((pad,), dtype=local_count.dtype, device=local_count.device)
    cnt[: local_count.numel()] = local_count
    # all_gather keeps the code path identical for NCCL and gloo (gather is not available on every NCCL build)
    parts = [torch.empty_like(buf) for _ in range(world)]
    cparts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    dist.all_gather(cparts, cnt, group=group)
    if rank != dst:
        return None, None
    return (torch.cat([p[:, :, :k] for p, k in zip(parts, sizes)], dim=2),
            torch.cat([c[:k] for c, k in zip(cparts, sizes)]))


def ring_rows(rows, count, e):
    """Chronological rows of environment ``e`` out of a (gathered) ring: [min(count, capacity), ncols]."""
    cap = rows.shape[0]
    c = int(count[e])
    k = min(c, cap)
    idx = [(c - k + i) % cap for i in range(k)]
    return rows[idx, :, e]
