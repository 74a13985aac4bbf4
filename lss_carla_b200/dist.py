"""Multi-GPU plumbing of the lift-splat path: batch sharding and timing reduction.

The path shards by batch with no exchange (the batch index is part of the reference's voxel key,
src/models.py:214-216, :229), so a rank only needs to know WHICH samples are its own; the one collective
used for measurement is a MAX over the ranks' device times.  Works with any torch.distributed backend
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of `rank`: sizes differ by at most one, earlier ranks take the extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: dict, rank: int, world: int, batch_keys=None):
    """Slice every per-sample tensor of a synthetic batch to this rank's shard.  `depthnet_out` is laid out
    [B*N, ...] (camera-major inside a sample) and is sliced by B*N rows."""
    B = batch["trans"].shape[0]
    N = batch["trans"].shape[1]
    lo, hi = shard_range(B, rank, world)
    out = {}
    for k, v in batch.items():
        if batch_keys is not None and k not in batch_keys:
            out[k] = v
        elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B:
            out[k] = v[lo:hi]
        elif torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B * N:
            out[k] = v[lo * N:hi * N]
        else:
            out[k] = v
    return out


def max_over_ranks(seconds: float, device=None) -> float:
    """Device time of the slowest rank (what a whole-job throughput has to be divided by)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_this_rank: float, seconds: float, device=None) -> float:
    """units of ALL ranks / time of the slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return units_this_rank / seconds
    u = torch.tensor([units_this_rank], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / max_over_ranks(seconds, device)
