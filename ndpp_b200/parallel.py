"""Sharding of the E_in axis over ranks and assembly of the moment arrays.

Every (nuclide, E_in) column of every matrix depends only on read-only per-nuclide tables, so the path
shards without any exchange step; the one collective is the gather of the `[NE_part][G*L]` slabs to the
rank that hands the matrix back to the reference's driver.  The only cross-column rule -- an E_in above
the top group edge copies the previous column (src/scatt.F90:669,770) -- is applied after the gather,
because the previous column may live on another rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous block of the E_in grid owned by `rank` (first blocks take the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_columns(local: torch.Tensor, n_total: int, dst: int = 0):
    """Gather per-rank column blocks `[n_local][GL]` (possibly of unequal height) to `dst`.
    Returns the assembled `[n_total][GL]` tensor on dst, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    GL = local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local
    if local.shape[0] < pad:
        buf = torch.zeros((pad, GL), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    full = torch.empty((n_total, GL), dtype=local.dtype, device=local.device)
    for (lo, hi), p in zip(sizes, parts):
        full[lo:hi] = p[: hi - lo]
    return full


def copy_top_columns(mat: torch.Tensor, Ein: torch.Tensor, e_top: float):
    """Columns whose E_in exceeds the top group edge copy their predecessor, in order."""
    idx = torch.nonzero(~(Ein <= e_top)).flatten().tolist()
    for j in idx:
        if j > 0:
            mat[j] = mat[j - 1]
    return mat
