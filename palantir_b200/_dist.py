"""Sharding helpers: one process per GPU, torch.distributed for the plumbing.

The path shards where the work is independent per unit -- kNN query rows and
shortest-path sources -- and exchanges only small results:
neighbour-list row shards (allgather) and per-cell partial sums (allreduce).
Without an initialised process group every helper is the identity, so single-GPU
runs take exactly the same code path.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, parts: int, align: int = 1) -> List[int]:
    """parts+1 monotone boundaries over range(n); interior ones are multiples of ``align``."""
    b = [0]
    for r in range(1, parts):
        x = (n * r // parts) // align * align
        b.append(min(n, max(b[-1], x)))
    b.append(n)
    return b


def row_shard(n: int, align: int = 1) -> Tuple[int, int]:
    """(first row, number of rows) owned by this rank."""
    rank, size = world()
    b = shard_bounds(n, size, align)
    return b[rank], b[rank + 1] - b[rank]


def allgather_rows(idx: torch.Tensor, val: torch.Tensor, n: int, align: int = 1, bounds=None):
    """Gather row shards (idx [nq,k], val [nq,k]) produced under ``row_shard(n, align)`` -- or under the explicit
    ``bounds`` (size+1 row boundaries) -- into full [n,k] tables on every rank (uneven shards are padded to the
    largest one)."""
    rank, size = world()
    if size == 1:
        return idx, val
    b = list(bounds) if bounds is not None else shard_bounds(n, size, align)
    most = max(b[r + 1] - b[r] for r in range(size))
    outs = []
    for t in (idx, val):
        k = t.shape[1]
        pad = torch.zeros((most, k), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        full = torch.empty((size * most, k), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, pad)
        outs.append(torch.cat([full[r * most: r * most + (b[r + 1] - b[r])] for r in range(size)], dim=0))
    return outs[0], outs[1]


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    _, size = world()
    if size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    _, size = world()
    if size > 1:
        dist.broadcast(t, src=src)
    return t
