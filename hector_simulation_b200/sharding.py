"""Batch sharding across GPUs.

Robots are independent, so the batch is cut into contiguous slices, one per rank (one process per
GPU); every rank runs the identical kernel on its slice and the path's only exchange is ONE
all_gather of the results (wrench floats + status words) — issued only when the batch spans more than
one device.  `torch.distributed` is plumbing here: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) per rank; the first `batch % world` ranks get one extra robot."""
    base, extra = divmod(batch, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def solve_sharded(records: np.ndarray, horizon: int, solve_local: Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]],
                  group=None, device: torch.device | None = None):
    """Each rank solves its slice of `records` with `solve_local` (-> wrench [b,12N] f64, status [b] i32)
    and all ranks end up with the full [B,12N] / [B] results.  With world_size 1 no collective is issued."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = records.shape[0]
    bounds = shard_bounds(B, world)
    lo, hi = bounds[rank]
    w_loc, s_loc = solve_local(records[lo:hi])
    if world == 1:
        return w_loc, s_loc
    dev = device if device is not None else torch.device("cpu")
    width = 12 * horizon
    bmax = max(h - l for l, h in bounds)
    # one fused buffer per rank: [bmax, width + 1] (status carried as the last column) -> ONE all_gather
    buf = torch.zeros((bmax, width + 1), dtype=torch.float64, device=dev)
    buf[: hi - lo, :width] = torch.from_numpy(np.ascontiguousarray(w_loc)).to(dev)
    buf[: hi - lo, width] = torch.from_numpy(s_loc.astype(np.float64)).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    wrench = np.zeros((B, width), dtype=np.float64)
    status = np.zeros(B, dtype=np.int32)
    for r, (l, h) in enumerate(bounds):
        g = gathered[r].cpu().numpy()
        wrench[l:h] = g[: h - l, :width]
        status[l:h] = g[: h - l, width].astype(np.int32)
    return wrench, status
