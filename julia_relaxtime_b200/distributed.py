"""Multi-GPU scan: one process per GPU, contiguous mu-slabs, one final gather (BASELINE.json north_star §e).

The unit of independence is one (xi, muB) line (its own tracker, run_gap_transport_scan.jl:408-416), so the
grid shards with no halo and no exchange during the solve.  Rank r owns the mu indices
[r*n_mu/W, (r+1)*n_mu/W) for every xi; the only collective is the gather of the result records to rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).  torch.distributed is plumbing only.
"""
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _abi as A
from .scan import ScanGrid


def slab_bounds(n_mu: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal mu-slabs: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_mu, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def rank_line_indices(n_xi: int, n_mu: int, rank: int, world_size: int) -> np.ndarray:
    """Global line indices (xi-major order of scan.build_grid) owned by `rank`."""
    lo, hi = slab_bounds(n_mu, world_size)[rank]
    return (np.arange(n_xi)[:, None] * n_mu + np.arange(lo, hi)[None, :]).reshape(-1)


def scan_sharded(grid: ScanGrid, n_xi: int, n_mu: int, compute: Callable, rank: int, world_size: int,
                 device: Optional[str] = None, group=None, gather: bool = True):
    """Run this rank's slab with `compute(line_indices) -> torch tensor [n_local, n_T, 32]` (on `device`) and
    gather everything on rank 0 in global line order.  Returns (records on rank 0 | None, local records)."""
    import torch
    import torch.distributed as dist

    mine = rank_line_indices(n_xi, n_mu, rank, world_size)
    local = compute(mine)
    if world_size == 1 or not gather:
        return (local if rank == 0 else None), local
    n_T = grid.n_T
    max_lines = max(hi - lo for lo, hi in slab_bounds(n_mu, world_size)) * n_xi
    dev = local.device if device is None else torch.device(device)
    send = torch.zeros((max_lines, n_T, A.REC_DOUBLES), dtype=torch.float64, device=dev)
    send[:local.shape[0]] = local
    if rank == 0:
        bufs = [torch.empty_like(send) for _ in range(world_size)]
        dist.gather(send, bufs, dst=0, group=group)
        out = torch.empty((grid.n_lines, n_T, A.REC_DOUBLES), dtype=torch.float64, device=dev)
        for r in range(world_size):
            idx = torch.as_tensor(rank_line_indices(n_xi, n_mu, r, world_size), device=dev)
            out[idx] = bufs[r][:idx.numel()]
        return out, local
    dist.gather(send, None, dst=0, group=group)
    return None, local
