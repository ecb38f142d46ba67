"""Multi-GPU scan: one process per GPU, the mu axis dealt out over the ranks, one final gather (BASELINE.json north_star §e).

The unit of independence is one (xi, muB) line (its own tracker, run_gap_transport_scan.jl:408-416), so the
grid shards with no halo and no exchange during the solve.  Two layouts, identical results:
  "interleaved" (default)  rank r owns the mu indices r, r + W, r + 2W, ... for every xi.  The cost of a line varies smoothly
                           with mu (the lines through the first-order / crossover band at mu_q = 280..360 MeV need more
                           iterations), so dealing mu round-robin gives every rank the same mix and the same run time;
  "slab"                   rank r owns the contiguous indices [r*n_mu/W, (r+1)*n_mu/W): the rank holding the band is the
                           slowest (round 1: 330 ms against 312 ms at 8 ranks).
The only collective is the gather of the result records to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests), or
none at all when the kernels store straight into rank 0's array (PeerRecords).  torch.distributed is plumbing only.
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


def rank_mu_indices(n_mu: int, rank: int, world_size: int, layout: str = "interleaved") -> np.ndarray:
    """mu indices owned by `rank`."""
    if layout == "interleaved":
        return np.arange(rank, n_mu, world_size)
    if layout == "slab":
        lo, hi = slab_bounds(n_mu, world_size)[rank]
        return np.arange(lo, hi)
    raise ValueError("layout must be 'interleaved' or 'slab'")


def rank_line_indices(n_xi: int, n_mu: int, rank: int, world_size: int, layout: str = "interleaved") -> np.ndarray:
    """Global line indices (xi-major order of scan.build_grid) owned by `rank`."""
    return (np.arange(n_xi)[:, None] * n_mu + rank_mu_indices(n_mu, rank, world_size, layout)[None, :]).reshape(-1)


def scan_sharded(grid: ScanGrid, n_xi: int, n_mu: int, compute: Callable, rank: int, world_size: int,
                 device: Optional[str] = None, group=None, gather: bool = True, layout: str = "interleaved"):
    """Run this rank's slab with `compute(line_indices) -> torch tensor [n_local, n_T, 32]` (on `device`) and
    gather everything on rank 0 in global line order.  Returns (records on rank 0 | None, local records)."""
    import torch
    import torch.distributed as dist

    mine = rank_line_indices(n_xi, n_mu, rank, world_size, layout)
    local = compute(mine)
    if world_size == 1 or not gather:
        return (local if rank == 0 else None), local
    n_T = grid.n_T
    max_lines = max(rank_mu_indices(n_mu, r, world_size, layout).size for r in range(world_size)) * n_xi
    dev = local.device if device is None else torch.device(device)
    send = torch.zeros((max_lines, n_T, A.REC_DOUBLES), dtype=torch.float64, device=dev)
    send[:local.shape[0]] = local
    if rank == 0:
        bufs = [torch.empty_like(send) for _ in range(world_size)]
        dist.gather(send, bufs, dst=0, group=group)
        out = torch.empty((grid.n_lines, n_T, A.REC_DOUBLES), dtype=torch.float64, device=dev)
        for r in range(world_size):
            idx = torch.as_tensor(rank_line_indices(n_xi, n_mu, r, world_size, layout), device=dev)
            out[idx] = bufs[r][:idx.numel()]
        return out, local
    dist.gather(send, None, dst=0, group=group)
    return None, local


# ------------------------------------------------------------------------------------------------------------
# Gather fused into the kernel: every rank's scan kernel stores its records straight into rank 0's result array
# (CUDA IPC mapping of one buffer, NVLink peer stores), so there is no collective after the kernel — only a barrier.
# ------------------------------------------------------------------------------------------------------------
class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a raw device allocation."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerRecords:
    """The result array [n_lines, n_T, 32] of the whole grid, allocated on rank 0 and mapped into every other rank of the
    node.  `ptr` is the address to hand to `Engine.scan_lines_device_indexed` on this rank; `tensor` (rank 0 only) is a torch
    view of it.  Collective constructor (one small broadcast of the IPC handle); close() is collective too."""

    def __init__(self, n_lines, n_T, rank, world_size, device, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from ._lib import PnjlError, load
        self.L, self.rank, self.world, self.group = load(), rank, world_size, group
        self.shape = (int(n_lines), int(n_T), A.REC_DOUBLES)
        nbytes = 8 * self.shape[0] * self.shape[1] * self.shape[2]
        handle = C.create_string_buffer(64)
        p = C.c_void_p()
        ok = torch.ones(1, dtype=torch.int32, device=device)
        if rank == 0:
            if self.L.pnjl_ipc_alloc(C.c_uint64(nbytes), C.byref(p), handle) != 0:
                ok.zero_()
        h = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(device)
        if world_size > 1:
            dist.broadcast(h, src=0, group=group)
        if rank != 0:
            hb = C.create_string_buffer(bytes(h.cpu().numpy().tobytes()), 64)
            if self.L.pnjl_ipc_open(hb, C.byref(p)) != 0:
                ok.zero_()
        if world_size > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.ptr = p.value
        if int(ok.item()) == 0:
            err = self.L.pnjl_last_error().decode()
            self.close()            # collective on this path too: the all_reduce above was reached by every rank
            raise PnjlError("peer-visible result buffer could not be set up on every rank: %s" % err)
        self.tensor = torch.as_tensor(_DevArray(self.ptr, self.shape), device=device) if rank == 0 else None

    def close(self, collective=True):
        """Unmap on the peers, barrier, free on the owner (a peer must not hold the mapping when the owner frees it — also
        on the set-up failure path, where some ranks hold a mapping and others do not)."""
        import torch.distributed as dist
        if getattr(self, "_closed", False):
            return
        self._closed = True
        self.tensor = None
        ptr = getattr(self, "ptr", None)
        if self.rank != 0 and ptr:
            self.L.pnjl_ipc_close(ptr)
        if collective and self.world > 1:
            dist.barrier(group=self.group)          # every peer has unmapped before the owner frees
        if self.rank == 0 and ptr:
            self.L.pnjl_ipc_free(ptr)
        self.ptr = None


def scan_sharded_peer(engine, grid: ScanGrid, n_xi: int, n_mu: int, peer: PeerRecords, rank: int, world_size: int,
                      device, stream=None, inputs=None, group=None, layout: str = "interleaved"):
    """This rank's mu-slab, written by the kernel directly into `peer` (global line order, the order of scan.build_grid).
    `inputs` caches the device copies of this rank's line parameters between calls.  Returns (records on rank 0 | None,
    inputs).  The only communication is the closing barrier."""
    import torch
    import torch.distributed as dist

    if stream is None:
        stream = torch.cuda.current_stream(device).cuda_stream      # the stream the input tensors are created on
    if inputs is None:
        mine = rank_line_indices(n_xi, n_mu, rank, world_size, layout)
        inputs = dict(muq=torch.as_tensor(grid.muq_MeV[mine], device=device), xi=torch.as_tensor(grid.xi[mine], device=device),
                      tidx=torch.as_tensor(grid.table_idx[mine], device=device), T=torch.as_tensor(grid.T_MeV, device=device),
                      out_index=torch.as_tensor(mine.astype(np.int64), device=device))
    engine.scan_lines_device_indexed(inputs["muq"], inputs["xi"], inputs["tidx"], inputs["T"], peer.ptr, inputs["out_index"],
                                     stream)
    torch.cuda.synchronize(device)
    if world_size > 1:
        dist.barrier(group=group)                   # all slabs are in rank 0's buffer
    return (peer.tensor if rank == 0 else None), inputs
