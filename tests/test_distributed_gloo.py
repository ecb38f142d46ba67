"""world_size-2 test of the multi-GPU host logic on CPU (gloo): mu-slab partition, per-rank compute, final gather
and re-ordering into global line order.  The per-rank compute here is the CPU oracle (test stand-in for the CUDA
engine, which needs a GPU); what is under test is julia_relaxtime_b200.distributed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200.distributed import scan_sharded
from julia_relaxtime_b200.scan import build_grid


def _oracle_compute(grid):
    from oracle.oracle import Oracle
    o = Oracle(p_num=8, t_num=4, max_iter=40, n_threads=1)

    def compute(lines):
        res = o.scan_lines(grid.muq_MeV[lines], grid.xi[lines], grid.T_MeV, grid.tables, grid.table_idx[lines])
        rec = np.zeros((len(lines) * grid.n_T, A.REC_DOUBLES))
        rec[:, 0:5] = res.x.T
        rec[:, 5:8] = res.mass.T
        rec[:, A.REC_OMEGA] = res.omega
        rec[:, A.REC_STATUS] = res.status
        return torch.from_numpy(rec.reshape(len(lines), grid.n_T, A.REC_DOUBLES))
    return compute


def _worker(rank, world, port, out_path, layout):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = build_grid([0.0, 0.2], np.linspace(0, 1200, 5), np.linspace(100, 200, 6))
        full, local = scan_sharded(grid, 2, 5, _oracle_compute(grid), rank, world, layout=layout)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
        assert local.shape[0] == (6 if rank == 0 else 4)     # 5 mu over 2 ranks: 3 + 2, times 2 xi
    finally:
        dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("layout", ["interleaved", "slab"])
def test_two_rank_scan_equals_single_process(tmp_path, layout):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out, layout), nprocs=2, join=True)
    got = np.load(out)
    grid = build_grid([0.0, 0.2], np.linspace(0, 1200, 5), np.linspace(100, 200, 6))
    ref = _oracle_compute(grid)(np.arange(grid.n_lines)).numpy()
    assert got.shape == ref.shape == (10, 6, A.REC_DOUBLES)
    np.testing.assert_array_equal(got, ref)
