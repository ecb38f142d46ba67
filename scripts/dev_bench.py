"""Development timings of the line kernels on ONE GPU (bench.py is the contract; this is the quick A/B tool).

    python scripts/dev_bench.py --workload cfg5 [--ranks 8 --rank 0] [--schedule 0|1|2] [--parts P] [--quantum Q] [--reps 2]

--ranks N --rank r runs the slab rank r of an N-GPU run would carry (mu interleaved round-robin over the ranks), so the
multi-GPU per-rank kernel time of a fixed grid can be measured on a single GPU.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A  # noqa: E402
from julia_relaxtime_b200._lib import Engine  # noqa: E402
from julia_relaxtime_b200.scan import build_grid  # noqa: E402

XI8 = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
WORK = {
    "cfg5": (XI8, 1024, (0.0, 400.0), 1024, (50.0, 300.0), 64, 16),
    "cfg4": ([0.0], 2048, (260.0, 330.0), 2048, (100.0, 160.0), 64, 16),
    "cfg2": ([0.0], 128, (0.0, 400.0), 128, (50.0, 300.0), 12, 6),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--ranks", type=int, default=1)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--contiguous", action="store_true", help="contiguous mu slabs instead of round-robin")
    ap.add_argument("--schedule", type=int, default=0)
    ap.add_argument("--parts", type=int, default=0)
    ap.add_argument("--quantum", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--n-t", type=int, default=0)
    args = ap.parse_args()
    if args.workload == "cfg3":
        xis, n_mu, (m0, m1), n_T, (t0, t1), p, t = XI8, 256, (0.0, 400.0), 256, (50.0, 300.0), 64, 16
        gx, gm, gT = np.meshgrid(np.asarray(xis), np.linspace(m0, m1, n_mu), np.linspace(t0, t1, n_T), indexing="ij")
        sel = np.arange(args.rank, gx.size, args.ranks)
        if args.n_t:
            sel = sel[:: max(1, len(sel) // args.n_t)]
        e = Engine(p_num=p, t_num=t, max_iter=40, schedule=args.schedule)
        if args.parts:
            e.set_option("march_parts", args.parts)
        rec = np.empty((len(sel), A.REC_DOUBLES))
        best = 1e30
        for _ in range(args.reps):
            e.solve_points(gT.ravel()[sel] / 197.327, gm.ravel()[sel] / 197.327, gx.ravel()[sel], A.SEED_MULTI, out=rec)
            best = min(best, e.stats()["kernel_ms"])
        st = e.stats()
        conv = ((rec[:, A.REC_STATUS].astype(np.int64) & 1) != 0).sum()
        passes = (rec[:, A.REC_NEVAL].sum() + rec[:, A.REC_NTHERMO].sum() + rec[:, A.REC_NFUSED].sum()) / len(sel)
        print("cfg3 rank %d/%d sched %d: %d MultiSeed points, converged %d | kernel %.2f ms -> %.3f M points/s (%.1f M passes/s) | "
              "passes/pt %.1f | lanes/solve %d blocks %d threads %d" % (args.rank, args.ranks, args.schedule, len(sel), conv, best,
              len(sel) / best / 1e3, len(sel) * passes / best / 1e3, passes, st["lanes_per_solve"], st["blocks"], st["threads"]), flush=True)
        return
    xis, n_mu, (m0, m1), n_T, (t0, t1), p, t = WORK[args.workload]
    if args.n_t:
        n_T = args.n_t
    mus = np.linspace(m0, m1, n_mu)
    T = np.linspace(t0, t1, n_T)
    grid = build_grid(xis, 3.0 * mus, T)
    idx = np.arange(grid.n_lines)
    im = idx % n_mu
    if args.contiguous:
        lo, hi = args.rank * n_mu // args.ranks, (args.rank + 1) * n_mu // args.ranks
        mine = idx[(im >= lo) & (im < hi)]
    else:
        mine = idx[im % args.ranks == args.rank]
    e = Engine(p_num=p, t_num=t, max_iter=40, schedule=args.schedule, lanes_per_solve=args.lanes)
    e.set_boundaries(grid.tables)
    if args.parts:
        e.set_option("march_parts", args.parts)
    if args.quantum:
        e.set_option("march_quantum", args.quantum)
    rec = np.empty((len(mine), n_T, A.REC_DOUBLES))
    best = 1e30
    for _ in range(args.reps):
        w0 = time.time()
        e.scan_lines(grid.muq_MeV[mine], grid.xi[mine], T, grid.table_idx[mine], out=rec)
        wall = time.time() - w0
        best = min(best, e.stats()["kernel_ms"])
    st = e.stats()
    r = rec.reshape(-1, A.REC_DOUBLES)
    npts = r.shape[0]
    conv = ((r[:, A.REC_STATUS].astype(np.int64) & 1) != 0).sum()
    passes = (r[:, A.REC_NEVAL].sum() + r[:, A.REC_NTHERMO].sum() + r[:, A.REC_NFUSED].sum()) / npts
    print("%s rank %d/%d sched %d: %d lines x %d T = %d pts, converged %d | kernel %.2f ms -> %.3f M points/s | passes/pt %.3f | "
          "lanes/solve %d blocks %d threads %d regs %d smem %d | wall %.0f ms" % (
              args.workload, args.rank, args.ranks, args.schedule, len(mine), n_T, npts, conv, best, npts / best / 1e3, passes,
              st["lanes_per_solve"], st["blocks"], st["threads"], st["regs_per_thread"], st["smem_bytes"], wall * 1e3), flush=True)


if __name__ == "__main__":
    main()
