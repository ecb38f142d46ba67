"""Small exercise of every kernel path for compute-sanitizer runs (development aid): the warp-specialised kernel in lines,
points, T-mu and dual-branch mode (64x16 mesh, unsplit and split passes), k_couplings, zero-copy output."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200._lib import Engine, PinnedArray
from julia_relaxtime_b200.boundary import default_tables

tables, index = default_tables([0.0, 0.2])
e = Engine(p_num=64, t_num=16, max_iter=40)
e.set_boundaries(tables)
T = np.linspace(60.0, 260.0, 6)
muq = np.linspace(0.0, 400.0, 40)
xi = np.tile([0.0, 0.2, -0.4, 0.6], 10)
tidx = np.array([index.get(x, -1) for x in xi], dtype=np.int32)
rec, aux = e.scan_lines_couplings(muq, xi, T, tidx)                      # split passes (few lines per SM) + k_couplings
print("lines", rec.shape, "converged", int(((rec[..., A.REC_STATUS].astype(int) & 1) != 0).sum()), "aux finite", np.isfinite(aux).all())
big_mu = np.linspace(0.0, 400.0, 2 * 148 * 4)                            # > 4 lines per SM: unsplit passes, round-robin deal
rb = e.scan_lines(big_mu, np.zeros(big_mu.size) + 0.2, T[:2], np.full(big_mu.size, index[0.2], dtype=np.int32))
print("many lines", rb.shape, int(((rb[..., A.REC_STATUS].astype(int) & 1) != 0).sum()))
rp = e.solve_points(np.full(20, 150.0 / 197.327), np.linspace(0, 400, 20) / 197.327, np.full(20, 0.2), A.SEED_MULTI)
print("points", int(((rp[:, A.REC_STATUS].astype(int) & 1) != 0).sum()))
rt = e.tmu_scan([80.0, 150.0], [0.0, 0.2], np.linspace(0, 400, 9), np.array([index[0.0], index[0.2]], dtype=np.int32))
rd = e.dual_branch([60.0, 110.0], [0.2, 0.0], np.linspace(300, 400, 6))
print("tmu", rt.shape, "dual", rd.shape)
pa = PinnedArray((muq.size, T.size, A.REC_DOUBLES))
e.scan_lines(muq, xi, T, tidx, out=pa.array)
print("zero-copy equal", np.array_equal(pa.array, rec))
pa.close()
