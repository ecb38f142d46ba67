"""Quadrature passes per point on a cfg5 slab (development aid): python scripts/pass_counts.py  [PNJL_LIB=... in the environment]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200._lib import Engine
from julia_relaxtime_b200.scan import build_grid
xis = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
grid = build_grid(xis, 3.0 * np.linspace(0, 400, 128), np.linspace(50, 300, 1024))
e = Engine(p_num=64, t_num=16, max_iter=40)
e.set_boundaries(grid.tables)
rec = e.scan_lines(grid.muq_MeV, grid.xi, grid.T_MeV, grid.table_idx)
n = rec.shape[0] * rec.shape[1]
fj, ft, th = rec[..., A.REC_NEVAL].sum() / n, rec[..., A.REC_NFUSED].sum() / n, rec[..., A.REC_NTHERMO].sum() / n
print("fj %.4f ft %.4f th %.4f total %.4f  work %.1f (169 fj + 140 ft + 110 th); wasted fused %.4f, missed %.4f" % (
    fj, ft, th, fj + ft + th, 169 * fj + 140 * ft + 110 * th, ft - (1 - th), th))
