import sys, os; sys.path.insert(0,'.')
import numpy as np
from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200._lib import Engine
from julia_relaxtime_b200.scan import build_grid
xis = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
mus = np.linspace(0.0, 400.0, 1024)[::8]
T = np.linspace(50.0, 300.0, 1024)
grid = build_grid(xis, 3.0 * mus, T)
e = Engine(p_num=64, t_num=16, max_iter=40)
e.set_boundaries(grid.tables)
rec = e.scan_lines(grid.muq_MeV, grid.xi, grid.T_MeV, grid.table_idx)
bad = np.argwhere(~(rec[..., 7] > rec[..., 5]))
print(len(bad), 'bad points; lines', np.unique(bad[:,0])[:20])
for l in np.unique(bad[:,0])[:4]:
    idx = bad[bad[:,0]==l][:,1]
    print('line', l, 'xi', grid.xi[l], 'muq', grid.muq_MeV[l], 'T idx range', idx.min(), idx.max(), 'T', T[idx.min()], T[idx.max()])
    for it in (idx.min()-1, idx.min(), idx.min()+1, idx.max(), min(idx.max()+1,1023)):
        r = rec[l, it]
        print('   T', T[it], 'x', r[:5], 'M', r[5:8], 'omega', r[8], 'it', r[20], 'st', int(r[21]), 'res', r[19])
