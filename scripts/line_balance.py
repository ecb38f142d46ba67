"""How uneven are the lines?  Per-line quadrature-pass totals of the cfg5 grid (development aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200._lib import Engine
from julia_relaxtime_b200.scan import build_grid
xis = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
n_mu, n_T = 1024, int(os.environ.get("NT", "1024"))
grid = build_grid(xis, 3.0 * np.linspace(0, 400, n_mu), np.linspace(50, 300, n_T))
e = Engine(p_num=64, t_num=16, max_iter=40)
e.set_boundaries(grid.tables)
rec = e.scan_lines(grid.muq_MeV, grid.xi, grid.T_MeV, grid.table_idx)
passes = rec[..., A.REC_NEVAL] + rec[..., A.REC_NTHERMO] + rec[..., A.REC_NFUSED]
per_line = passes.sum(axis=1)
print("lines", per_line.size, "passes/line: mean %.0f min %.0f max %.0f  max/mean %.3f  p99/mean %.3f" % (
    per_line.mean(), per_line.min(), per_line.max(), per_line.max() / per_line.mean(), np.percentile(per_line, 99) / per_line.mean()))
first = passes[:, 0]
print("bootstrap passes (first T): mean %.1f max %.0f; share of all passes %.3f" % (first.mean(), first.max(), first.sum() / passes.sum()))
# lines are handed out in index order, 56 per SM (one wave): per-SM load if CTA b takes lines [56 b, 56 b + 56)
for per_sm in (56,):
    n_sm = (per_line.size + per_sm - 1) // per_sm
    load = np.array([per_line[i * per_sm:(i + 1) * per_sm].sum() for i in range(n_sm)])
    print("contiguous blocks of %d lines: SM load max/mean %.3f" % (per_sm, load.max() / load.mean()))
    perm = np.random.default_rng(0).permutation(per_line.size)
    load = np.array([per_line[perm[i * per_sm:(i + 1) * per_sm]].sum() for i in range(n_sm)])
    print("random blocks of %d lines:     SM load max/mean %.3f" % (per_sm, load.max() / load.mean()))
st = rec[..., A.REC_STATUS].astype(int)
print("TR attempted points", ((st & 4) != 0).sum(), "multiseed points", ((st & 8) != 0).sum(), "phase switch", ((st & 128) != 0).sum())
print("kernel ms", e.stats()["kernel_ms"])
# per-line cost in FP64 loop instructions (paired loops: FJ 177, fused final 265, thermo 153 per node) for offline analysis
cost = 177.0 * rec[..., A.REC_NEVAL].sum(axis=1) + 265.0 * rec[..., A.REC_NFUSED].sum(axis=1) + 153.0 * rec[..., A.REC_NTHERMO].sum(axis=1)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/line_cost.npy", np.stack([rec[..., A.REC_NEVAL].sum(axis=1), rec[..., A.REC_NFUSED].sum(axis=1), rec[..., A.REC_NTHERMO].sum(axis=1)]))
print("per-line loop cost: max/mean %.3f p90/mean %.3f p10/mean %.3f min/mean %.3f" % (
    cost.max() / cost.mean(), np.percentile(cost, 90) / cost.mean(), np.percentile(cost, 10) / cost.mean(), cost.min() / cost.mean()))
