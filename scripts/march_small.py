"""Small exercise of the line-march kernel for compute-sanitizer runs (development aid): one warp per line with time slices,
teams of 2, 3, 4 and 8 warps (leader + followers, named barriers, command blocks and partial sums in shared memory), mixed
isotropic / anisotropic lines, records stored through an output index; the warp-specialised kernel on the same lines."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A  # noqa: E402
from julia_relaxtime_b200._lib import Engine  # noqa: E402
from julia_relaxtime_b200.boundary import default_tables  # noqa: E402

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
tables, index = default_tables([0.0, 0.2])
T = np.linspace(120.0, 260.0, n_T)
muq = np.linspace(0.0, 400.0, n_lines)
xi = np.tile([0.0, 0.2, -0.4, 0.6], (n_lines + 3) // 4)[:n_lines]
tidx = np.array([index.get(x, -1) for x in xi], dtype=np.int32)
ref = None
CASES = ((2, 0, 0), (3, 1, 2), (3, 2, 0), (3, 3, 0), (3, 4, 3), (3, 8, 0))
if len(sys.argv) > 3 and sys.argv[3] == "solo":       # no team barriers: what compute-sanitizer's synccheck can follow (see profiles/r02_sanitizer.txt)
    CASES = CASES[:2]
for schedule, parts, quantum in CASES:
    e = Engine(p_num=64, t_num=16, max_iter=40, schedule=schedule)
    e.set_boundaries(tables)
    if parts:
        e.set_option("march_parts", parts)
    if quantum:
        e.set_option("march_quantum", quantum)
    rec = e.scan_lines(muq, xi, T, tidx)
    conv = int(((rec[..., A.REC_STATUS].astype(int) & 1) != 0).sum())
    if ref is None:
        ref = rec
    err = float(np.nanmax(np.abs(rec[..., :5] - ref[..., :5]) / np.maximum(1e-300, np.abs(ref[..., :5]))))
    print("schedule %d parts %d quantum %d: converged %d of %d, worst rel. difference to the first run %.2e, threads %d lanes/solve %d" % (
        schedule, parts, quantum, conv, rec.shape[0] * rec.shape[1], err, e.stats()["threads"], e.stats()["lanes_per_solve"]), flush=True)
