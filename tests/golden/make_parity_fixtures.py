#!/usr/bin/env python
"""Generate the oracle-side fixtures of the full-resolution parity tests (tests/test_gpu_parity_configs.py).

    python tests/golden/make_parity_fixtures.py [cfg3] [cfg4] [cfg5]      (about 20 minutes on 8 cores for all three)

The AD oracle (oracle/pnjl_oracle.cpp: the reference's algorithm with nested-dual Jacobians) is too slow to run inside a
GPU test at these sizes (≈ 250 points/s on 8 cores), so its results on stratified samples of BASELINE configs 3, 4 and 5 —
at the configs' true T / mu resolution and 64x16 nodes — are computed once here and committed as compressed .npz files:

  parity_cfg5.npz  128 complete lines (every xi x 16 mu values, half of them inside the 280-360 MeV first-order / crossover
                   band) x all 1024 T of the 1024x1024x8 grid
  parity_cfg4.npz  32 complete mu-lines x all 2048 T of the 2048x2048 CEP window (T 100-160, mu_q 260-330 MeV, xi = 0)
  parity_cfg3.npz  2048 points of the 256x256x8 grid (fixed RNG seed), MultiSeed at every point

Stored per point: the state x (5), Omega, residual norm, iterations, status bits and the oracle's evaluation count
(n_fj: how long the Newton path was — the tests use it to recognise far-from-root wanders).  Masses follow from x in closed
form (Thermodynamics.jl:81-88) and are recomputed by the test.  Inputs (which lines / points) are stored alongside so the
test feeds the GPU exactly the same sample.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from julia_relaxtime_b200.scan import build_grid           # noqa: E402  (pure host logic, no GPU)
from oracle.oracle import HBARC, Oracle                     # noqa: E402

XI8 = [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8]
MAX_ITER = 40


def pack(res):
    return dict(x=res.x.T.copy(), omega=res.omega.copy(), residual_norm=res.residual_norm.copy(),
                iterations=res.iterations.astype(np.int16), status=res.status.astype(np.int32),
                n_fj=res.n_fj.astype(np.int32))


def cfg5_sample():
    """Line indices (into build_grid(XI8, 3 mus, T), xi-major) of the stratified sample: per xi 8 mu values spread over
    0..400 MeV and 8 inside 280..360 MeV, with a different offset for every xi."""
    mus = np.linspace(0.0, 400.0, 1024)
    lines = []
    band = np.nonzero((mus >= 280.0) & (mus <= 360.0))[0]
    for ix in range(len(XI8)):
        wide = (np.arange(8) * 128 + 16 * ix + 5) % 1024
        inband = band[(np.arange(8) * (len(band) // 8) + 3 * ix + 1) % len(band)]
        for im in sorted(set(wide.tolist() + inband.tolist())):
            lines.append(ix * 1024 + im)
    return np.array(lines, dtype=np.int64)


def make_cfg5(o):
    mus = np.linspace(0.0, 400.0, 1024)
    T = np.linspace(50.0, 300.0, 1024)
    grid = build_grid(XI8, 3.0 * mus, T)
    sel = cfg5_sample()
    t0 = time.time()
    res = o.scan_lines(grid.muq_MeV[sel], grid.xi[sel], T, grid.tables, grid.table_idx[sel])
    print("cfg5: %d lines x %d T in %.0f s" % (len(sel), len(T), time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "parity_cfg5.npz"), lines=sel, **pack(res))


def cfg4_sample():
    return (np.arange(32) * 64 + 17).astype(np.int64)          # 32 of the 2048 mu lines, evenly spread


def make_cfg4(o):
    mus = np.linspace(260.0, 330.0, 2048)
    T = np.linspace(100.0, 160.0, 2048)
    grid = build_grid([0.0], 3.0 * mus, T)
    sel = cfg4_sample()
    t0 = time.time()
    res = o.scan_lines(grid.muq_MeV[sel], grid.xi[sel], T, grid.tables, grid.table_idx[sel])
    print("cfg4: %d lines x %d T in %.0f s" % (len(sel), len(T), time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "parity_cfg4.npz"), lines=sel, **pack(res))


def make_cfg3(o):
    mus = np.linspace(0.0, 400.0, 256)
    T = np.linspace(50.0, 300.0, 256)
    rng = np.random.default_rng(20261018)
    n = 2048
    ix, im, it = rng.integers(0, 8, n), rng.integers(0, 256, n), rng.integers(0, 256, n)
    t0 = time.time()
    res = o.solve_points(T[it] / HBARC, mus[im] / HBARC, np.asarray(XI8)[ix], "multi")
    print("cfg3: %d points in %.0f s" % (n, time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "parity_cfg3.npz"), ix=ix, im=im, it=it, **pack(res))


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    o = Oracle(p_num=64, t_num=16, max_iter=MAX_ITER, n_threads=int(os.environ.get("ORACLE_THREADS", "0")))
    for w in which:
        {"cfg3": make_cfg3, "cfg4": make_cfg4, "cfg5": make_cfg5}[w](o)
