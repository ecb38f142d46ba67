"""Quick exploratory timings of the solve kernels (development aid; bench.py is the contract)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200 import _abi as A  # noqa: E402
from julia_relaxtime_b200._lib import Engine  # noqa: E402
from julia_relaxtime_b200.boundary import default_tables  # noqa: E402


def flops_alg(rec, n_nodes):
    nev = rec[..., A.REC_NEVAL].sum()
    nth = rec[..., A.REC_NTHERMO].sum()
    return n_nodes * 3 * (123.0 * nev + 54.0 * nth), nev, nth


def run_lines(p, t, n_mu, n_T, xis, lanes=0, reps=2):
    e = Engine(p_num=p, t_num=t, max_iter=40, lanes_per_solve=lanes)
    tables, index = default_tables(xis)
    e.set_boundaries(tables)
    T = np.linspace(50, 300, n_T)
    mus = np.linspace(0, 400, n_mu)
    muq = np.tile(mus, len(xis))
    xi = np.repeat(np.array(xis, dtype=float), n_mu)
    tidx = np.array([index[x] for x in xi], dtype=np.int32)
    for _ in range(reps):
        t0 = time.time()
        rec = e.scan_lines(muq, xi, T, tidx)
        wall = time.time() - t0
    st = e.stats()
    npts = rec.shape[0] * rec.shape[1]
    fl, nev, nth = flops_alg(rec, p * t)
    conv = ((rec[..., A.REC_STATUS].astype(int) & 1) != 0).mean()
    ms = st["kernel_ms"]
    print("lines %dx%d mesh, %d lines x %d T = %d pts: kernel %.1f ms (wall %.1f ms) -> %.3f Mpts/s; evals/pt %.2f thermo/pt %.2f; "
          "alg %.2f TFLOP/s; conv %.4f; G=%d blocks=%d regs=%d" % (
              p, t, len(muq), n_T, npts, ms, wall * 1e3, npts / ms / 1e3, nev / npts, nth / npts, fl / ms / 1e9, conv,
              st["lanes_per_solve"], st["blocks"], st["regs_per_thread"]))
    return rec


def run_points(p, t, n, mode=A.SEED_MULTI, lanes=0):
    e = Engine(p_num=p, t_num=t, max_iter=40, lanes_per_solve=lanes)
    rng = np.random.default_rng(0)
    T = rng.uniform(50, 300, n) / 197.327
    mu = rng.uniform(0, 400, n) / 197.327
    xi = rng.choice([-0.6, -0.4, -0.2, 0, 0.2, 0.4, 0.6, 0.8], n)
    for _ in range(2):
        rec = e.solve_points(T, mu, xi, mode)
    st = e.stats()
    fl, nev, nth = flops_alg(rec, p * t)
    ms = st["kernel_ms"]
    conv = ((rec[..., A.REC_STATUS].astype(int) & 1) != 0).mean()
    print("points %dx%d mesh, n=%d mode=%d: kernel %.1f ms -> %.3f Mpts/s; evals/pt %.2f; alg %.2f TFLOP/s; conv %.4f" % (
        p, t, n, mode, ms, n / ms / 1e3, nev / n, fl / ms / 1e9, conv))


if __name__ == "__main__":
    e = Engine(p_num=12, t_num=6)
    print("fp64 peak TFLOP/s (burst, sustained):", e.measure_fp64_peak())
    del e
    run_lines(12, 6, 128, 128, [0.0])
    run_lines(64, 16, 256, 64, [0.0, 0.2])
    run_lines(64, 16, 1024, 64, [-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6, 0.8])
    run_lines(64, 16, 128, 64, [0.0])
    run_points(64, 16, 65536)
    run_points(12, 6, 65536)
