"""Time k_couplings (one-loop A + effective couplings) on cfg5-sized input — development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200._lib import Engine
n = 8 * 1024 * 1024
rng = np.random.default_rng(0)
T = rng.uniform(50, 300, n) / 197.327
mu = rng.uniform(0, 400, n) / 197.327
m_u = rng.uniform(0.02, 2.0, n); m_s = rng.uniform(0.7, 2.9, n)
P = rng.uniform(0, 1, n); Pb = rng.uniform(0, 1, n)
e = Engine(p_num=64, t_num=16, max_iter=40)
for _ in range(3):
    t0 = time.time(); aux = e.effective_couplings(T, mu, m_u, m_s, P, Pb); wall = time.time() - t0
ms = e.stats()["kernel_ms"]
print("k_couplings: %d points, 64 nodes x 2 flavours: kernel %.2f ms (%.0f M points/s), host call %.1f ms; finite %s" % (
    n, ms, n / ms / 1e3, wall * 1e3, np.isfinite(aux).all()))
