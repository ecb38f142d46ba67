"""Time the bare FJ quadrature pass (k_eval_fj: loop + reduction + finish, no solver code) — development aid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from julia_relaxtime_b200._lib import Engine
n = 2_000_000
rng = np.random.default_rng(0)
T = rng.uniform(50, 300, n) / 197.327
mu = rng.uniform(0, 400, n) / 197.327
xi = rng.choice([-0.6, -0.4, -0.2, 0, 0.2, 0.4, 0.6, 0.8], n)
x = np.tile(np.array([-1.0, -1.0, -2.0, 0.3, 0.35]), (n, 1)) + rng.uniform(-0.1, 0.1, (n, 5))
x[:, 1] = x[:, 0]
for v in os.environ.get("VARIANTS", "1").split(","):
    os.environ["PNJL_LOOP_VARIANT"] = v
    e = Engine(p_num=64, t_num=16, max_iter=40)
    for _ in range(2):
        e.eval_fj(T, mu, xi, x)
    ms = e.stats()["kernel_ms"]
    inst = n * (1024 / 32) * 185.0          # FP64 warp-instructions (2 flavours x 92.5 per node)
    peak = 148 * 4 * 0.5 * 1.965e9
    print("variant %s: %d FJ passes in %.1f ms -> %.2f Mpass/s; FP64 issue utilisation ~%.1f%%" % (
        v, n, ms, n / ms / 1e3, 100 * inst / (ms * 1e-3) / peak))
