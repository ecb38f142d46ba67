import sys, numpy as np
sys.path.insert(0, '.')
from oracle.oracle import HBARC, Oracle
from julia_relaxtime_b200 import thermo_derivatives as td
from julia_relaxtime_b200._lib import Engine
orc = Oracle(p_num=24, t_num=8, max_iter=1000)
e = Engine(p_num=24, t_num=8, max_iter=1000, nodes=(orc.p_nodes, orc.p_w, orc.c_nodes, orc.c_w))
rng = np.random.default_rng(0)
n = 200
T_MeV = rng.uniform(60, 280, n); mu_MeV = rng.uniform(10, 330, n); xi = rng.choice([-0.4, 0.0, 0.2, 0.6], n)
T, mu = T_MeV / HBARC, mu_MeV / HBARC
bulk = td.bulk_viscosity_coefficients(T, mu, xi=xi, engine=e)
thr = td.thermo_derivatives(T, mu, xi=xi, engine=e)
r = orc.solve_points(T, mu, xi, "auto")
x = np.array([r.x[q] for q in range(5)]).T
d = orc.thermo_derivatives(T, mu, xi, x)
ok = r.converged & thr["converged"]
def rel(a, b): return np.abs(a - b)[ok] / np.maximum(np.abs(b)[ok], 1e-3 * np.abs(b)[ok].max())
for a, b in (("v_n_sq", "v_n_sq"), ("dmuB_dT_sigma", "dmuB_dT_sigma")):
    print(a, "max rel err %.2e median %.2e" % (rel(bulk[a], d[b]).max(), np.median(rel(bulk[a], d[b]))))
for a, b in (("dEpsilon_dT", "dEps_dT"), ("dn_dmu", "dn_dmu"), ("dP_dT", "dP_dT")):
    print(a, "max rel err %.2e" % rel(thr[a], d[b]).max())
print("dM_u_dT max rel err %.2e" % rel(bulk["dM_dT"][:, 0], d["dM_u_dT"]).max(), "points", ok.sum())
