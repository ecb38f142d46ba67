"""ThermoDerivatives: the oracle (exact AD over (x, T, mu)) against the reference's committed
data/outputs/results/pnjl/{bulk_viscosity,derivatives}_xi0.0.csv (p_num=24, t_num=8; the T = 150 MeV rows — the T = 160 MeV
rows of those files sit on an unphysical root, M_u = -636 MeV, of the solver version that wrote them), and the GPU path
(pnjl_eval_derivs_host: analytic partial derivatives in (T, mu) summed over the mesh, then -J^-1 dF/dtheta) against the oracle."""
import numpy as np
import pytest

from oracle.oracle import HBARC, Oracle

# rows 1-2 of bulk_viscosity_xi0.0.csv and derivatives_xi0.0.csv (T = 150 MeV; mu = 0, 50 MeV), %.6e / %.6f as printed
GOLD_BULK = {
    "v_n_sq": (7.879846e-02, 7.951898e-02), "dmuB_dT_sigma": (None, -2.179307e+00),
    "M_u": (363.909334, 363.453557), "M_d": (363.909334, 363.453557), "M_s": (546.453381, 546.184966),
    "dM_u_dT": (-3.035097e-01, -3.475445e-01), "dM_s_dT": (-1.929164e-01, -2.197904e-01),
    "dM_u_dmuB": (None, -6.608376e-03), "dM_s_dmuB": (None, -3.890179e-03),
    "s": (1.668448e-01, 1.832041e-01), "n_B": (None, 2.278966e-03)}
GOLD_DERIV = {
    "dM_u_dmu": (None, -1.982513e-02), "dM_s_dmu": (None, -1.167054e-02), "dP_dT": (1.668448e-01, 1.832041e-01),
    "dP_dmu": (None, 6.836898e-03), "dEps_dT": (2.117361e+00, 2.415916e+00), "dEps_dmu": (None, 1.147412e-01),
    "dn_dT": (None, 4.634741e-02), "dn_dmu": (7.617229e-03, 1.190139e-02), "P": (2.161791e+01, 2.161871e+01),
    "eps": (-2.149108e+01, -2.147771e+01)}


@pytest.fixture(scope="module")
def orc():
    return Oracle(p_num=24, t_num=8, max_iter=1000)


def _oracle_at(orc, T_MeV, mu_MeV, xi=0.0):
    T, mu = np.asarray(T_MeV, dtype=float) / HBARC, np.asarray(mu_MeV, dtype=float) / HBARC
    r = orc.solve_points(T, mu, np.broadcast_to(xi, T.shape), "auto")
    assert r.converged.all()
    x = np.array([r.x[q] for q in range(5)]).T
    return x, orc.thermo_derivatives(T, mu, xi, x)


def test_oracle_reproduces_reference_derivative_tables(orc):
    _, d = _oracle_at(orc, [150.0, 150.0], [0.0, 50.0])
    for table in (GOLD_BULK, GOLD_DERIV):
        for k, gold in table.items():
            for i, g in enumerate(gold):
                if g is None:
                    assert abs(d[k][i]) < 1e-10, (k, d[k][i])       # round-off noise upstream too (1e-16 .. 1e-12)
                    continue
                v = d[k][i] * (HBARC if k in ("M_u", "M_d", "M_s") else 1.0)
                assert abs(v - g) <= 6e-7 * abs(g), (k, i, v, g)    # 7 printed digits
    assert abs(d["dmuB_dT_sigma"][0]) < 1e-10                        # upstream prints -2.6e-12 at mu = 0


def test_oracle_derivatives_against_its_own_finite_differences(orc):
    """The AD Hessian is consistent with differences of the oracle's own solve (an independent route through NLsolve)."""
    T0, mu0 = 150.0 / HBARC, 50.0 / HBARC
    _, d = _oracle_at(orc, [150.0], [50.0])
    h = 1e-4 * T0
    Ts = np.array([T0 - h, T0 + h, T0, T0])
    Ms = np.array([mu0, mu0, mu0 - h, mu0 + h])
    r = orc.solve_points(Ts, Ms, np.zeros(4), "auto")
    M = np.asarray(r.mass)          # solver masses (bare masses differ by constants only -> same derivatives)
    assert abs((M[0][1] - M[0][0]) / (2 * h) - d["dM_u_dT"][0]) < 1e-6
    assert abs((M[2][3] - M[2][2]) / (2 * h) - d["dM_s_dmu"][0]) < 1e-6
    assert abs((r.entropy[1] - r.entropy[0]) / (2 * h) - (d["dEps_dT"][0] - 3 * mu0 * d["dn_dT"][0]) / T0) < 1e-5


class _HostEngine:
    """Stands in for Engine in the host-logic test below: same methods, evaluated by the TEST-ONLY host build of the
    product's math/solver headers (tests/hostsim) — the finite-difference / implicit-differentiation algebra of
    thermo_derivatives.py is then checked on the CPU against the oracle's exact AD."""

    def __init__(self, orc):
        from julia_relaxtime_b200.constants import DEFAULT
        from tests.hostsim.hostsim import HostSim
        self.hs = HostSim(orc.p_nodes, orc.p_w, orc.c_nodes, orc.c_w, max_iter=1000)
        self.consts = DEFAULT

    def solve_points(self, T, mu, xi, seed_mode, seeds=None):
        return self.hs.solve_points(T, mu, xi, seed_mode, seeds)

    def eval_derivs(self, T, mu, xi, x):
        x = np.asarray(x, dtype=float).reshape(-1, 5)
        n = x.shape[0]
        out = dict(F=np.zeros((n, 5)), J=np.zeros((n, 5, 5)), entropy=np.zeros(n), pressure=np.zeros(n), rho=np.zeros((n, 3)),
                   rho_norm=np.zeros(n), dF_dT=np.zeros((n, 5)), dF_dmu=np.zeros((n, 5)), s_T=np.zeros(n), s_mu=np.zeros(n),
                   nB_T=np.zeros(n), nB_mu=np.zeros(n))
        for i in range(n):
            out["F"][i], out["J"][i] = self.hs.fj(x[i], T[i], mu[i], xi[i])
            th = self.hs.thermo(x[i], T[i], mu[i], xi[i])
            out["entropy"][i], out["pressure"][i], out["rho"][i], out["rho_norm"][i] = th["entropy"], th["pressure"], th["rho"], th["rho_norm"]
            d = self.hs.derivs(x[i], T[i], mu[i], xi[i])
            for k in ("dF_dT", "dF_dmu", "s_T", "s_mu", "nB_T", "nB_mu"):
                out[k][i] = d[k]
        return out


def test_host_algebra_of_thermo_derivatives_matches_oracle_on_cpu(orc):
    from julia_relaxtime_b200 import thermo_derivatives as td
    e = _HostEngine(orc)
    T_MeV = np.array([150.0, 150.0, 110.0, 220.0])
    mu_MeV = np.array([0.0, 50.0, 280.0, 120.0])
    xi = np.array([0.0, 0.0, 0.2, -0.4])
    T, mu = T_MeV / HBARC, mu_MeV / HBARC
    bulk = td.bulk_viscosity_coefficients(T, mu, xi=xi, engine=e)
    thr = td.thermo_derivatives(T, mu, xi=xi, engine=e)
    _, d = _oracle_at(orc, T_MeV, mu_MeV, xi)
    nz = mu_MeV > 0
    # every first derivative is analytic now: 1e-10 against the oracle's exact AD (2e-8 with the finite differences of round 1)
    assert np.allclose(bulk["v_n_sq"], d["v_n_sq"], rtol=1e-10, atol=0)
    assert np.allclose(bulk["dmuB_dT_sigma"][nz], d["dmuB_dT_sigma"][nz], rtol=1e-10, atol=0)
    for i, f in enumerate("uds"):
        assert np.allclose(bulk["dM_dT"][:, i], d["dM_%s_dT" % f], rtol=1e-10, atol=1e-13)
        assert np.allclose(thr["dM_dmu"][:, i], d["dM_%s_dmu" % f], rtol=1e-10, atol=1e-13)
        assert np.allclose(bulk["masses"][:, i], d["M_" + f], rtol=1e-10, atol=0)
    for a, b in (("dP_dT", "dP_dT"), ("dP_dmu", "dP_dmu"), ("dEpsilon_dT", "dEps_dT"), ("dEpsilon_dmu", "dEps_dmu"), ("dn_dT", "dn_dT"),
                 ("dn_dmu", "dn_dmu"), ("energy", "eps")):
        assert np.allclose(thr[a], d[b], rtol=1e-10, atol=1e-12), a
    assert abs(bulk["v_n_sq"][1] - 7.951898e-02) < 6e-9          # the reference's printed value
    # order = 2 (ThermoDerivatives.jl:150-172): against central differences of the ORACLE's exact first derivatives
    md = td.mass_derivatives(T[1:3], mu[1:3], order=2, xi=xi[1:3], engine=e)
    h = 2e-4 * T[1:3]
    k = 2e-4 * np.maximum(T[1:3], mu[1:3])
    _, dp = _oracle_at(orc, (T[1:3] + h) * HBARC, mu_MeV[1:3], xi[1:3])
    _, dm = _oracle_at(orc, (T[1:3] - h) * HBARC, mu_MeV[1:3], xi[1:3])
    _, ep = _oracle_at(orc, T_MeV[1:3], (mu[1:3] + k) * HBARC, xi[1:3])
    _, em = _oracle_at(orc, T_MeV[1:3], (mu[1:3] - k) * HBARC, xi[1:3])
    for i, f in enumerate("uds"):
        ref_TT = (dp["dM_%s_dT" % f] - dm["dM_%s_dT" % f]) / (2 * h)
        ref_mm = (ep["dM_%s_dmu" % f] - em["dM_%s_dmu" % f]) / (2 * k)
        ref_Tm = (dp["dM_%s_dmu" % f] - dm["dM_%s_dmu" % f]) / (2 * h)
        assert np.allclose(md["d2M_dT2"][:, i], ref_TT, rtol=2e-5, atol=1e-8), (f, md["d2M_dT2"][:, i], ref_TT)
        assert np.allclose(md["d2M_dmu2"][:, i], ref_mm, rtol=2e-5, atol=1e-8), (f, md["d2M_dmu2"][:, i], ref_mm)
        assert np.allclose(md["d2M_dTdmu"][:, i], ref_Tm, rtol=2e-5, atol=1e-8), (f, md["d2M_dTdmu"][:, i], ref_Tm)
    with pytest.raises(ValueError):
        td.mass_derivatives(T[0], mu[0], order=3, engine=e)


@pytest.mark.gpu
def test_gpu_thermo_derivatives_match_oracle_and_reference_tables(orc):
    from julia_relaxtime_b200 import thermo_derivatives as td
    from julia_relaxtime_b200._lib import Engine
    e = Engine(p_num=24, t_num=8, max_iter=1000, nodes=(orc.p_nodes, orc.p_w, orc.c_nodes, orc.c_w))
    T_MeV = np.array([150.0, 150.0, 100.0, 120.0, 200.0, 250.0, 140.0])
    mu_MeV = np.array([0.0, 50.0, 200.0, 300.0, 100.0, 20.0, 250.0])
    xi = np.array([0.0, 0.0, 0.0, 0.2, -0.4, 0.6, 0.0])
    T, mu = T_MeV / HBARC, mu_MeV / HBARC
    bulk = td.bulk_viscosity_coefficients(T, mu, xi=xi, engine=e)
    thr = td.thermo_derivatives(T, mu, xi=xi, engine=e)
    x, d = _oracle_at(orc, T_MeV, mu_MeV, xi)

    def close(a, b, rel=1e-10, floor=1e-3):
        return (np.abs(a - b) <= rel * np.abs(b) + floor * rel).all()

    nz = mu_MeV > 0
    # the two ratios divide by differences of products of derivatives: an order of magnitude of slack over the 1e-10 below
    assert close(bulk["v_n_sq"], d["v_n_sq"], 1e-9) and close(bulk["dmuB_dT_sigma"][nz], d["dmuB_dT_sigma"][nz], 1e-9)
    assert close(bulk["s"], d["s"], 1e-10) and close(bulk["n_B"], d["n_B"], 1e-9, 1e-3)
    for i, f in enumerate("uds"):
        assert close(bulk["masses"][:, i], d["M_" + f], 1e-10)
        assert close(bulk["dM_dT"][:, i], d["dM_%s_dT" % f], 1e-10, 1e-3)
        assert close(bulk["dM_dmuB"][:, i], d["dM_%s_dmuB" % f], 1e-10, 1e-3)
        assert close(thr["dM_dmu"][:, i], d["dM_%s_dmu" % f], 1e-10, 1e-3)
    for a, b in (("dP_dT", "dP_dT"), ("dP_dmu", "dP_dmu"), ("dEpsilon_dT", "dEps_dT"), ("dEpsilon_dmu", "dEps_dmu"),
                 ("dn_dT", "dn_dT"), ("dn_dmu", "dn_dmu"), ("pressure", "P"), ("energy", "eps")):
        assert close(thr[a], d[b], 1e-10, 1e-2), a
    assert thr["converged"].all()
    # the reference's own numbers, straight from the GPU path
    assert abs(bulk["v_n_sq"][1] - 7.951898e-02) < 6e-9 and abs(bulk["dmuB_dT_sigma"][1] + 2.179307) < 2e-6
    assert abs(bulk["masses"][0, 0] * HBARC - 363.909334) < 1e-6 and abs(bulk["masses"][0, 2] * HBARC - 546.453381) < 1e-6
    assert abs(thr["dEpsilon_dT"][1] - 2.415916) < 2e-6
    one = td.bulk_viscosity_coefficients(T[1], mu[1], engine=e)
    # scalar call vs the same point inside a batch: the host algebra (numpy reductions, batched 5x5 solves) groups its
    # additions differently for different batch lengths, so the two agree to round-off of the difference quotients, not bitwise
    assert np.ndim(one["v_n_sq"]) == 0 and abs(one["v_n_sq"] - bulk["v_n_sq"][1]) < 1e-12 * abs(bulk["v_n_sq"][1]) + 1e-13
    md = td.mass_derivatives(T[:2], mu[:2], engine=e)
    assert np.allclose(md["dM_dT"], bulk["dM_dT"][:2], rtol=0, atol=1e-13)
    # order = 2: central differences of the oracle's exact first derivatives as the yardstick
    m2 = td.mass_derivatives(T[1:4], mu[1:4], order=2, xi=xi[1:4], engine=e)
    h = 2e-4 * T[1:4]
    _, dp = _oracle_at(orc, (T[1:4] + h) * HBARC, mu_MeV[1:4], xi[1:4])
    _, dm = _oracle_at(orc, (T[1:4] - h) * HBARC, mu_MeV[1:4], xi[1:4])
    for i, f in enumerate("uds"):
        ref_TT = (dp["dM_%s_dT" % f] - dm["dM_%s_dT" % f]) / (2 * h)
        assert np.allclose(m2["d2M_dT2"][:, i], ref_TT, rtol=2e-5, atol=1e-8), (f, m2["d2M_dT2"][:, i], ref_TT)
    assert set(m2) == {"masses", "dM_dT", "dM_dmu", "d2M_dT2", "d2M_dTdmu", "d2M_dmu2"}
