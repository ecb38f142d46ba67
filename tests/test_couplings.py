"""One-loop integral A and effective couplings: oracle vs independent quadrature, the product's header
(host build) and the GPU kernel vs the oracle, and the properties the reference's own tests check
(tests/unit/relaxtime/test_oneloopintegrals.jl:125-163, test_effective_couplings.jl:41-160)."""
import numpy as np
import pytest

from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200.constants import DEFAULT
from oracle.oracle import HBARC, Oracle

LAM = DEFAULT.Lambda_inv_fm


@pytest.fixture(scope="module")
def orc():
    return Oracle(p_num=12, t_num=6, max_iter=40)


def _states(n, seed=0):
    rng = np.random.default_rng(seed)
    T = rng.uniform(50, 300, n) / HBARC
    mu = rng.uniform(0, 400, n) / HBARC
    m_u = rng.uniform(0.02, 2.0, n)
    m_s = rng.uniform(0.7, 2.9, n)
    Phi = rng.uniform(0.0, 1.0, n)
    Phib = rng.uniform(0.0, 1.0, n)
    return T, mu, m_u, m_s, Phi, Phib


def test_oracle_A_against_adaptive_quadrature(orc):
    """test_oneloopintegrals.jl:125-163: C(m) is the closed form of int_0^Lambda p^2/E dp; A equals 4(-C + thermal integral)
    within rtol 5e-5 for the 64-node rule, and doubling the nodes changes it by < 5e-5."""
    from scipy.integrate import quad
    m, mu, T, Phi, Phib = 0.3, 0.25, 0.15, 0.5, 0.5      # TEST_PARAMS-like values (fm^-1)
    C_num = quad(lambda p: p * p / np.sqrt(p * p + m * m), 0.0, LAM, epsrel=1e-12)[0]

    def dist(p):
        E = np.sqrt(p * p + m * m)
        y, z = np.exp(-(E - mu) / T), np.exp(-(E + mu) / T)
        fq = (Phi * y + 2 * Phib * y**2 + y**3) / (1 + 3 * Phi * y + 3 * Phib * y**2 + y**3)
        fa = (Phib * z + 2 * Phi * z**2 + z**3) / (1 + 3 * Phib * z + 3 * Phi * z**2 + z**3)
        return p * p / E * (fq + fa)

    thermal = quad(dist, 0.0, 10.0, epsrel=1e-11, limit=200)[0]
    expected = 4.0 * (-C_num + thermal)
    a64 = orc.oneloop_A(m, mu, T, Phi, Phib)
    assert abs(a64 - expected) <= 5e-5 * abs(expected) + 1e-6
    n128, w128 = orc.gauleg(0.0, 10.0, 128)
    assert abs(orc.oneloop_A(m, mu, T, Phi, Phib, n128, w128) - a64) <= 5e-5 * abs(a64) + 1e-6
    # m -> 0 limit of the constant term (OneLoopIntegrals.jl:509-512) and negative masses (max(m, 0))
    a0 = orc.oneloop_A(0.0, 0.0, T, Phi, Phib)
    assert np.isfinite(a0) and abs(orc.oneloop_A(1e-15, 0.0, T, Phi, Phib) - a0) < 1e-12


def test_effective_coupling_properties(orc):
    """test_effective_couplings.jl:41-160 on the oracle and on the host mirror."""
    from julia_relaxtime_b200 import couplings as cp
    G, K = DEFAULT.G_fm2, DEFAULT.K_fm5
    for impl in (lambda *a: np.asarray(orc.effective_couplings(*a)), lambda *a: np.asarray(cp.calculate_effective_couplings(*a))):
        k = impl(G, 0.0, -0.3, -0.2)                     # K = 0: every K_alpha = G, mixing 0
        assert np.abs(k[:8] - G).max() < 1e-15 and np.abs(k[8:10]).max() < 1e-15
        k = impl(G, K, 0.0, 0.0)                         # G^f = 0
        assert np.abs(k[:8] - G).max() < 1e-15 and np.abs(k[8:10]).max() < 1e-15
        k = impl(G, K, -0.25, -0.25)                     # G^u = G^s: flavour degenerate
        assert abs(k[2] - k[4]) < 1e-15 and abs(k[3] - k[5]) < 1e-15 and abs(k[2] - k[6]) < 1e-15 and abs(k[3] - k[7]) < 1e-15
        assert np.abs(k[8:10]).max() < 1e-15
        k = impl(G, K, -0.3, -0.2)                       # magnitudes (:198-219)
        assert (k[10:] > 0).all() and (np.abs(k[:8] - G) / G < 0.5).all() and (np.abs(k[8:10]) < 0.1 * G).all()
        assert abs(k[10] - (k[0] * k[6] - k[8] ** 2)) < 1e-20 and abs(k[11] - (k[1] * k[7] - k[9] ** 2)) < 1e-20
    assert cp.calculate_G_from_A(2.0, 1.5) == pytest.approx(-3 / (4 * np.pi**2) * 3.0, abs=1e-15)
    with pytest.raises(ValueError):
        cp.calculate_G_from_A(2.0)
    assert np.allclose(orc.effective_couplings(G, K, -0.3, -0.2), cp.calculate_effective_couplings(G, K, -0.3, -0.2), rtol=0, atol=1e-17)


def test_product_header_couplings_match_oracle(orc):
    """The product's oneloop_A / effective_couplings (scale-free occupation numbers) against the literal restatement."""
    from tests.hostsim.hostsim import HostSim
    hs = HostSim(orc.p_nodes, orc.p_w, orc.c_nodes, orc.c_w, max_iter=40)
    st = _states(400, seed=3)
    nodes, weights = orc.gauleg(0.0, 10.0, 64)
    a = hs.couplings(*st, nodes, weights)
    b = orc.couplings_batch(*st)
    assert np.abs(a - b).max() / np.abs(b).max() < 1e-13
    assert (np.abs(a - b) <= 2e-13 * np.abs(b) + 1e-15).all()


@pytest.mark.gpu
def test_gpu_couplings_match_oracle(orc):
    from julia_relaxtime_b200._lib import Engine
    e = Engine(p_num=12, t_num=6, max_iter=40)
    st = _states(5000, seed=1)
    aux = e.effective_couplings(*st)
    ref = orc.couplings_batch(*st)
    assert aux.shape == (5000, A.AUX_DOUBLES)
    assert (np.abs(aux - ref) <= 1e-12 * np.abs(ref) + 1e-15).all()
    assert e.stats()["kernel_launches"] == 1
    # edge cases: m = 0 and negative mass (const-term branch), empty batch, another rule
    z = e.effective_couplings([0.7], [0.1], [0.0], [-0.5], [0.3], [0.4])
    zr = orc.couplings_batch([0.7], [0.1], [0.0], [-0.5], [0.3], [0.4])
    assert np.allclose(z, zr, rtol=1e-12, atol=1e-15)
    assert e.effective_couplings(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0)).shape == (0, 16)
    n20, w20 = orc.gauleg(0.0, 20.0, 16)              # the rule the reference's docstring recommends as an alternative
    e.set_oneloop_rule(n20, w20)
    a20 = e.effective_couplings(*[s[:64] for s in st])
    r20 = orc.couplings_batch(*[s[:64] for s in st], n20, w20)
    assert (np.abs(a20 - r20) <= 1e-12 * np.abs(r20) + 1e-15).all()
    with pytest.raises(Exception):
        e.set_oneloop_rule(np.zeros(0), np.zeros(0))


@pytest.mark.gpu
def test_gpu_scan_with_couplings_equals_oracle_from_records(orc):
    """pnjl_scan_lines_couplings_host: same records as the plain scan, aux = build_K_data of every record."""
    from julia_relaxtime_b200._lib import Engine
    from julia_relaxtime_b200.boundary import default_tables
    from julia_relaxtime_b200.couplings import build_K_data
    tables, index = default_tables([0.0, 0.2])
    e = Engine(p_num=12, t_num=6, max_iter=40, nodes=(orc.p_nodes, orc.p_w, orc.c_nodes, orc.c_w))
    e.set_boundaries(tables)
    T = np.linspace(60.0, 260.0, 21)
    muq = np.array([0.0, 150.0, 330.0, 300.0])
    xi = np.array([0.0, 0.0, 0.0, 0.2])
    tidx = np.array([index[x] for x in xi], dtype=np.int32)
    rec0 = e.scan_lines(muq, xi, T, tidx)
    rec, aux = e.scan_lines_couplings(muq, xi, T, tidx)
    assert np.array_equal(rec0, rec)
    r = rec.reshape(-1, A.REC_DOUBLES)
    ref = orc.couplings_batch(r[:, A.REC_T], r[:, A.REC_MU], r[:, A.REC_MASS], r[:, A.REC_MASS + 2], r[:, A.REC_X + 3],
                              r[:, A.REC_X + 4])
    assert (np.abs(aux.reshape(-1, 16) - ref) <= 1e-12 * np.abs(ref) + 1e-15).all()
    kd = build_K_data(r[5, A.REC_T], r[5, A.REC_MU], r[5, A.REC_MASS:A.REC_MASS + 3], r[5, A.REC_X + 3], r[5, A.REC_X + 4], engine=e)
    assert np.allclose(np.asarray(kd.K_coeffs), ref[5, 4:], rtol=1e-12) and kd.A_vals[0] == kd.A_vals[1]
    assert (ref[:, 14:] > 0).all()                       # det K > 0 along physical solutions (test_effective_couplings.jl:192-193)
