"""Host mirror of the step that follows the gap solve in scripts/relaxtime/run_gap_transport_scan.jl:297-305
(`build_K_data`): the one-loop integral A, G^f and the effective couplings K_alpha^+-.

    A(m, mu, T, Phi, Phibar, nodes, weights)                    src/relaxtime/OneLoopIntegrals.jl:531-543
    calculate_G_from_A(A_f, m_f; Nc=3)                          src/relaxtime/EffectiveCouplings.jl:56-60
    coupling_matrix_determinant(K0, K8, K08)                    EffectiveCouplings.jl:117-119
    calculate_effective_couplings(G, K, G_u, G_s)               EffectiveCouplings.jl:232-279
    build_K_data(T_fm, mu_fm, masses, Phi, Phibar)              run_gap_transport_scan.jl:297-305

The quadrature (A) runs on the GPU through `pnjl_effective_couplings_host`; the closed forms are host arithmetic like in
the Julia host code.  No CPU path for A: without the library / a B200 the calls raise PnjlError.
"""
import math
from collections import namedtuple

import numpy as np

from . import _abi
from ._lib import Engine
from .constants import DEFAULT

KCoeffs = namedtuple("KCoeffs", _abi.AUX_NAMES[4:])
KData = namedtuple("KData", ["K_coeffs", "A_vals"])

_ENGINE = None


def _engine():
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(p_num=64, t_num=8)
    return _ENGINE


def A(m, mu, T, Phi, Phibar, nodes_p=None, weights_p=None, engine=None):
    """One-loop integral A for one state (floats) or arrays of states; the rule defaults to gauleg(0, 10, 64)."""
    e = engine or _engine()
    if nodes_p is not None:
        e.set_oneloop_rule(nodes_p, weights_p)
    scalar = np.ndim(m) == 0
    aux = e.effective_couplings(np.atleast_1d(T), mu, m, m, Phi, Phibar)
    if nodes_p is not None and engine is None:
        from ._lib import gauleg
        e.set_oneloop_rule(*gauleg(0.0, 10.0, 64))
    a = aux[:, _abi.AUX["A_u"]]
    return float(a[0]) if scalar else a


def calculate_G_from_A(A_f, m_f=None, Nc=3):
    if m_f is None:
        raise ValueError("calculate_G_from_A(A_f) is deprecated; call calculate_G_from_A(A_f, m_f)")   # ArgumentError upstream
    return -Nc / (4.0 * math.pi ** 2) * (m_f * A_f)


def coupling_matrix_determinant(K0, K8, K08):
    return K0 * K8 - K08 ** 2


def calculate_effective_couplings(G, K, G_u, G_s):
    term_0 = (1.0 / 3.0) * K * (2.0 * G_u + G_s)
    term_123 = 0.5 * K * G_s
    term_4567 = 0.5 * K * G_u
    term_8 = (1.0 / 6.0) * K * (4.0 * G_u - G_s)
    term_08 = (1.0 / 6.0) * math.sqrt(2.0) * K * (G_u - G_s)
    K0p, K0m, K8p, K8m = G - term_0, G + term_0, G + term_8, G - term_8
    return KCoeffs(K0p, K0m, G + term_123, G - term_123, G + term_4567, G - term_4567, K8p, K8m, term_08, -term_08,
                   coupling_matrix_determinant(K0p, K8p, term_08), coupling_matrix_determinant(K0m, K8m, -term_08))


def build_K_data(T_fm, mu_fm, masses, Phi, Phibar, engine=None):
    """masses: (u, d, s) in fm^-1.  Returns KData(K_coeffs, A_vals=(u, d, s)) like the script's NamedTuple."""
    e = engine or _engine()
    aux = e.effective_couplings([T_fm], [mu_fm], [masses[0]], [masses[2]], [Phi], [Phibar])[0]
    return KData(KCoeffs(*aux[4:]), (aux[0], aux[0], aux[1]))


def k_coeffs_from_aux(aux_row):
    return KCoeffs(*aux_row[4:])


__all__ = ["A", "calculate_G_from_A", "coupling_matrix_determinant", "calculate_effective_couplings", "build_K_data",
           "KCoeffs", "KData", "k_coeffs_from_aux", "DEFAULT"]
