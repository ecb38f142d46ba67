"""Host mirror of src/pnjl/derivatives/ThermoDerivatives.jl:

    mass_derivatives(T_fm, mu_fm; order=1|2, xi, p_num, t_num)               :132-175
    thermo_derivatives(T_fm, mu_fm; xi, p_num, t_num)                        :186-250
    bulk_derivative_coeffs(T_fm, mu_fm; ...)                                 :252-262
    bulk_viscosity_coefficients(T_fm, mu_fm; xi, p_num, t_num)               :342-467

The reference differentiates through the solve with ImplicitDifferentiation.jl (dx/dtheta = -J^-1 dF/dtheta, :80-109) and
ForwardDiff on calculate_thermo / calculate_rho.  Here every quadrature runs on the GPU and every first derivative is
ANALYTIC: one `solve` batch for the states, then one `pnjl_eval_derivs_host` call at the solutions, whose derivative pass
sums the closed-form partial derivatives of the integrand in (T, mu) at fixed x over the same mesh (csrc/pnjl_math.cuh,
dtheta_node):
    J = dF/dx,   dF/dT, dF/dmu,   ds/dT|x, ds/dmu|x, dn_B/dT|x, dn_B/dmu|x,
    ds/dx = dF/dT and dn_B/dx = (1/3) dF/dmu  (symmetry of the mixed partials of P),
and the rest is the reference's algebra (5x5 solves on the host).  Agreement with exact AD (the oracle): <= 1e-10 relative
(tests/test_thermo_derivatives.py).  The second derivatives of mass_derivatives(order = 2) are fourth-order central
differences of those analytic first derivatives along T and mu (re-solved at the shifted points, continuity-seeded):
~1e-8 relative.  All functions take scalars or arrays (batched over points); xi may be an array too.
"""
import numpy as np

from . import _abi as A
from ._lib import Engine

HBARC = 197.327
_ENGINES = {}


def _engine(p_num, t_num):
    key = (int(p_num), int(t_num))
    if key not in _ENGINES:
        _ENGINES[key] = Engine(p_num=p_num, t_num=t_num)          # NLsolve default iterations, like solve() upstream
    return _ENGINES[key]


def compute_masses_from_state(x, consts):
    """ThermoDerivatives.jl:111-121 — note ITS bare masses 5.5/197.327 and 140.0/197.327 (not the configured 140.7)."""
    x = np.asarray(x, dtype=float)
    m_u0, m_s0 = 0.0055 / 0.197327, 0.140 / 0.197327
    G, K = consts.G_fm2, consts.K_fm5
    return np.stack([m_u0 - 4 * G * x[..., 0] + 2 * K * x[..., 1] * x[..., 2],
                     m_u0 - 4 * G * x[..., 1] + 2 * K * x[..., 0] * x[..., 2],
                     m_s0 - 4 * G * x[..., 2] + 2 * K * x[..., 0] * x[..., 1]], axis=-1)


def implicit_derivatives(engine, T_fm, mu_fm, xi, x):
    """Everything the functions below need at states x [n, 5] (normally the converged solutions): one batched GPU call."""
    x = np.ascontiguousarray(x, dtype=float).reshape(-1, 5)
    n = x.shape[0]
    T, mu, xi = A.as_f64(T_fm, n), A.as_f64(mu_fm, n), A.as_f64(xi, n)
    st = engine.eval_derivs(T, mu, xi, x)
    J = st["J"]
    dx_dT = np.linalg.solve(J, -st["dF_dT"][..., None])[..., 0]
    dx_dmu = np.linalg.solve(J, -st["dF_dmu"][..., None])[..., 0]
    return dict(x=x, T=T, mu=mu, F=st["F"], J=J, P=st["pressure"], s=st["entropy"], rho_sum=st["rho"].sum(axis=-1),
                rho_norm=st["rho_norm"], dF_dT=st["dF_dT"], dF_dmu=st["dF_dmu"], dx_dT=dx_dT, dx_dmu=dx_dmu,
                s_T=st["s_T"], s_mu=st["s_mu"], n_T=st["nB_T"], n_mu=st["nB_mu"])


def _solve_states(engine, T, mu, xi):
    rec = engine.solve_points(T, mu, xi, A.SEED_AUTO)            # solve(FixedMu(), T, mu; xi, p_num, t_num): DefaultSeed()
    return rec


def _prepare(T_fm, mu_fm, xi, p_num, t_num, engine):
    scalar = np.ndim(T_fm) == 0 and np.ndim(mu_fm) == 0
    T = A.as_f64(T_fm)
    n = max(T.size, np.size(mu_fm))
    T, mu, xi = A.as_f64(T, n), A.as_f64(mu_fm, n), A.as_f64(xi, n)
    e = engine or _engine(p_num, t_num)
    rec = _solve_states(e, T, mu, xi)
    d = implicit_derivatives(e, T, mu, xi, rec[:, A.REC_X:A.REC_X + 5])
    d["rec"] = rec
    d["consts"] = e.consts
    return scalar, d


def _dM(d):
    x, c = d["x"], d["consts"]
    G4, K2 = -4.0 * c.G_fm2, 2.0 * c.K_fm5
    n = x.shape[0]
    dM_dx = np.zeros((n, 3, 5))
    dM_dx[:, 0, 0] = G4; dM_dx[:, 0, 1] = K2 * x[:, 2]; dM_dx[:, 0, 2] = K2 * x[:, 1]
    dM_dx[:, 1, 0] = K2 * x[:, 2]; dM_dx[:, 1, 1] = G4; dM_dx[:, 1, 2] = K2 * x[:, 0]
    dM_dx[:, 2, 0] = K2 * x[:, 1]; dM_dx[:, 2, 1] = K2 * x[:, 0]; dM_dx[:, 2, 2] = G4
    return (np.einsum("nij,nj->ni", dM_dx, d["dx_dT"]), np.einsum("nij,nj->ni", dM_dx, d["dx_dmu"]))


def _out(scalar, res):
    if not scalar:
        return res
    return {k: (v[0] if isinstance(v, np.ndarray) else v) for k, v in res.items()}


def _first_order_at(e, T, mu, xi, seeds):
    """dM/dT, dM/dmu [n, 3] at (T, mu) with the solve started from `seeds` (continuity: neighbours of a solved point)."""
    rec = e.solve_points(T, mu, xi, A.SEED_EXPLICIT, seeds.reshape(-1, 1, 5))
    d = implicit_derivatives(e, T, mu, xi, rec[:, A.REC_X:A.REC_X + 5])
    d["consts"] = e.consts
    return _dM(d)


def mass_derivatives(T_fm, mu_fm, order=1, xi=0.0, p_num=64, t_num=8, engine=None, rel_step=1e-3):
    """order = 1: masses, dM_dT, dM_dmu.  order = 2 (ThermoDerivatives.jl:150-172) adds d2M_dT2, d2M_dTdmu, d2M_dmu2:
    fourth-order central differences of the analytic first derivatives at (T +- h, +- 2h) and (mu +- k, +- 2k), each of the
    eight neighbours re-solved from the centre's solution."""
    if order not in (1, 2):
        raise ValueError("order must be 1 or 2, got %r" % (order,))          # the reference: error("order must be 1 or 2, ...")
    scalar, d = _prepare(T_fm, mu_fm, xi, p_num, t_num, engine)
    dM_dT, dM_dmu = _dM(d)
    res = dict(masses=compute_masses_from_state(d["x"], d["consts"]), dM_dT=dM_dT, dM_dmu=dM_dmu)
    if order == 2:
        e = engine or _engine(p_num, t_num)
        T, mu, x = d["T"], d["mu"], d["x"]
        n = T.size
        xi_a = A.as_f64(xi, n)
        hT = rel_step * T
        hM = rel_step * np.maximum(T, np.abs(mu))
        offs = (-2.0, -1.0, 1.0, 2.0)
        Ts = np.concatenate([T + o * hT for o in offs] + [T for _ in offs])
        Ms = np.concatenate([mu for _ in offs] + [mu + o * hM for o in offs])
        gT, gM = _first_order_at(e, Ts, Ms, np.tile(xi_a, 8), np.tile(x, (8, 1)))
        gT, gM = gT.reshape(8, n, 3), gM.reshape(8, n, 3)

        def d4(f, h):         # f at offsets (-2, -1, +1, +2) h
            return (f[0] - 8.0 * f[1] + 8.0 * f[2] - f[3]) / (12.0 * h[:, None])
        res["d2M_dT2"] = d4(gT[:4], hT)
        res["d2M_dTdmu"] = 0.5 * (d4(gM[:4], hT) + d4(gT[4:], hM))       # both orders of differentiation, averaged
        res["d2M_dmu2"] = d4(gM[4:], hM)
    return _out(scalar, res)


def _totals(d):
    ds_dx, dn_dx = d["dF_dT"], d["dF_dmu"] / 3.0
    dot = lambda a, b: np.einsum("ni,ni->n", a, b)
    return dict(ds_dT=d["s_T"] + dot(ds_dx, d["dx_dT"]), ds_dmu=d["s_mu"] + dot(ds_dx, d["dx_dmu"]),
                dn_dT=d["n_T"] + dot(dn_dx, d["dx_dT"]), dn_dmu=d["n_mu"] + dot(dn_dx, d["dx_dmu"]),
                P_T=d["s"] + dot(d["F"], d["dx_dT"]), P_mu=d["rho_sum"] + dot(d["F"], d["dx_dmu"]))


def thermo_derivatives(T_fm, mu_fm, xi=0.0, p_num=64, t_num=8, engine=None):
    scalar, d = _prepare(T_fm, mu_fm, xi, p_num, t_num, engine)
    t = _totals(d)
    T, mu, s, P = d["T"], d["mu"], d["s"], d["P"]
    eps = -P + mu * d["rho_sum"] + T * s
    E_T = -t["P_T"] + mu * 3.0 * t["dn_dT"] + s + T * t["ds_dT"]
    E_mu = -t["P_mu"] + d["rho_sum"] + mu * 3.0 * t["dn_dmu"] + T * t["ds_dmu"]
    den_e = E_T * t["dn_dmu"] - E_mu * t["dn_dT"]
    den_n = t["dn_dT"] * E_mu - t["dn_dmu"] * E_T
    with np.errstate(divide="ignore", invalid="ignore"):
        dP_de = np.where(den_e == 0, np.nan, (t["P_T"] * t["dn_dmu"] - t["P_mu"] * t["dn_dT"]) / den_e)
        dP_dn = np.where(den_n == 0, np.nan, (t["P_T"] * E_mu - t["P_mu"] * E_T) / den_n)
    dM_dT, dM_dmu = _dM(d)
    rec = d["rec"]
    return _out(scalar, dict(pressure=P, energy=eps, rho=d["rho_sum"] / 3.0, rho_norm=rec[:, A.REC_RHO_NORM], entropy=s,
                             dP_dT=t["P_T"], dP_dmu=t["P_mu"], dEpsilon_dT=E_T, dEpsilon_dmu=E_mu, dn_dT=t["dn_dT"],
                             dn_dmu=t["dn_dmu"], dP_depsilon_n=dP_de, dP_dn_epsilon=dP_dn,
                             masses=compute_masses_from_state(d["x"], d["consts"]), dM_dT=dM_dT, dM_dmu=dM_dmu,
                             converged=(rec[:, A.REC_STATUS].astype(int) & A.ST_CONVERGED) != 0,
                             iterations=rec[:, A.REC_ITER].astype(int), residual_norm=rec[:, A.REC_RESNORM]))


def bulk_derivative_coeffs(T_fm, mu_fm, **kw):
    r = thermo_derivatives(T_fm, mu_fm, **kw)
    return {k: r[k] for k in ("dP_depsilon_n", "dP_dn_epsilon", "dM_dT", "dM_dmu")}


def bulk_viscosity_coefficients(T_fm, mu_fm, xi=0.0, p_num=64, t_num=8, engine=None):
    scalar, d = _prepare(T_fm, mu_fm, xi, p_num, t_num, engine)
    t = _totals(d)
    T, s = d["T"], d["s"]
    n_B = d["rho_sum"] / 3.0
    dM_dT, dM_dmu = _dM(d)
    ds_dmuB, dn_dmuB = t["ds_dmu"] / 3.0, t["dn_dmu"] / 3.0
    with np.errstate(divide="ignore", invalid="ignore"):
        v_n_sq = (s * dn_dmuB - n_B * t["dn_dT"]) / (T * (t["ds_dT"] * dn_dmuB - ds_dmuB * t["dn_dT"]))
        dmuB_dT = -(n_B * t["ds_dT"] - s * t["dn_dT"]) / (n_B * ds_dmuB - s * dn_dmuB)
    return _out(scalar, dict(v_n_sq=v_n_sq, dmuB_dT_sigma=dmuB_dT, masses=compute_masses_from_state(d["x"], d["consts"]),
                             dM_dT=dM_dT, dM_dmuB=dM_dmu / 3.0, s=s, n_B=n_B))
