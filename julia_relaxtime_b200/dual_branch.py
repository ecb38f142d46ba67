"""Host mirror of src/pnjl/scans/DualBranchScan.jl: the two continuity branches of every (T, xi) line are marched on the
GPU (`pnjl_dual_branch_host`, one task per branch), everything after that is host arithmetic like in the reference:

    run_dual_branch_scan(T_mev, mu_range; xi, p_num=24, t_num=8)      DualBranchScan.jl:104-182
    find_phase_transition(result) -> PhaseTransitionInfo              :190-236   (Omega crossing by linear interpolation :450-487)
    merge_branches(result; output_path) -> rows / 18-column CSV       :243-300, :490-520
    _select_physical_branch, _is_same_solution                        :380-418

Units as upstream: MeV for T and mu, fm^-1 inside the records.
"""
import math
import os
from collections import namedtuple

import numpy as np

from . import _abi as A
from ._lib import Engine

HBARC = 197.327

BranchPoint = namedtuple("BranchPoint", ["mu_mev", "converged", "omega", "pressure", "rho_norm", "entropy", "energy",
                                         "x_state", "masses", "iterations", "residual_norm"])
DualBranchResult = namedtuple("DualBranchResult", ["T_mev", "xi", "mu_values", "hadron_branch", "quark_branch",
                                                   "physical_branch"])
PhaseTransitionInfo = namedtuple("PhaseTransitionInfo", ["found", "mu_c", "omega_at_transition", "mu_hadron_spinodal",
                                                         "mu_quark_spinodal", "coexistence_region"])


def _branch_points(rec, mu_values):
    """records [n_mu][32] of one branch -> list of BranchPoint / None."""
    out = []
    for i, mu in enumerate(mu_values):
        r = rec[i]
        st = int(r[A.REC_STATUS])
        if st & A.ST_NO_RESULT:
            out.append(None)
            continue
        out.append(BranchPoint(float(mu), bool(st & A.ST_CONVERGED), r[A.REC_OMEGA], r[A.REC_PRESSURE], r[A.REC_RHO_NORM],
                               r[A.REC_ENTROPY], r[A.REC_ENERGY], r[A.REC_X:A.REC_X + 5].copy(),
                               r[A.REC_MASS:A.REC_MASS + 3].copy(), int(r[A.REC_ITER]), r[A.REC_RESNORM]))
    return out


def is_same_solution(h, q, tol=0.01):
    return (abs(h.x_state[0] - q.x_state[0]) < tol and abs(h.x_state[3] - q.x_state[3]) < tol
            and abs(h.masses[0] - q.masses[0]) < tol * 197.327)


def select_physical_branch(hadron, quark):
    phys = []
    for h, q in zip(hadron, quark):
        if h is None and q is None:
            phys.append(None)
        elif h is None:
            phys.append(q)
        elif q is None:
            phys.append(h)
        elif is_same_solution(h, q):
            phys.append(h)
        else:
            phys.append(h if h.omega <= q.omega else q)
    return phys


def results_from_records(records, T_mev, xi, mu_values):
    """records [n_lines][2][n_mu][32] -> list of DualBranchResult."""
    mu_values = [float(m) for m in mu_values]
    out = []
    for l in range(records.shape[0]):
        h = _branch_points(records[l, 0], mu_values)
        q = _branch_points(records[l, 1], mu_values)
        out.append(DualBranchResult(float(np.atleast_1d(T_mev)[l]), float(np.atleast_1d(xi)[l]), mu_values, h, q,
                                    select_physical_branch(h, q)))
    return out


def run_dual_branch_scans(T_mev, mu_range, xi=0.0, p_num=24, t_num=8, iterations=1000, engine=None):
    """Batched form: one DualBranchResult per entry of T_mev (xi scalar or per line)."""
    T_mev = A.as_f64(T_mev)
    xi = A.as_f64(xi, T_mev.size)
    mu = A.as_f64(list(mu_range))
    e = engine or Engine(p_num=p_num, t_num=t_num, max_iter=iterations)
    rec = e.dual_branch(T_mev, xi, mu)
    return results_from_records(rec, T_mev, xi, mu)


def run_dual_branch_scan(T_mev, mu_range, xi=0.0, p_num=24, t_num=8, iterations=1000, engine=None):
    return run_dual_branch_scans([T_mev], mu_range, [xi], p_num, t_num, iterations, engine)[0]


def _find_branch_endpoint(branch, mu_values, forward):
    idx = range(len(branch) - 1, -1, -1) if forward else range(len(branch))
    for i in idx:
        if branch[i] is not None:
            return mu_values[i]
    return math.nan


def _find_omega_crossing(mu_vals, om_h, om_q):
    delta = [a - b for a, b in zip(om_h, om_q)]
    sig = []
    for i in range(len(mu_vals)):
        thr = max(1e-6 * 0.5 * (abs(om_h[i]) + abs(om_q[i])), 1e-10)
        if abs(delta[i]) > thr:
            sig.append(i)
    if len(sig) < 2:
        return math.nan, math.nan
    for i1, i2 in zip(sig[:-1], sig[1:]):
        if delta[i1] * delta[i2] < 0:
            t = delta[i1] / (delta[i1] - delta[i2])
            return mu_vals[i1] + t * (mu_vals[i2] - mu_vals[i1]), om_h[i1] + t * (om_h[i2] - om_h[i1])
    return math.nan, math.nan


def find_phase_transition(result):
    mu, h, q = result.mu_values, result.hadron_branch, result.quark_branch
    mu_h = _find_branch_endpoint(h, mu, True)
    mu_q = _find_branch_endpoint(q, mu, False)
    co = [i for i in range(len(mu)) if h[i] is not None and q[i] is not None and not is_same_solution(h[i], q[i])]
    if not co:
        return PhaseTransitionInfo(False, math.nan, math.nan, mu_h, mu_q, (math.nan, math.nan))
    mu_c, om_c = _find_omega_crossing([mu[i] for i in co], [h[i].omega for i in co], [q[i].omega for i in co])
    return PhaseTransitionInfo(not math.isnan(mu_c), mu_c, om_c, mu_h, mu_q, (mu[co[0]], mu[co[-1]]))


MERGED_HEADER = ["T_MeV", "mu_MeV", "xi", "branch", "omega", "pressure", "rho", "entropy", "energy", "phi_u", "phi_d", "phi_s",
                 "Phi1", "Phi2", "M_u_MeV", "M_d_MeV", "M_s_MeV", "delta_omega"]


def merge_branches(result, output_path=None):
    rows = []
    for i, mu in enumerate(result.mu_values):
        pt, h, q = result.physical_branch[i], result.hadron_branch[i], result.quark_branch[i]
        if pt is None:
            continue
        branch = "unknown"
        if h is not None and q is not None:
            if is_same_solution(h, q):
                branch = "same"
            elif abs(pt.omega - h.omega) < 1e-10:
                branch = "hadron"
            else:
                branch = "quark"
        elif h is not None:
            branch = "hadron"
        elif q is not None:
            branch = "quark"
        d_om = math.nan
        if h is not None and q is not None and not is_same_solution(h, q):
            d_om = h.omega - q.omega
        rows.append(dict(T_MeV=result.T_mev, mu_MeV=mu, xi=result.xi, branch=branch, omega=pt.omega, pressure=pt.pressure,
                         rho=pt.rho_norm, entropy=pt.entropy, energy=pt.energy, phi_u=pt.x_state[0], phi_d=pt.x_state[1],
                         phi_s=pt.x_state[2], Phi1=pt.x_state[3], Phi2=pt.x_state[4], M_u_MeV=pt.masses[0] * HBARC,
                         M_d_MeV=pt.masses[1] * HBARC, M_s_MeV=pt.masses[2] * HBARC, delta_omega=d_om))
    if output_path is not None:
        write_merged_csv(output_path, rows)
    return rows


def _fmt(v, nd):
    return "NaN" if (isinstance(v, float) and math.isnan(v)) else ("%." + str(nd) + "f") % v


def write_merged_csv(path, rows):
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "w") as f:
        f.write(",".join(MERGED_HEADER) + "\n")
        for r in rows:
            vals = [_fmt(r["T_MeV"], 6), _fmt(r["mu_MeV"], 6), _fmt(r["xi"], 6), r["branch"]]
            vals += [_fmt(float(r[k]), 10) for k in MERGED_HEADER[4:14]]
            vals += [_fmt(float(r[k]), 6) for k in MERGED_HEADER[14:17]]
            vals.append(_fmt(float(r["delta_omega"]), 10))
            f.write(",".join(vals) + "\n")
