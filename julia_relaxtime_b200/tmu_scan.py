"""T-mu scan with the semantics and CSV of `TmuScan.run_tmu_scan` (src/pnjl/scans/TmuScan.jl:120-234).

Loop order of the reference: for xi → for T → mu in the given order, one fresh PhaseAwareContinuitySeed tracker per
(xi, T) line (:166-172), the four-candidate seed list (:269-300), the 1e-4 acceptance / refine / force-promote rules
(:327-458).  Here every (xi, T) line is marched on the GPU at once (`pnjl_tmu_scan_host`), the host only formats the
19-column `%.6f` CSV (:62-82, :460-504) and handles resume keys (:236-266).

Units as in the reference: T and mu in MeV on this surface (mu is the quark chemical potential), masses written in
MeV (`masses .* ħc`), everything else in fm units.
"""
import os
from typing import Optional, Sequence

import numpy as np

from . import _abi as A
from ._lib import Engine
from .boundary import default_tables
from .constants import DEFAULT, PNJLConstants

DEFAULT_T_VALUES = [50.0 + 10.0 * i for i in range(16)]       # collect(50.0:10.0:200.0)   TmuScan.jl:56
DEFAULT_MU_VALUES = [0.0 + 10.0 * i for i in range(41)]       # collect(0.0:10.0:400.0)    TmuScan.jl:57
DEFAULT_OUTPUT_PATH = os.path.join("data", "outputs", "results", "pnjl", "tmu_scan.csv")
ACCEPTABLE_RESIDUAL = 1e-4                                    # TmuScan.jl:60

HEADER = ["T_MeV", "mu_MeV", "xi", "pressure_fm4", "rho", "entropy_fm3", "energy_fm4", "phi_u", "phi_d", "phi_s",
          "Phi1", "Phi2", "M_u_MeV", "M_d_MeV", "M_s_MeV", "iterations", "residual_norm", "converged", "message"]

CANDIDATE_LABELS = ("phase_aware", "continuation", "default_1", "default_2")


def _fmt(x):
    """@sprintf("%.6f", x) (TmuScan.jl:510); Julia prints NaN / Inf / -Inf for non-finite values."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    return "%.6f" % x


def _key(T, mu, xi):
    return (round(float(T), 6), round(float(mu), 6), round(float(xi), 6))       # TmuScan.jl:240


def load_completed(path):
    """_load_completed (TmuScan.jl:243-266): keys of the rows already in the file; malformed lines are ignored."""
    done = set()
    with open(path) as f:
        first = True
        for line in f:
            if first:
                first = False
                continue
            if not line.strip():
                continue
            cols = line.split(",")
            if len(cols) < 3:
                continue
            try:
                done.add(_key(float(cols[0]), float(cols[1]), float(cols[2])))
            except ValueError:
                continue
    return done


def message_of(status, residual):
    """The reference's free-text message column carries the refine / promote notes (TmuScan.jl:340-347, :457); the
    per-candidate failure texts of earlier candidates are reduced here to the index of the candidate that succeeded."""
    st = int(status)
    if st & A.ST_NO_RESULT:
        return "all seed candidates failed"
    parts = []
    ci = (st & A.ST_CAND_MASK) >> A.ST_CAND_SHIFT
    if ci:
        parts.append("succeeded with seed[%s]" % CANDIDATE_LABELS[ci])
    if st & A.ST_PROMOTED:
        parts.append("force-marked converged (residual %s)" % _fmt(residual))
    if st & A.ST_REFINED:
        parts.append("refined from near-converged seed")
    return " | ".join(parts)


def format_row(T, mu, xi, rec, consts: PNJLConstants = DEFAULT):
    """_write_row (TmuScan.jl:460-504) from one 32-double record."""
    st = int(rec[A.REC_STATUS])
    msg = message_of(st, rec[A.REC_RESNORM])
    q = '"%s"' % msg if msg else ""
    if st & A.ST_NO_RESULT:
        vals = [_fmt(T), _fmt(mu), _fmt(xi)] + ["NaN"] * 12 + ["-1", "NaN", "false", q]
        return ",".join(vals)
    vals = [_fmt(T), _fmt(mu), _fmt(xi), _fmt(rec[A.REC_PRESSURE]), _fmt(rec[A.REC_RHO_NORM]), _fmt(rec[A.REC_ENTROPY]),
            _fmt(rec[A.REC_ENERGY]), _fmt(rec[0]), _fmt(rec[1]), _fmt(rec[2]), _fmt(rec[3]), _fmt(rec[4]),
            _fmt(rec[A.REC_MASS] * consts.hbarc), _fmt(rec[A.REC_MASS + 1] * consts.hbarc),
            _fmt(rec[A.REC_MASS + 2] * consts.hbarc), str(int(rec[A.REC_ITER])), _fmt(rec[A.REC_RESNORM]),
            "true" if st & A.ST_CONVERGED else "false", q]
    return ",".join(vals)


def is_success(rec):
    """_is_success (TmuScan.jl:411-419) on a record: converged (after refine / promote) rows."""
    return (int(rec[A.REC_STATUS]) & A.ST_CONVERGED) != 0


def solve_tmu_grid(T_values, mu_values, xi_values, p_num=24, t_num=8, max_iter=1000, use_phase_aware=True,
                   engine: Optional[Engine] = None):
    """All (xi, T) lines at once → (records [n_xi][n_T][n_mu][32], line xi, line T)."""
    xi_values = [float(x) for x in xi_values]
    T_values = np.asarray(T_values, dtype=np.float64)
    mu_values = np.asarray(mu_values, dtype=np.float64)
    e = engine if engine is not None else Engine(p_num=p_num, t_num=t_num, max_iter=max_iter)
    if use_phase_aware:
        tables, index = default_tables(xi_values)       # PhaseAwareContinuitySeed(xi): missing data → no table
    else:
        tables, index = [], {x: -1 for x in xi_values}
    e.set_boundaries(tables)
    lx = np.repeat(np.asarray(xi_values, dtype=np.float64), T_values.size)
    lT = np.tile(T_values, len(xi_values))
    tidx = np.array([index[x] for x in lx], dtype=np.int32)
    rec = e.tmu_scan(lT, lx, mu_values, tidx)
    return rec.reshape(len(xi_values), T_values.size, mu_values.size, A.REC_DOUBLES), lx, lT


def run_tmu_scan(T_values: Sequence[float] = DEFAULT_T_VALUES, mu_values: Sequence[float] = DEFAULT_MU_VALUES,
                 xi_values: Sequence[float] = (0.0,), output_path: str = DEFAULT_OUTPUT_PATH, overwrite: bool = False,
                 resume: bool = True, use_phase_aware: bool = True, p_num: int = 24, t_num: int = 8, progress_cb=None,
                 iterations: int = 1000, engine: Optional[Engine] = None):
    """run_tmu_scan(; T_values, mu_values, xi_values, output_path, overwrite, resume, use_phase_aware, p_num, t_num,
    progress_cb, nlsolve_kwargs...) → dict(total, success, failure, skipped, output)   (TmuScan.jl:120-234).

    Resume: rows whose (T, mu, xi) key is already in the file are skipped when writing (as in the reference), but
    because continuity seeding makes later points of a line depend on earlier ones the lines that still miss rows are
    re-marched from their first mu — the rows that get appended are the ones a from-scratch run would write."""
    d = os.path.dirname(output_path)
    if d:
        os.makedirs(d, exist_ok=True)
    have_file = os.path.isfile(output_path)
    completed = load_completed(output_path) if (resume and not overwrite and have_file) else set()
    mode = "w" if (overwrite or not have_file) else "a"
    stats = dict(total=0, success=0, failure=0, skipped=0)
    xi_values = [float(x) for x in xi_values]
    T_values = [float(t) for t in T_values]
    mu_values = [float(m) for m in mu_values]
    todo = any(_key(T, mu, xi) not in completed for xi in xi_values for T in T_values for mu in mu_values)
    rec = None
    if todo:
        rec, _, _ = solve_tmu_grid(T_values, mu_values, xi_values, p_num, t_num, iterations, use_phase_aware, engine)
    with open(output_path, mode) as io:
        if mode == "w":
            io.write(",".join(HEADER) + "\n")
        for ix, xi in enumerate(xi_values):
            for iT, T in enumerate(T_values):
                for im, mu in enumerate(mu_values):
                    stats["total"] += 1
                    key = _key(T, mu, xi)
                    if key in completed:
                        stats["skipped"] += 1
                        continue
                    r = rec[ix, iT, im]
                    io.write(format_row(T, mu, xi, r) + "\n")
                    completed.add(key)
                    if is_success(r):
                        stats["success"] += 1
                    else:
                        stats["failure"] += 1
                    if progress_cb is not None:
                        try:
                            progress_cb(dict(T=T, mu=mu, xi=xi), r)
                        except Exception:      # the reference ignores callback errors (TmuScan.jl:219-223)
                            pass
                io.flush()
    stats["output"] = output_path
    return stats
