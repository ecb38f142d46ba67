"""Batched (xi, muB, T) scan — the loop of scripts/relaxtime/run_gap_transport_scan.jl:362-550 with every
line marched on the GPU at once, and its 47-column CSV (header :274-291, row :498-529).

Loop order, seeding and units follow the script: for xi → for muB → T ascending; muq = muB/3; T_fm = T/ħc;
MultiSeed until the line's tracker holds a converged solution, then PhaseAwareContinuitySeed(xi).
The relaxation-time / transport columns (tau_*, tauinv_*, eta, sigma, zeta, eta_over_s, zeta_over_s) belong to
the reference's src/relaxtime chain, which is outside the accelerated path: they are written as NaN here and
are meant to be filled by the unchanged Julia chain from the returned masses/Phi/densities.
"""
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi as A
from ._lib import Engine
from .boundary import default_tables
from .constants import DEFAULT, PNJLConstants

HEADER = [
    "T_MeV", "muq_MeV", "muB_MeV", "xi", "T_fm", "muq_fm", "converged", "iterations", "residual_norm",
    "Phi", "Phibar", "m_u", "m_d", "m_s", "rho_baryon", "rho_norm",
    "omega_fm4inv", "P_fm4inv", "epsilon_fm4inv", "s_fm3inv", "omega_MeV_fm3", "P_MeV_fm3", "epsilon_MeV_fm3",
    "eps_minus_3P_over_T4", "n_u", "n_d", "n_s", "n_ubar", "n_dbar", "n_sbar",
    "tau_u", "tau_d", "tau_s", "tau_ubar", "tau_dbar", "tau_sbar",
    "tauinv_u", "tauinv_d", "tauinv_s", "tauinv_ubar", "tauinv_dbar", "tauinv_sbar",
    "eta", "sigma", "zeta", "eta_over_s", "zeta_over_s"]
N_EQUILIBRIUM_COLS = 30


def julia_float(x):
    """`string(::Float64)` of Julia: shortest round-trip digits; positional for 1e-4 <= |x| < 1e6 (always with a
    decimal point), otherwise `d.ddde±N` with no zero padding; NaN / Inf / -Inf."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    r = repr(abs(x))
    if "e" in r:
        mant, ex = r.split("e")
        ex = int(ex)
    else:
        mant, ex = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    if ip.strip("0"):
        e10 = len(ip.lstrip("0")) - 1 + ex
    else:
        e10 = -(len(fp) - len(fp.lstrip("0")) + 1) + ex
    digits = digits.rstrip("0") or "0"
    sign = "-" if x < 0 else ""
    if -5 < e10 < 6:
        if e10 >= 0:
            ipart = digits[:e10 + 1].ljust(e10 + 1, "0")
            fpart = digits[e10 + 1:] or "0"
        else:
            ipart = "0"
            fpart = "0" * (-e10 - 1) + digits
        return "%s%s.%s" % (sign, ipart, fpart)
    return "%s%s.%se%d" % (sign, digits[0], digits[1:] or "0", e10)


def julia_range(start, stop, step):
    """collect(range(start; stop, step)) for Float64 inputs (run_gap_transport_scan.jl:400-401)."""
    n = int(math.floor((stop - start) / step + 1e-10)) + 1
    return [start + i * step for i in range(max(n, 0))]


@dataclass
class ScanOptions:
    """Subset of run_gap_transport_scan.jl's options that concern the equilibrium solve (:97-124)."""
    output: str = os.path.join("data", "outputs", "results", "relaxtime", "gap_transport_scan.csv")
    xi_values: Sequence[float] = (0.0,)
    tmin: float = 50.0
    tmax: float = 200.0
    tstep: float = 10.0
    mubmin: float = 0.0
    mubmax: float = 1200.0
    mubstep: float = 60.0
    overwrite: bool = False
    resume: bool = True
    p_num: int = 12
    t_num: int = 6
    max_iter: int = 40
    resume_mode: str = "line"      # "line": a line with missing rows is re-marched from its first T (continuity seeding as in an
                                   # uninterrupted run) and only the missing rows are appended;  "reference": like the script, the
                                   # existing rows are skipped WITHOUT being solved, so the tracker is empty at the first missing T of a
                                   # line, which is then bootstrapped with MultiSeed (run_gap_transport_scan.jl:417-443) — row for row
                                   # what the reference writes when it resumes the same file
    T_values: Optional[Sequence[float]] = None     # explicit grids override the ranges
    muB_values: Optional[Sequence[float]] = None
    metadata: dict = field(default_factory=dict)


@dataclass
class ScanGrid:
    T_MeV: np.ndarray          # [n_T] ascending
    muq_MeV: np.ndarray        # [n_lines]
    muB_MeV: np.ndarray        # [n_lines]
    xi: np.ndarray             # [n_lines]
    table_idx: np.ndarray      # [n_lines] int32
    tables: list

    @property
    def n_lines(self):
        return int(self.muq_MeV.size)

    @property
    def n_T(self):
        return int(self.T_MeV.size)


def build_grid(xi_values, muB_values, T_values) -> ScanGrid:
    """Lines in the script's order: xi outer, muB inner (:407-408); muq = muB / 3 (:426)."""
    xi_values = [float(x) for x in xi_values]
    muB = sorted(set(float(m) for m in muB_values))           # unique(sort(...)) :402
    tables, index = default_tables(xi_values)
    lx = np.repeat(np.asarray(xi_values, dtype=np.float64), len(muB))
    lm = np.tile(np.asarray(muB, dtype=np.float64), len(xi_values))
    tidx = np.array([index[x] for x in lx], dtype=np.int32)
    return ScanGrid(np.asarray(T_values, dtype=np.float64), lm / 3.0, lm, lx, tidx, tables)


def derived_columns(rec, T_MeV, muq_MeV, muB_MeV, xi, consts: PNJLConstants = DEFAULT):
    """Columns 1-30 of the CSV from result records (run_gap_transport_scan.jl:498-507).  rec: [..., 32]."""
    hb = consts.hbarc
    nq = rec[..., A.REC_NQ:A.REC_NQ + 3]
    nqb = rec[..., A.REC_NQBAR:A.REC_NQBAR + 3]
    rho_quark_net = (nq[..., 0] - nqb[..., 0]) + (nq[..., 1] - nqb[..., 1]) + (nq[..., 2] - nqb[..., 2])
    rho_baryon = rho_quark_net / 3.0
    T_fm = T_MeV / hb
    om, P, eps, s = (rec[..., A.REC_OMEGA], rec[..., A.REC_PRESSURE], rec[..., A.REC_ENERGY], rec[..., A.REC_ENTROPY])
    with np.errstate(all="ignore"):
        e3p = np.where(np.isfinite(eps) & np.isfinite(P) & np.isfinite(T_fm) & (T_fm != 0.0),
                       (eps - 3.0 * P) / T_fm ** 4, np.nan)
    st = rec[..., A.REC_STATUS].astype(np.int64)
    return {
        "T_MeV": T_MeV, "muq_MeV": muq_MeV, "muB_MeV": muB_MeV, "xi": xi, "T_fm": T_fm, "muq_fm": muq_MeV / hb,
        "converged": (st & A.ST_CONVERGED) != 0, "iterations": rec[..., A.REC_ITER].astype(np.int64),
        "residual_norm": rec[..., A.REC_RESNORM], "Phi": rec[..., 3], "Phibar": rec[..., 4],
        "m_u": rec[..., 5], "m_d": rec[..., 6], "m_s": rec[..., 7],
        "rho_baryon": rho_baryon, "rho_norm": rho_baryon / consts.rho0_fm3,
        "omega_fm4inv": om, "P_fm4inv": P, "epsilon_fm4inv": eps, "s_fm3inv": s,
        "omega_MeV_fm3": om * hb, "P_MeV_fm3": P * hb, "epsilon_MeV_fm3": eps * hb, "eps_minus_3P_over_T4": e3p,
        "n_u": nq[..., 0], "n_d": nq[..., 1], "n_s": nq[..., 2],
        "n_ubar": nqb[..., 0], "n_dbar": nqb[..., 1], "n_sbar": nqb[..., 2]}


def format_rows(cols, n_rows):
    out = []
    names = HEADER[:N_EQUILIBRIUM_COLS]
    arrs = [np.broadcast_to(np.asarray(cols[n]), (n_rows,)) for n in names]
    tail = ",".join(["NaN"] * (len(HEADER) - N_EQUILIBRIUM_COLS))
    for i in range(n_rows):
        parts = []
        for n, a in zip(names, arrs):
            v = a[i]
            if n == "converged":
                parts.append("true" if v else "false")
            elif n == "iterations":
                parts.append(str(int(v)))
            else:
                parts.append(julia_float(v))
        out.append(",".join(parts) + "," + tail)
    return out


def read_existing_keys(path, key_cols=("T_MeV", "muB_MeV", "xi")):
    """scripts/utils/scan_csv.jl:53-91."""
    keys = set()
    if not os.path.isfile(path):
        return keys
    with open(path) as f:
        header = None
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            if header is None:
                header = {c.strip(): i for i, c in enumerate(s.split(","))}
                if any(c not in header for c in key_cols):
                    return keys
                continue
            parts = s.split(",")
            try:
                keys.add(tuple(float(parts[header[c]]) for c in key_cols))
            except (ValueError, IndexError):
                continue
    return keys


def ensure_output_header_compatible(path):
    """run_gap_transport_scan.jl:248-272."""
    if not os.path.isfile(path):
        return
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            for c in ("omega_fm4inv", "P_fm4inv", "epsilon_fm4inv", "s_fm3inv", "eps_minus_3P_over_T4",
                      "eta_over_s", "zeta_over_s"):
                if c not in s:
                    raise RuntimeError("existing output CSV header is incompatible with current script (missing "
                                       "column: %s). Please rerun with --overwrite or choose a new --output path." % c)
            return


def solve_grid(grid: ScanGrid, p_num=12, t_num=6, max_iter=40, engine: Optional[Engine] = None, lines=None):
    """Run the lines of `grid` (optionally a subset `lines` of line indices) → records [n_lines][n_T][32]."""
    e = engine if engine is not None else Engine(p_num=p_num, t_num=t_num, max_iter=max_iter)
    e.set_boundaries(grid.tables)
    sel = slice(None) if lines is None else np.asarray(lines)
    return e.scan_lines(grid.muq_MeV[sel], grid.xi[sel], grid.T_MeV, grid.table_idx[sel])


def run_scan(opts: ScanOptions, engine: Optional[Engine] = None):
    """run_scan(opts) of run_gap_transport_scan.jl:362-550 for the equilibrium columns.

    Resume (existing rows are never rewritten), two modes (ScanOptions.resume_mode):
      "line" (default)  a line is recomputed from its first T when any of its points is missing from the existing file
                        (continuity seeding makes later points depend on earlier ones) and only the missing rows are appended: the
                        appended rows are those of an uninterrupted run.
      "reference"       what the script does: rows already in the file are skipped without solving, so the line's tracker holds
                        no previous solution at its first missing T — that point is bootstrapped with MultiSeed and continuity
                        resumes from there.  Near the first-order line this can land on another branch / iteration count than
                        "line" mode; it is the mode to use when the file is later compared row for row with one resumed upstream.
    Returns the number of rows written."""
    out = opts.output
    d = os.path.dirname(out)
    if d and not os.path.isdir(d):
        os.makedirs(d)
    if opts.resume and os.path.isfile(out) and not opts.overwrite:
        ensure_output_header_compatible(out)
    existing = read_existing_keys(out) if (opts.resume and os.path.isfile(out) and not opts.overwrite) else set()
    if opts.overwrite and os.path.isfile(out):
        os.remove(out)
    new_file = (not os.path.isfile(out)) or os.path.getsize(out) == 0
    T_values = list(opts.T_values) if opts.T_values is not None else julia_range(opts.tmin, opts.tmax, opts.tstep)
    muB_values = list(opts.muB_values) if opts.muB_values is not None else julia_range(opts.mubmin, opts.mubmax, opts.mubstep)
    grid = build_grid(opts.xi_values, muB_values, T_values)
    todo = [l for l in range(grid.n_lines)
            if any((float(T), float(grid.muB_MeV[l]), float(grid.xi[l])) not in existing for T in grid.T_MeV)]
    written = 0
    with open(out, "a") as io:
        if new_file:
            meta = {"schema": "scan_csv_v1", "title": "gap_transport_scan",
                    "script": "julia_relaxtime_b200.scan.run_scan (B200 equilibrium path of scripts/relaxtime/run_gap_transport_scan.jl)",
                    "p_num": str(opts.p_num), "t_num": str(opts.t_num), "max_iter": str(opts.max_iter)}
            meta.update(opts.metadata)
            for k, v in meta.items():
                io.write("# %s: %s\n" % (k, v))                 # scan_csv.jl:18-22
            io.write(",".join(HEADER) + "\n")
        if todo and opts.resume_mode == "reference":
            # per line: the missing T values in ascending order form the march (a line's tracker never sees the skipped points)
            e = engine if engine is not None else Engine(p_num=opts.p_num, t_num=opts.t_num, max_iter=opts.max_iter)
            e.set_boundaries(grid.tables)
            groups = {}
            for l in todo:
                miss = tuple(i for i, T in enumerate(grid.T_MeV)
                             if (float(T), float(grid.muB_MeV[l]), float(grid.xi[l])) not in existing)
                groups.setdefault(miss, []).append(l)
            for miss, lines in groups.items():               # lines that miss the same T values share one launch
                Tm = grid.T_MeV[list(miss)]
                sel = np.asarray(lines)
                rec = e.scan_lines(grid.muq_MeV[sel], grid.xi[sel], Tm, grid.table_idx[sel])
                for j, l in enumerate(lines):
                    cols = derived_columns(rec[j], Tm, grid.muq_MeV[l], grid.muB_MeV[l], grid.xi[l])
                    for row in format_rows(cols, len(Tm)):
                        io.write(row + "\n")
                        written += 1
                io.flush()
        elif todo:
            if opts.resume_mode != "line":
                raise ValueError("resume_mode must be 'line' or 'reference'")
            rec = solve_grid(grid, opts.p_num, opts.t_num, opts.max_iter, engine, todo)
            for j, l in enumerate(todo):
                cols = derived_columns(rec[j], grid.T_MeV, grid.muq_MeV[l], grid.muB_MeV[l], grid.xi[l])
                rows = format_rows(cols, grid.n_T)
                for T, row in zip(grid.T_MeV, rows):
                    if (float(T), float(grid.muB_MeV[l]), float(grid.xi[l])) in existing:
                        continue
                    io.write(row + "\n")
                    written += 1
                io.flush()
    return written
