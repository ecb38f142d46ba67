"""First-order phase boundary tables for PhaseAwareContinuitySeed.

Mirror of `PhaseBoundaryData` / `load_phase_boundary` / `interpolate_mu_c`
(src/pnjl/solver/SeedStrategies.jl:365-475 of the reference).  The shipped data files
(data/boundary.csv, data/cep.csv) are the reference's data/reference/pnjl/{boundary,cep}.csv.
"""
import csv
import math
import os
from dataclasses import dataclass, field
from typing import List

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_BOUNDARY = os.path.join(DATA_DIR, "boundary.csv")
DEFAULT_CEP = os.path.join(DATA_DIR, "cep.csv")


@dataclass
class PhaseBoundaryData:
    T_values: List[float] = field(default_factory=list)   # MeV, ascending
    mu_values: List[float] = field(default_factory=list)  # MeV
    T_CEP: float = math.nan
    mu_CEP: float = math.nan
    xi: float = 0.0

    def as_table(self):
        return (list(self.T_values), list(self.mu_values), self.T_CEP)

    @property
    def empty(self):
        return not self.T_values and math.isnan(self.T_CEP)


def _rows(path):
    if not os.path.isfile(path):
        return
    with open(path) as f:
        for parts in csv.reader(f):
            if not parts or parts[0].startswith("xi") or len(parts) < 3:
                continue
            try:
                yield float(parts[0]), parts
            except ValueError:
                continue


def load_phase_boundary(xi, boundary_path=DEFAULT_BOUNDARY, cep_path=DEFAULT_CEP):
    """SeedStrategies.jl:388-436: rows with |xi_row - xi| <= 1e-6, sorted by T; first matching CEP row."""
    T_CEP = mu_CEP = math.nan
    for x, parts in _rows(cep_path):
        if abs(x - xi) > 1e-6:
            continue
        try:
            T_CEP, mu_CEP = float(parts[1]), float(parts[2])
        except ValueError:
            T_CEP = mu_CEP = math.nan
        break
    pairs = []
    for x, parts in _rows(boundary_path):
        if abs(x - xi) > 1e-6:
            continue
        try:
            pairs.append((float(parts[1]), float(parts[2])))
        except ValueError:
            continue
    pairs.sort(key=lambda p: p[0])
    return PhaseBoundaryData([p[0] for p in pairs], [p[1] for p in pairs], T_CEP, mu_CEP, float(xi))


def interpolate_mu_c(data: PhaseBoundaryData, T_MeV):
    """SeedStrategies.jl:446-475: NaN above the CEP or without data; clamped at the table ends."""
    if not math.isnan(data.T_CEP) and T_MeV > data.T_CEP:
        return math.nan
    if not data.T_values:
        return math.nan
    Ts, ms = data.T_values, data.mu_values
    T = float(T_MeV)
    if T <= Ts[0]:
        return ms[0]
    if T >= Ts[-1]:
        return ms[-1]
    for i in range(len(Ts) - 1):
        if Ts[i] <= T <= Ts[i + 1]:
            t = (T - Ts[i]) / (Ts[i + 1] - Ts[i])
            return ms[i] + t * (ms[i + 1] - ms[i])
    return math.nan


def default_tables(xis, boundary_path=DEFAULT_BOUNDARY, cep_path=DEFAULT_CEP):
    """Tables for a set of xi values in the form Engine.set_boundaries() takes, plus xi -> table index
    (-1 when the reference has no data for that xi: plain continuity, SeedStrategies.jl:777-779)."""
    tables, index = [], {}
    for xi in xis:
        xi = float(xi)
        if xi in index:
            continue
        d = load_phase_boundary(xi, boundary_path, cep_path)
        if d.empty:
            index[xi] = -1
        else:
            index[xi] = len(tables)
            tables.append(d.as_table())
    return tables, index
