"""Helpers to read the reference's golden scan CSV (tests/golden/, see make_fixtures.py)."""
import csv
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_scan_csv(path):
    """Returns dict column -> np.array (bools for true/false columns)."""
    with open(path) as f:
        rd = csv.reader(line for line in f if not line.startswith("#") and line.strip())
        hdr = next(rd)
        rows = list(rd)
    cols = {}
    for j, h in enumerate(hdr):
        vals = [r[j] for r in rows]
        if vals and vals[0] in ("true", "false"):
            cols[h] = np.array([v == "true" for v in vals])
        else:
            try:
                cols[h] = np.array([float(v) for v in vals])
            except ValueError:
                cols[h] = np.array(vals)
    return cols


def golden_lines(cols):
    """Group the golden rows into (xi, muB) lines in file order → list of (xi, muB, row indices)."""
    lines = []
    for i, (x, m) in enumerate(zip(cols["xi"], cols["muB_MeV"])):
        if not lines or (lines[-1][0], lines[-1][1]) != (x, m):
            lines.append((x, m, []))
        lines[-1][2].append(i)
    return lines


def rel(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
