"""Regenerate tests/golden/*.csv from the reference's committed outputs.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_fixtures.py

The reference cannot be executed here (no Julia toolchain), so the golden vectors are the
reference's own committed scan outputs, trimmed to the columns this path produces:

* gap_transport_scan_xi-0p6to0p6.csv  columns 1-30 (T_MeV … n_sbar) of
  /root/reference/data/outputs/results/relaxtime/gap_transport_scan_xi-0p6to0p6.csv
  (406 rows: 7 xi × 2 muB × 29 T; p_num=12, t_num=6, iterations=40 — SURVEY.md §8c)
* tmu_scan.csv, dual_branch_T50.csv — branch-level checks (older TmuScan, %.6f / %.10f)
* boundary.csv, cep.csv — inputs of PhaseAwareContinuitySeed (SeedStrategies.jl:388-436); the same two
  files are shipped as package data in julia_relaxtime_b200/data/.
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DATA = os.path.join(HERE, "..", "..", "julia_relaxtime_b200", "data")


def trim_scan(src, dst, ncols=30):
    with open(src) as f, open(dst, "w") as g:
        g.write("# trimmed copy (columns 1-%d) of %s\n" % (ncols, os.path.relpath(src, REF)))
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            g.write(",".join(line.rstrip("\n").split(",")[:ncols]) + "\n")


def main():
    trim_scan(os.path.join(REF, "data/outputs/results/relaxtime/gap_transport_scan_xi-0p6to0p6.csv"),
              os.path.join(HERE, "gap_transport_scan_xi-0p6to0p6.csv"))
    for name in ("tmu_scan.csv", "dual_branch_T50.csv"):
        shutil.copyfile(os.path.join(REF, "data/outputs/results/pnjl", name), os.path.join(HERE, name))
    os.makedirs(PKG_DATA, exist_ok=True)
    for name in ("boundary.csv", "cep.csv"):
        shutil.copyfile(os.path.join(REF, "data/reference/pnjl", name), os.path.join(HERE, name))
        shutil.copyfile(os.path.join(REF, "data/reference/pnjl", name), os.path.join(PKG_DATA, name))


if __name__ == "__main__":
    main()
