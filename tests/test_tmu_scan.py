"""TmuScan.run_tmu_scan semantics (src/pnjl/scans/TmuScan.jl:120-504).

CPU: the oracle's restatement against the reference's committed data/outputs/results/pnjl/tmu_scan.csv (branch level:
the file was written by an older TmuScan and is printed with %.6f), the product's solver headers (test-only host build)
against the oracle, and the CSV writer / resume logic.  GPU: `pnjl_tmu_scan_host` against the oracle and the written file.
"""
import os

import numpy as np
import pytest

from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200 import tmu_scan
from oracle.oracle import Oracle, load_phase_tables
from tests.golden_io import GOLDEN, read_scan_csv

HBARC = 197.327


def _tables(xis):
    return load_phase_tables(os.path.join(GOLDEN, "boundary.csv"), os.path.join(GOLDEN, "cep.csv"), xis)


def _records_from_oracle(res, n):
    rec = np.zeros((n, A.REC_DOUBLES))
    rec[:, 0:5] = res.x.T
    rec[:, 5:8] = res.mass.T
    rec[:, A.REC_OMEGA], rec[:, A.REC_PRESSURE], rec[:, A.REC_ENTROPY], rec[:, A.REC_ENERGY] = (
        res.omega, res.pressure, res.entropy, res.energy)
    rec[:, A.REC_RHO_NORM] = res.rho_norm
    rec[:, A.REC_RESNORM], rec[:, A.REC_ITER], rec[:, A.REC_STATUS] = res.residual_norm, res.iterations, res.status
    return rec


def test_oracle_tmu_scan_vs_reference_csv_branch_level():
    """The reference's tmu_scan.csv (656 rows, xi = 0, T 50..200, mu 0..400, p=24, t=8) was produced by an older
    TmuScan (its messages mention a removed seed cache) and carries 6 decimals, so it pins branches, not iterates:
    every row away from the metastable strip next to the first-order line and away from the T = 170 MeV line must
    agree to the printed precision.  (On the T = 170 line today's candidate order starts from the quark seed, whose
    Newton run ends on the phi_u ≈ -5.28 root that passes the reference's physicality filter — DESIGN.md §2.)"""
    g = read_scan_csv(os.path.join(GOLDEN, "tmu_scan.csv"))
    T, mu = np.unique(g["T_MeV"]), np.unique(g["mu_MeV"])
    assert T.size * mu.size == g["T_MeV"].size == 656
    tables, index = _tables([0.0])
    o = Oracle(p_num=24, t_num=8, max_iter=1000)
    r = o.tmu_scan(T, 0.0, mu, tables, np.full(T.size, index[0.0], dtype=np.int32))
    assert ((r.status & A.ST_CONVERGED) != 0).all() and not (r.status & A.ST_NO_RESULT).any()
    assert (r.residual_norm <= 1e-9).all()
    worst = np.zeros(656)
    for arr, col in ((r.x[0], "phi_u"), (r.x[1], "phi_d"), (r.x[2], "phi_s"), (r.x[3], "Phi1"), (r.x[4], "Phi2"),
                     (r.pressure, "pressure_fm4"), (r.rho_norm, "rho"), (r.entropy, "entropy_fm3"),
                     (r.energy, "energy_fm4")):
        worst = np.maximum(worst, np.abs(arr - g[col]))
    same = worst <= 2e-6
    assert same.sum() >= 600
    off = ~same
    near_first_order = (g["T_MeV"] <= 125.0) & (g["mu_MeV"] >= 300.0)
    assert (near_first_order | (g["T_MeV"] == 170.0))[off].all()
    assert same[(g["T_MeV"] != 170.0) & ~near_first_order].all()


@pytest.fixture(scope="module")
def sim_and_oracle():
    from tests.hostsim.hostsim import HostSim
    o = Oracle(p_num=12, t_num=6, max_iter=40)
    return HostSim(o.p_nodes, o.p_w, o.c_nodes, o.c_w, max_iter=40), o


def _compare_records(rec, res, min_same_iter=0.98, max_wander=2):
    """Status word, state, masses, Omega and iteration counts.  Points whose ORACLE path is a long far-from-root Newton
    wander (> 25 quadrature passes; such paths amplify last-ulp differences) may differ within the Newton
    tolerance (1e-7) on at most `max_wander` points; everything else must agree to 1e-9."""
    st = rec[:, A.REC_STATUS].astype(np.int64) & ~A.ST_MASS_INVERSION          # product-only flag, not an oracle status
    assert (st == res.status).mean() >= 0.995, np.nonzero(st != res.status)[0][:10]
    live = (res.status & A.ST_NO_RESULT) == 0
    assert (((st & A.ST_NO_RESULT) == 0) == live).all()
    worst = np.zeros(rec.shape[0])
    for q in range(5):
        scale = np.maximum(np.abs(res.x[q][live]), 1e-3 if q >= 3 else 1e-300)
        worst[live] = np.maximum(worst[live], np.abs(rec[live, q] - res.x[q][live]) / scale)
    for q in range(3):
        worst[live] = np.maximum(worst[live], np.abs(rec[live, 5 + q] - res.mass[q][live]) / np.abs(res.mass[q][live]))
    wander = res.n_fj > 25
    assert (worst[~wander] <= 1e-9).all(), (worst[~wander].max(), np.nonzero((worst > 1e-9) & ~wander)[0][:10])
    assert (worst[wander] <= 1e-7).all() and (worst[wander] > 1e-9).sum() <= max_wander
    assert np.abs(rec[live, A.REC_OMEGA] - res.omega[live]).max() <= 1e-9
    assert (rec[live, A.REC_ITER].astype(int) == res.iterations[live]).mean() >= min_same_iter
    assert np.isnan(rec[~live, 0:13]).all()


def test_product_headers_tmu_line_matches_oracle(sim_and_oracle):
    """scan_tmu_line of csrc/pnjl_solver.cuh (host build) vs the oracle: lines below, at and above the CEP for xi with
    and without a boundary table, marching mu up (the default) and down."""
    hs, o = sim_and_oracle
    tables, index = _tables([0.0, 0.4, -0.3])
    T = np.array([50.0, 100.0, 125.0, 131.0, 150.0, 170.0, 200.0, 90.0, 140.0, 60.0, 180.0])
    xi = np.array([0.0] * 7 + [0.4, 0.4, -0.3, -0.3])
    tidx = np.array([index[x] for x in xi], dtype=np.int32)
    mu = np.arange(0.0, 401.0, 12.5)
    for grid in (mu, mu[::-1].copy()):
        res = o.tmu_scan(T, xi, grid, tables, tidx)
        rec = hs.tmu_scan(T, xi, grid, tables, tidx).reshape(-1, A.REC_DOUBLES)
        _compare_records(rec, res)
        assert ((res.status & A.ST_CONVERGED) != 0).mean() > 0.99


def test_tmu_scan_with_crippled_solver_exercises_promote_and_failure_rows(sim_and_oracle):
    """With a single Newton iteration and no fallbacks most candidates stop short: rows then come out force-promoted
    (residual <= 1e-4), refined, from a later candidate, or as all-NaN rows — identically in oracle and product."""
    from tests.hostsim.hostsim import HostSim
    _, o0 = sim_and_oracle
    o = Oracle(p_num=12, t_num=6, max_iter=1, tr_fallback=False, auto_multiseed_fallback=False)
    hs = HostSim(o0.p_nodes, o0.p_w, o0.c_nodes, o0.c_w, max_iter=1, tr_fallback=False, auto_multiseed_fallback=False)
    tables, index = _tables([0.0])
    T = np.array([60.0, 120.0, 160.0, 220.0])
    mu = np.arange(0.0, 401.0, 2.0)
    tidx = np.full(4, index[0.0], dtype=np.int32)
    res = o.tmu_scan(T, 0.0, mu, tables, tidx)
    rec = hs.tmu_scan(T, np.zeros(4), mu, tables, tidx).reshape(-1, A.REC_DOUBLES)
    st = res.status
    assert (st & A.ST_PROMOTED).any() and (st & A.ST_REFINED).any() and (st & A.ST_NO_RESULT).any()
    assert ((st >> A.ST_CAND_SHIFT) & 3).max() >= 1
    _compare_records(rec, res, min_same_iter=1.0)
    assert (rec[:, A.REC_STATUS].astype(np.int64) == st).all()


def test_tmu_csv_rows_and_resume(tmp_path):
    """_write_row / _fmt / _key / _load_completed (TmuScan.jl:236-266, 460-512)."""
    rec = np.zeros(A.REC_DOUBLES)
    rec[0:5] = [-1.843295, -1.843295, -2.2270094, 2.0e-5, 2.0e-5]
    rec[A.REC_MASS:A.REC_MASS + 3] = [1.8, 1.8, 2.7]
    rec[A.REC_PRESSURE], rec[A.REC_RHO_NORM], rec[A.REC_ENTROPY], rec[A.REC_ENERGY] = 21.6080123, 0.0, 1e-7, -21.6080123
    rec[A.REC_ITER], rec[A.REC_RESNORM], rec[A.REC_STATUS] = 2, 3.1e-11, A.ST_CONVERGED
    row = tmu_scan.format_row(50.0, 0.0, 0.0, rec)
    assert row == ("50.000000,0.000000,0.000000,21.608012,0.000000,0.000000,-21.608012,-1.843295,-1.843295,-2.227009,"
                   "0.000020,0.000020,355.188600,355.188600,532.782900,2,0.000000,true,")
    assert len(row.split(",")) == len(tmu_scan.HEADER) == 19
    # same text as the reference's file on its first row (the 16-column older layout has no mass columns)
    first = open(os.path.join(GOLDEN, "tmu_scan.csv")).read().splitlines()[1].split(",")
    ours = row.split(",")
    assert ours[:12] == first[:12] and ours[15:18] == first[12:15]
    rec[A.REC_STATUS] = A.ST_CONVERGED | A.ST_PROMOTED | (2 << A.ST_CAND_SHIFT)
    rec[A.REC_RESNORM] = 5e-5
    assert tmu_scan.format_row(50.0, 0.0, 0.0, rec).endswith(
        ',0.000050,true,"succeeded with seed[default_1] | force-marked converged (residual 0.000050)"')
    rec[A.REC_STATUS] = A.ST_NO_RESULT
    nan_row = tmu_scan.format_row(50.0, 10.0, 0.2, rec).split(",")
    assert nan_row[:3] == ["50.000000", "10.000000", "0.200000"] and nan_row[3:15] == ["NaN"] * 12
    assert nan_row[15:18] == ["-1", "NaN", "false"]
    p = tmp_path / "t.csv"
    p.write_text(",".join(tmu_scan.HEADER) + "\n" + row + "\n\nbroken,line\nx,y,z\n" + ",".join(nan_row) + "\n")
    assert tmu_scan.load_completed(str(p)) == {(50.0, 0.0, 0.0), (50.0, 10.0, 0.2)}
    assert tmu_scan.DEFAULT_T_VALUES[0] == 50.0 and tmu_scan.DEFAULT_T_VALUES[-1] == 200.0
    assert tmu_scan.DEFAULT_MU_VALUES[-1] == 400.0 and len(tmu_scan.DEFAULT_MU_VALUES) == 41


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("p_num,t_num,schedule", [(24, 8, 0), (12, 6, 0), (64, 16, 0), (64, 16, 1)])
def test_gpu_tmu_scan_matches_oracle(p_num, t_num, schedule):
    """pnjl_tmu_scan_host vs the oracle for the reference's default grid (T 50..200 step 10, mu 0..400 step 10) at
    xi = 0 and 0.4, in every kernel organisation (8/16-lane groups, warp-specialised, one warp per line)."""
    from julia_relaxtime_b200._lib import Engine
    tables, index = _tables([0.0, 0.4])
    o = Oracle(p_num=p_num, t_num=t_num, max_iter=1000)
    e = Engine(p_num=p_num, t_num=t_num, max_iter=1000, schedule=schedule, nodes=(o.p_nodes, o.p_w, o.c_nodes, o.c_w))
    e.set_boundaries(tables)
    n_T = 16 if p_num < 64 else 6
    T = np.tile(np.linspace(50.0, 200.0, n_T), 2)
    xi = np.repeat([0.0, 0.4], n_T)
    tidx = np.array([index[x] for x in xi], dtype=np.int32)
    mu = np.asarray(tmu_scan.DEFAULT_MU_VALUES)
    rec = e.tmu_scan(T, xi, mu, tidx).reshape(-1, A.REC_DOUBLES)
    res = o.tmu_scan(T, xi, mu, tables, tidx)
    _compare_records(rec, res)
    assert e.stats()["kernel_launches"] == 1


@pytest.mark.gpu
def test_gpu_run_tmu_scan_writes_reference_layout_and_resumes(tmp_path):
    """run_tmu_scan end to end: header, row count, loop order xi → T → mu, text equal to rows formatted from the
    oracle's results, agreement with the reference's committed file where today's algorithm stays on its branch, and
    resume (nothing recomputed or rewritten for complete files; missing rows appended)."""
    out = str(tmp_path / "pnjl" / "tmu_scan.csv")
    stats = tmu_scan.run_tmu_scan(output_path=out)
    assert stats == dict(total=656, success=656, failure=0, skipped=0, output=out)
    lines = open(out).read().splitlines()
    assert lines[0] == ",".join(tmu_scan.HEADER) and len(lines) == 657
    tables, index = _tables([0.0])
    o = Oracle(p_num=24, t_num=8, max_iter=1000)
    T, mu = np.asarray(tmu_scan.DEFAULT_T_VALUES), np.asarray(tmu_scan.DEFAULT_MU_VALUES)
    res = o.tmu_scan(T, 0.0, mu, tables, np.full(T.size, index[0.0], dtype=np.int32))
    orec = _records_from_oracle(res, 656)
    n_text_equal = 0
    for i, line in enumerate(lines[1:]):
        want = tmu_scan.format_row(T[i // mu.size], mu[i % mu.size], 0.0, orec[i]).split(",")
        got = line.split(",")
        assert got[:3] == want[:3] and got[17] == want[17]
        for a, b in zip(got[3:15], want[3:15]):
            assert abs(float(a) - float(b)) <= 1.5e-6
        n_text_equal += got[:18] == want[:18]
    assert n_text_equal >= 640        # %.6f rounding can flip the last digit on a few rows
    g = read_scan_csv(os.path.join(GOLDEN, "tmu_scan.csv"))
    mine = read_scan_csv(out)
    agree = np.abs(mine["phi_u"] - g["phi_u"]) <= 2e-6
    assert agree.sum() >= 600
    # resume: complete file → everything skipped; drop the last 50 rows → exactly those are appended again
    stats2 = tmu_scan.run_tmu_scan(output_path=out)
    assert stats2["skipped"] == 656 and stats2["success"] == 0 and open(out).read().splitlines() == lines
    open(out, "w").write("\n".join(lines[:-50]) + "\n")
    stats3 = tmu_scan.run_tmu_scan(output_path=out)
    assert stats3["skipped"] == 606 and stats3["success"] == 50
    assert open(out).read().splitlines() == lines
