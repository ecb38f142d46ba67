"""CPU tests of the host-side logic: seed strategies mirror, boundary tables, CSV schema/formatting,
slab partition, C-ABI symbol export, and the test-only host build of the product's solver headers
against the oracle (so the analytic Jacobian and the cascade are checked even without a GPU)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from julia_relaxtime_b200 import _abi as A
from julia_relaxtime_b200 import _lib, boundary, scan, seeds
from julia_relaxtime_b200.distributed import rank_line_indices, slab_bounds
from oracle.oracle import HBARC, Oracle, load_phase_tables
from tests.golden_io import GOLDEN, read_scan_csv, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    """The library loads without a GPU and exports exactly what include/pnjl_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "pnjl_b200.h")).read()
    declared = set(re.findall(r"\b(pnjl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.pnjl_abi_version() == A.ABI_VERSION
    cfg = A.PnjlConfig()
    L.pnjl_default_config(ctypes.byref(cfg))
    assert cfg.p_num == 64 and cfg.t_num == 8 and cfg.max_iter == 1000 and cfg.Nc == 3
    assert abs(cfg.Lambda - 602.3 / 197.327) < 1e-15 and abs(cfg.G - 1.835 / cfg.Lambda ** 2) < 1e-15


def test_julia_shim_binds_only_exported_symbols_with_the_declared_arity():
    """julia/PNJLB200.jl cannot be executed here (no Julia in the image): at least every `ccall` in it must name a symbol
    include/pnjl_b200.h declares, with as many argument types as the C prototype has parameters."""
    root = ROOT
    src = open(os.path.join(root, "julia", "PNJLB200.jl")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "pnjl_b200.h")).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(pnjl_\w+)\s*\(([^;{]*?)\)\s*;", hdr):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    calls = list(re.finditer(r"ccall\(\(:(pnjl_\w+), LIB\),\s*\w+,\s*\(([^)]*)\)", src))
    assert len(calls) >= 15
    for m in calls:
        name, types = m.group(1), [t for t in m.group(2).split(",") if t.strip()]
        assert name in protos, name
        assert len(types) == protos[name], (name, len(types), protos[name])



def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PnjlError):
        _lib.Engine(p_num=12, t_num=6)


def test_gauleg_matches_numpy_and_reference_properties():
    """tests/unit/integration/test_gausslegendre.jl: weights sum to b-a, symmetry, polynomial exactness."""
    for n in (1, 2, 6, 8, 12, 16, 64, 65, 128):
        x, w = _lib.gauleg(-1.0, 1.0, n)
        xs, ws = np.polynomial.legendre.leggauss(n)
        assert np.abs(x - xs).max() < 5e-15 and np.abs(w - ws).max() < 2e-14
        assert abs(w.sum() - 2.0) < 1e-14
        assert np.abs(x + x[::-1]).max() < 1e-15
    x, w = _lib.gauleg(0.0, 10.0, 12)
    for k in range(0, 24):
        assert abs((w * x ** k).sum() - 10.0 ** (k + 1) / (k + 1)) < 1e-12 * 10.0 ** (k + 1)
    with pytest.raises(ValueError):
        _lib.gauleg(0.0, 1.0, 0)
    with pytest.raises(ValueError):
        _lib.gauleg(1.0, 1.0, 4)


def test_seed_strategy_mirror():
    """tests/unit/pnjl/test_solver_seed_strategies.jl: lengths, update!/reset!/set_phase!, phases."""
    th = [100.0 / HBARC, 100.0 / HBARC]
    assert seeds.get_seed(seeds.DefaultSeed(), th) == seeds.HADRON_SEED_5
    assert seeds.get_seed(seeds.DefaultSeed(), [160.0 / HBARC, 0.0]) == seeds.QUARK_SEED_5
    assert seeds.get_seed(seeds.DefaultSeed(), [100.0 / HBARC, 301.0 / HBARC]) == seeds.QUARK_SEED_5
    assert seeds.get_seed(seeds.DefaultSeed(phase_hint="quark"), [300.0 / HBARC * 1.0000001, 0.0]) == seeds.VERY_HIGH_TEMP_SEED_5
    all6 = seeds.get_all_seeds(seeds.MultiSeed(), th, seeds.FixedMu())
    assert len(all6) == 6 and all(len(s) == 5 for s in all6)
    assert all6[0] == seeds.HADRON_SEED_5 and all6[2] == seeds.WEAK_CHIRAL_CONF_SEED_5 and all6[5] == seeds.HT_GUESS_0p95_SEED_5
    c = seeds.ContinuitySeed()
    assert seeds.get_seed(c, th) == seeds.HADRON_SEED_5
    seeds.update_(c, [1, 2, 3, 4, 5])
    assert seeds.get_seed(c, th) == [1.0, 2.0, 3.0, 4.0, 5.0]
    seeds.reset_(c)
    assert c.previous_solution is None
    t = seeds.PhaseAwareContinuitySeed(0.0)
    assert t.boundary_data is not None and t.boundary_data.T_values == sorted(t.boundary_data.T_values)
    assert seeds.get_seed(t, [100 / HBARC, 300 / HBARC]) == seeds.HADRON_SEED_5
    assert seeds.get_seed(t, [100 / HBARC, 350 / HBARC]) == seeds.QUARK_SEED_5
    seeds.update_(t, [1, 2, 3, 4, 5], 100.0, 300.0)
    assert t.previous_phase == "hadron"
    assert seeds.get_seed(t, [100 / HBARC, 350 / HBARC]) == seeds.QUARK_SEED_5      # flip → fresh quark seed
    assert seeds.get_seed(t, [100 / HBARC, 310 / HBARC]) == [1.0, 2.0, 3.0, 4.0, 5.0]
    assert seeds.get_seed(t, [140 / HBARC, 350 / HBARC]) == [1.0, 2.0, 3.0, 4.0, 5.0]  # crossover: continuity
    seeds.set_phase_(t, "quark")
    assert t.previous_phase == "quark"
    seeds.reset_(t)
    assert t.previous_solution is None and t.previous_phase == "unknown"
    assert math.isnan(boundary.interpolate_mu_c(t.boundary_data, 135.0))                # above CEP
    assert boundary.interpolate_mu_c(t.boundary_data, 10.0) == t.boundary_data.mu_values[0]
    t6 = seeds.PhaseAwareContinuitySeed(0.6)                                            # no data for xi=0.6
    assert seeds._get_current_phase(t6, 100.0, 300.0) == "unknown"


def test_boundary_tables_match_oracle_loader():
    tabs, idx = boundary.default_tables([-0.6, 0.0, 0.2, 0.4, 0.8])
    otabs, oidx = load_phase_tables(os.path.join(GOLDEN, "boundary.csv"), os.path.join(GOLDEN, "cep.csv"),
                                    [-0.6, 0.0, 0.2, 0.4, 0.8])
    assert idx == {k: v for k, v in oidx.items()}
    for a, b in zip(tabs, otabs):
        assert list(a[0]) == list(b[0]) and list(a[1]) == list(b[1]) and a[2] == b[2]
    assert [len(t[0]) for t in tabs] == [25, 21, 19]


def test_julia_float_formatting_reproduces_golden_text():
    n = 0
    for line in open(os.path.join(GOLDEN, "gap_transport_scan_xi-0p6to0p6.csv")):
        if line.startswith("#") or line.startswith("T_MeV"):
            continue
        for tok in line.strip().split(","):
            if tok in ("true", "false") or tok.isdigit():
                continue
            assert scan.julia_float(float(tok)) == tok
            n += 1
    assert n > 10000
    assert scan.julia_float(float("nan")) == "NaN" and scan.julia_float(1e-5) == "1.0e-5"
    assert scan.julia_range(120.0, 400.0, 10.0) == [120.0 + 10.0 * i for i in range(29)]


def test_csv_rows_from_oracle_records_match_golden_text_columns():
    """derived_columns/format_rows build the script's row (cols 1-30); fed with oracle results on a golden line
    the text of the exactly-reproducible columns equals the golden file's."""
    cols = read_scan_csv(os.path.join(GOLDEN, "gap_transport_scan_xi-0p6to0p6.csv"))
    sel = np.nonzero((cols["xi"] == 0.0) & (cols["muB_MeV"] == 800.0))[0]
    T = cols["T_MeV"][sel]
    o = Oracle(p_num=12, t_num=6, max_iter=40)
    tables, index = load_phase_tables(os.path.join(GOLDEN, "boundary.csv"), os.path.join(GOLDEN, "cep.csv"), [0.0])
    res = o.scan_lines([800.0 / 3.0], [0.0], T, tables, np.array([index[0.0]], dtype=np.int32))
    rec = np.zeros((len(T), A.REC_DOUBLES))
    rec[:, 0:5] = res.x.T
    rec[:, 5:8] = res.mass.T
    rec[:, A.REC_OMEGA], rec[:, A.REC_PRESSURE], rec[:, A.REC_ENTROPY], rec[:, A.REC_ENERGY] = (
        res.omega, res.pressure, res.entropy, res.energy)
    rec[:, A.REC_NQ:A.REC_NQ + 3] = res.n_q.T
    rec[:, A.REC_NQBAR:A.REC_NQBAR + 3] = res.n_qbar.T
    rec[:, A.REC_RESNORM], rec[:, A.REC_ITER], rec[:, A.REC_STATUS] = res.residual_norm, res.iterations, res.status
    c = scan.derived_columns(rec, T, 800.0 / 3.0, 800.0, 0.0)
    rows = scan.format_rows(c, len(T))
    assert all(len(r.split(",")) == 47 for r in rows)
    golden_rows = [l.strip() for l in open(os.path.join(GOLDEN, "gap_transport_scan_xi-0p6to0p6.csv"))
                   if not l.startswith("#") and not l.startswith("T_MeV")]
    for j, i in enumerate(sel[1:], start=1):
        g = golden_rows[i].split(",")
        r = rows[j].split(",")
        assert r[:8] == g[:8]                      # T, muq, muB, xi, T_fm, muq_fm, converged, iterations: textual
        for a, b in zip(r[9:30], g[9:30]):         # the rest: numerically (1e-10; round-off differs in the last digits)
            assert abs(float(a) - float(b)) <= 1e-10 * max(abs(float(b)), 1e-6)
    assert len(scan.HEADER) == 47


def test_scan_resume_keys_and_header_check(tmp_path):
    p = tmp_path / "scan.csv"
    p.write_text("# schema: scan_csv_v1\n" + ",".join(scan.HEADER) + "\n" +
                 "120.0,0.0,0.0,0.0," + ",".join(["1.0"] * 43) + "\n")
    assert scan.read_existing_keys(str(p)) == {(120.0, 0.0, 0.0)}
    scan.ensure_output_header_compatible(str(p))
    bad = tmp_path / "bad.csv"
    bad.write_text("T_MeV,muB_MeV,xi\n1,2,3\n")
    with pytest.raises(RuntimeError):
        scan.ensure_output_header_compatible(str(bad))
    g = scan.build_grid([0.0, 0.2, 0.6], [0.0, 800.0, 800.0], [120.0, 130.0])
    assert g.n_lines == 6 and list(g.muB_MeV) == [0.0, 800.0] * 3 and list(g.xi) == [0, 0, 0.2, 0.2, 0.6, 0.6]
    assert list(g.table_idx) == [0, 0, 1, 1, -1, -1]


def test_slab_partition():
    for n_mu, w in ((1024, 8), (10, 4), (3, 4), (7, 1)):
        b = slab_bounds(n_mu, w)
        assert b[0][0] == 0 and b[-1][1] == n_mu and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
        for layout in ("interleaved", "slab"):
            per_rank = [rank_line_indices(3, n_mu, r, w, layout) for r in range(w)]
            assert sorted(np.concatenate(per_rank).tolist()) == list(range(3 * n_mu))
            assert max(len(p) for p in per_rank) - min(len(p) for p in per_rank) <= 3
        # interleaved: every rank gets a sample of the whole mu range (what balances the expensive 280-360 MeV band)
        if n_mu >= 4 * w:
            for r in range(w):
                mu_idx = rank_line_indices(1, n_mu, r, w, "interleaved")
                assert mu_idx.min() < w and mu_idx.max() >= n_mu - w


# ---- the product's math/solver headers, built for the host (tests/hostsim), against the oracle -------------
@pytest.fixture(scope="module")
def sim_and_oracle():
    from tests.hostsim.hostsim import HostSim
    o = Oracle(p_num=12, t_num=6, max_iter=40)
    return HostSim(o.p_nodes, o.p_w, o.c_nodes, o.c_w, max_iter=40), o


def test_analytic_fj_and_thermo_match_ad_oracle(sim_and_oracle):
    hs, o = sim_and_oracle
    rng = np.random.default_rng(0)
    for _ in range(60):
        T, mu, xi = rng.uniform(30, 400) / HBARC, rng.uniform(0, 400) / HBARC, rng.uniform(-0.6, 0.8)
        x = np.array([rng.uniform(-2.2, 0.3), rng.uniform(-2.2, 0.3), rng.uniform(-2.4, -0.3),
                      rng.uniform(-0.05, 1.02), rng.uniform(-0.05, 1.02)])
        F0, J0 = o.FJ(x, T, mu, xi)
        F1, J1 = hs.fj(x, T, mu, xi)
        assert np.abs(F0 - F1).max() <= 2e-12 * (np.abs(F0).max() + 1e-3)
        assert np.abs(J0 - J1).max() <= 2e-12 * np.abs(J0).max()
        t0, t1 = o.thermo(x, T, mu, xi), hs.thermo(x, T, mu, xi)
        for k in t0:
            a, b = np.asarray(t0[k]), np.asarray(t1[k])
            assert np.abs(a - b).max() <= 1e-11 * (np.abs(a).max() + 1e-9), k


def test_floor_branches_match_ad_oracle(sim_and_oracle):
    """Non-physical iterates where the reference's max(., 1e-16) floors and safe_log are active."""
    hs, o = sim_and_oracle
    for x, T, mu in (([-0.2, -0.2, -0.7, -0.6, -0.7], 60.0, 380.0), ([-1.8, -1.8, -2.2, 1.3, 1.4], 150.0, 0.0),
                     ([0.4, 0.4, -0.5, -0.34, 0.2], 40.0, 390.0), ([-0.1, -0.1, -0.4, -2.0, -2.0], 30.0, 395.0)):
        F0, J0 = o.FJ(np.array(x), T / HBARC, mu / HBARC, 0.3)
        F1, J1 = hs.fj(np.array(x), T / HBARC, mu / HBARC, 0.3)
        assert np.abs(F0 - F1).max() <= 1e-11 * (np.abs(F0).max() + 1e-3), (x, F0, F1)
        assert np.abs(J0 - J1).max() <= 1e-11 * np.abs(J0).max(), (x,)


def test_solver_cascade_matches_oracle_on_lines_and_points(sim_and_oracle):
    hs, o = sim_and_oracle
    tables, index = load_phase_tables(os.path.join(GOLDEN, "boundary.csv"), os.path.join(GOLDEN, "cep.csv"), [0.0])
    T = np.linspace(50, 300, 48)
    muq = np.linspace(0, 400, 12)
    tidx = np.full(12, index[0.0], dtype=np.int32)
    res = o.scan_lines(muq, np.zeros(12), T, tables, tidx)
    rec = hs.scan_lines(muq, np.zeros(12), T, tables, tidx).reshape(-1, A.REC_DOUBLES)
    assert (((rec[:, A.REC_STATUS].astype(int) & 1) != 0) == res.converged).all() and res.converged.all()
    for q in range(5):
        scale = np.maximum(np.abs(res.x[q]), 1e-3 if q >= 3 else 0)
        assert (np.abs(rec[:, q] - res.x[q]) / scale).max() <= 1e-9
    assert (rec[:, A.REC_ITER].astype(int) == res.iterations).mean() > 0.99
    rng = np.random.default_rng(3)
    n = 120
    Tp, mup, xip = rng.uniform(20, 400, n) / HBARC, rng.uniform(0, 450, n) / HBARC, rng.uniform(-0.8, 0.8, n)
    for mode, omode in ((A.SEED_MULTI, "multi"), (A.SEED_AUTO, "auto")):
        res = o.solve_points(Tp, mup, xip, omode)
        rec = hs.solve_points(Tp, mup, xip, mode)
        assert (((rec[:, A.REC_STATUS].astype(int) & 1) != 0) == res.converged).all()
        ok = res.converged
        for q in range(3):
            assert rel(rec[ok, 5 + q], res.mass[q][ok]).max() <= 1e-9


def test_isotropic_collapse_host_build_matches_full_mesh(sim_and_oracle):
    """xi == 0 with the cos(theta) weights pre-summed (p_num nodes per pass) vs the full p_num * t_num mesh of the
    reference: F, J and the thermo sums agree to round-off, and so does the oracle's full-mesh AD evaluation."""
    from tests.hostsim.hostsim import HostSim
    hs, o = sim_and_oracle
    full = HostSim(o.p_nodes, o.p_w, o.c_nodes, o.c_w, max_iter=40, isotropic_collapse=False)
    rng = np.random.default_rng(8)
    for _ in range(30):
        T, mu = rng.uniform(30, 400) / HBARC, rng.uniform(0, 400) / HBARC
        x = np.array([rng.uniform(-2.2, 0.3), rng.uniform(-2.2, 0.3), rng.uniform(-2.4, -0.3),
                      rng.uniform(-0.05, 1.02), rng.uniform(-0.05, 1.02)])
        F1, J1 = hs.fj(x, T, mu, 0.0)
        F2, J2 = full.fj(x, T, mu, 0.0)
        F0, J0 = o.FJ(x, T, mu, 0.0)
        assert np.abs(F1 - F2).max() <= 2e-13 * (np.abs(F2).max() + 1e-3) and np.abs(J1 - J2).max() <= 2e-13 * np.abs(J2).max()
        assert np.abs(F1 - F0).max() <= 2e-12 * (np.abs(F0).max() + 1e-3) and np.abs(J1 - J0).max() <= 2e-12 * np.abs(J0).max()
        t1, t2 = hs.thermo(x, T, mu, 0.0), full.thermo(x, T, mu, 0.0)
        for k in t1:
            a, b = np.asarray(t1[k]), np.asarray(t2[k])
            assert np.abs(a - b).max() <= 1e-12 * (np.abs(b).max() + 1e-9), k
        # xi != 0 never takes the collapsed mesh
        F3, J3 = hs.fj(x, T, mu, 0.3)
        F4, J4 = full.fj(x, T, mu, 0.3)
        assert (F3 == F4).all() and (J3 == J4).all()


def test_sliced_line_march_is_bit_identical_to_one_go(sim_and_oracle):
    """scan_line_slice (the resumable march of the line-march kernel: tracker state parked in a LineState between time slices,
    a fresh solver object per slice) gives bit-identical records to scan_line for every slice length, on lines that cross the
    first-order boundary (phase switches, fallbacks) and on lines without a table."""
    hs, o = sim_and_oracle
    tables, index = load_phase_tables(os.path.join(GOLDEN, "boundary.csv"), os.path.join(GOLDEN, "cep.csv"), [0.0, 0.2])
    T = np.linspace(50, 300, 61)
    muq = np.array([0.0, 150.0, 310.0, 335.0, 350.0, 400.0, 320.0, 345.0])
    xi = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.2, -0.4])
    tidx = np.array([index.get(x, -1) for x in xi], dtype=np.int32)
    whole = hs.scan_lines(muq, xi, T, tables, tidx, mode=0)
    assert ((whole[:, :, A.REC_STATUS].astype(int) & A.ST_PHASE_SWITCH) != 0).any()
    for q in (1, 7, 32, 61, 500):
        sliced = hs.scan_lines(muq, xi, T, tables, tidx, mode=100 + q)
        assert np.array_equal(whole, sliced, equal_nan=True), q


def test_boundary_table_bisection_equals_reference_scan():
    """current_phase finds the table segment by bisection; the reference's interpolate_mu_c (SeedStrategies.jl:446-475, mirrored
    in boundary.py) scans linearly.  Same phase for every T, including exactly on the nodes, on a table far longer than the 64 rows
    the first ABI allowed (tables regenerated by the dual-branch scan can be that long)."""
    from tests.hostsim.hostsim import HostSim
    o = Oracle(p_num=12, t_num=6, max_iter=40)
    hs = HostSim(o.p_nodes, o.p_w, o.c_nodes, o.c_w, max_iter=40)
    rng = np.random.default_rng(4)
    n = 1500
    Ts = np.sort(rng.uniform(40.0, 130.0, n))
    Ts[100] = Ts[101]                                   # a repeated node
    mus = 360.0 - 0.5 * (Ts - 40.0) + rng.normal(0, 0.05, n)
    data = boundary.PhaseBoundaryData(list(Ts), list(mus), 131.0, 290.0, 0.0)
    probe = np.concatenate([rng.uniform(30.0, 140.0, 400), Ts[::50], [Ts[0], Ts[-1], Ts[100]]])
    # the host build exposes the tracker through a scan: phase switches happen exactly where mu crosses mu_c(T); compare the
    # interpolation itself instead through the seeds mirror
    for T in probe:
        mu_c = boundary.interpolate_mu_c(data, float(T))
        # bisection re-implemented here exactly as in csrc/pnjl_solver.cuh
        if T > data.T_CEP:
            got = math.nan                              # current_phase: crossover above the CEP, before any table lookup
        elif T <= Ts[0]:
            got = mus[0]
        elif T >= Ts[-1]:
            got = mus[-1]
        else:
            lo, hi = 1, n - 1
            while lo < hi:
                mid = (lo + hi) >> 1
                if Ts[mid] >= T:
                    hi = mid
                else:
                    lo = mid + 1
            i = lo - 1
            got = mus[i] + (T - Ts[i]) / (Ts[i + 1] - Ts[i]) * (mus[i + 1] - mus[i])
        assert got == mu_c or (math.isnan(got) and math.isnan(mu_c)), (T, got, mu_c)
    # and through the compiled header: a line at fixed mu marched over T switches phase exactly where the mirror says
    Tg = np.linspace(45.0, 129.0, 300)
    muq = 330.0
    want = np.array([boundary.interpolate_mu_c(data, float(t)) for t in Tg])
    rec = hs.scan_lines([muq], [0.0], Tg, [data.as_table()], np.array([0], dtype=np.int32))[0]
    sw = (rec[:, A.REC_STATUS].astype(int) & A.ST_PHASE_SWITCH) != 0
    phase = np.where(muq < want, 1, 2)
    flips = np.nonzero(phase[1:] != phase[:-1])[0] + 1
    assert set(np.nonzero(sw)[0]) <= set(flips.tolist()) and (len(flips) == 0 or sw.any())


def test_struct_layout_self_check():
    """pnjl_sizeof_* / pnjl_config_field_offset agree with the ctypes mirror (the same check julia/PNJLB200.jl runs at load)."""
    L = _lib.load()
    assert A.check_layout(L) == []
    assert L.pnjl_config_field_offset(b"no_such_field") == -1
    assert L.pnjl_sizeof_config() == ctypes.sizeof(A.PnjlConfig)


def test_out_buffer_validation():
    """Caller-supplied result buffers are checked before their address goes to the C ABI."""
    good = np.empty((4, 3, A.REC_DOUBLES))
    assert A.check_records(good, 4 * 3 * A.REC_DOUBLES) is good
    for bad in (np.empty((4, 3, A.REC_DOUBLES), dtype=np.float32), np.empty((4, 3, A.REC_DOUBLES)).transpose(1, 0, 2),
                np.empty((4, 2, A.REC_DOUBLES)), np.empty(4 * 3 * A.REC_DOUBLES + 1)[1:]):
        with pytest.raises(ValueError):
            A.check_records(bad, 4 * 3 * A.REC_DOUBLES)
    ro = np.empty((4, 3, A.REC_DOUBLES))
    ro.setflags(write=False)
    with pytest.raises(ValueError):
        A.check_records(ro, 4 * 3 * A.REC_DOUBLES)


def test_lane_parallel_finish_and_elimination_match_redundant_version(sim_and_oracle):
    """csrc/pnjl_lean.cuh (what the line-march kernel runs after every Jacobian pass: one matrix entry per lane, elimination
    through a shared scratch line), emulated lane by lane on the CPU, against finish_fj + lu_solve5_regs: same F, same Newton
    direction — on physical states, near-singular Jacobians (close to the CEP) and states with phi_u != phi_d."""
    hs, o = sim_and_oracle
    rng = np.random.default_rng(17)
    worst_F = worst_p = 0.0
    n_done = 0
    for trial in range(300):
        T, mu, xi = rng.uniform(40, 350) / HBARC, rng.uniform(0, 400) / HBARC, rng.choice([0.0, -0.4, 0.3, 0.8])
        if trial < 20:
            T, mu, xi = (130.94 + rng.normal(0, 0.3)) / HBARC, (291.21 + rng.normal(0, 0.3)) / HBARC, 0.0     # around the CEP
        u = rng.uniform(-2.0, -0.05)
        x = np.array([u, u if trial % 3 else rng.uniform(-2.0, -0.05), rng.uniform(-2.3, -0.4), rng.uniform(0.0, 0.98),
                      rng.uniform(0.0, 0.98)])
        F0, p0, rc0 = hs.fj_step(x, T, mu, xi)
        F1, p1, rc1 = hs.fj_step(x, T, mu, xi, lean=True)
        if rc1 < 0:
            continue
        n_done += 1
        assert rc0 == rc1 == 1
        worst_F = max(worst_F, np.abs(F0 - F1).max() / (np.abs(F0).max() + 1e-3))
        worst_p = max(worst_p, np.abs(p0 - p1).max() / (np.abs(p0).max() + 1e-12))
    assert n_done > 250 and worst_F <= 1e-14 and worst_p <= 1e-9, (n_done, worst_F, worst_p)
