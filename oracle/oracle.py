"""ctypes binding of the CPU oracle (oracle/pnjl_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from the product package julia_relaxtime_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pnjl_oracle.cpp")
LIB = os.path.join(HERE, "_build", "libpnjl_oracle.so")

HBARC = 197.327


def build(force=False):
    """Compile the oracle with g++ (-O2, no fast-math, no FMA contraction, OpenMP)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fno-fast-math", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
           "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


class Config(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("hbarc", "Lambda", "m_ud0", "m_s0", "G", "K", "T0", "a0", "a1", "a2", "b3", "rho0")] + [
        ("Nc", C.c_int32), ("p_num", C.c_int32), ("t_num", C.c_int32),
        ("p_nodes", C.POINTER(C.c_double)), ("p_w", C.POINTER(C.c_double)),
        ("c_nodes", C.POINTER(C.c_double)), ("c_w", C.POINTER(C.c_double)),
        ("xtol", C.c_double), ("ftol", C.c_double), ("residual_norm_max", C.c_double), ("phi_tol", C.c_double),
        ("max_iter", C.c_int32), ("tr_fallback", C.c_int32), ("auto_multiseed_fallback", C.c_int32),
        ("omega_tie_rel", C.c_double), ("n_threads", C.c_int32)]


class Out(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in
                ("x", "mass", "omega", "pressure", "rho_norm", "entropy", "energy", "n_q", "n_qbar",
                 "residual_norm")] + [
        ("iterations", C.POINTER(C.c_int32)), ("status", C.POINTER(C.c_int32)), ("n_fj", C.POINTER(C.c_int32))]


class Table(C.Structure):
    _fields_ = [("T_MeV", C.POINTER(C.c_double)), ("mu_c_MeV", C.POINTER(C.c_double)), ("n", C.c_int32),
                ("T_CEP", C.c_double)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


ST_CONVERGED, ST_USED_TR, ST_TR_ATTEMPTED, ST_USED_MULTISEED = 1, 2, 4, 8
ST_SEED_SHIFT, ST_PHASE_SWITCH, ST_NONFINITE, ST_ALL_SEEDS_FAILED = 4, 128, 256, 512
ST_PROMOTED, ST_REFINED, ST_CAND_SHIFT, ST_NO_RESULT = 1024, 2048, 12, 16384


class Result:
    """SoA result block; arrays are numpy, leading dim = component."""

    def __init__(self, n):
        self.n = n
        self.x = np.zeros((5, n))
        self.mass = np.zeros((3, n))
        self.omega = np.zeros(n)
        self.pressure = np.zeros(n)
        self.rho_norm = np.zeros(n)
        self.entropy = np.zeros(n)
        self.energy = np.zeros(n)
        self.n_q = np.zeros((3, n))
        self.n_qbar = np.zeros((3, n))
        self.residual_norm = np.zeros(n)
        self.iterations = np.zeros(n, dtype=np.int32)
        self.status = np.zeros(n, dtype=np.int32)
        self.n_fj = np.zeros(n, dtype=np.int32)

    def c_struct(self):
        return Out(_dp(self.x), _dp(self.mass), _dp(self.omega), _dp(self.pressure), _dp(self.rho_norm),
                   _dp(self.entropy), _dp(self.energy), _dp(self.n_q), _dp(self.n_qbar), _dp(self.residual_norm),
                   _ip(self.iterations), _ip(self.status), _ip(self.n_fj))

    @property
    def converged(self):
        return (self.status & ST_CONVERGED) != 0


class Oracle:
    """One oracle configuration: constants (config/pnjl/default.toml), quadrature mesh, solver options."""

    def __init__(self, p_num=64, t_num=8, max_iter=1000, tr_fallback=True, auto_multiseed_fallback=True,
                 omega_tie_rel=1e-12, n_threads=0, nodes=None):
        self.lib = C.CDLL(build())
        L = self.lib
        L.oracle_omega.restype = C.c_double
        L.oracle_omega.argtypes = [C.POINTER(Config), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
        L.oracle_FJ.argtypes = [C.POINTER(Config), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double,
                                C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_thermo.argtypes = [C.POINTER(Config), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double,
                                    C.POINTER(C.c_double)]
        L.oracle_nlsolve_trace.restype = C.c_int32
        L.oracle_nlsolve_trace.argtypes = [C.POINTER(Config), C.POINTER(C.c_double), C.c_double, C.c_double,
                                           C.c_double, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.c_int32]
        L.oracle_solve_points.argtypes = [C.POINTER(Config), C.c_int64, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_int32,
                                          C.POINTER(C.c_double), C.POINTER(Out), C.POINTER(C.c_double)]
        L.oracle_scan_lines.argtypes = [C.POINTER(Config), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                        C.POINTER(Table), C.POINTER(Out)]
        L.oracle_scan_lines_from.argtypes = [C.POINTER(Config), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                             C.POINTER(Table), C.POINTER(C.c_double), C.c_double, C.POINTER(Out)]
        L.oracle_tmu_scan.argtypes = [C.POINTER(Config), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                      C.POINTER(Table), C.POINTER(Out)]
        L.oracle_gauleg.argtypes = [C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_num_threads.restype = C.c_int32

        self.p_num, self.t_num = p_num, t_num
        if nodes is None:
            self.p_nodes, self.p_w = self.gauleg(0.0, 10.0, p_num)   # Integrals.jl:75-80
            self.c_nodes, self.c_w = self.gauleg(0.0, 1.0, t_num)    # Integrals.jl:67-73 (weights doubled in C)
        else:
            self.p_nodes, self.p_w, self.c_nodes, self.c_w = [np.ascontiguousarray(a, dtype=np.float64)
                                                              for a in nodes]
        hb = HBARC
        Lam = 602.3 / hb
        self.cfg = Config(hbarc=hb, Lambda=Lam, m_ud0=5.5 / hb, m_s0=140.7 / hb, G=1.835 / (Lam * Lam),
                          K=12.36 / Lam ** 5, T0=210.0 / hb, a0=3.51, a1=-2.47, a2=15.2, b3=-1.75, rho0=0.16,
                          Nc=3, p_num=p_num, t_num=t_num, p_nodes=_dp(self.p_nodes), p_w=_dp(self.p_w),
                          c_nodes=_dp(self.c_nodes), c_w=_dp(self.c_w), xtol=1e-9, ftol=1e-9,
                          residual_norm_max=1e-6, phi_tol=1e-8, max_iter=max_iter,
                          tr_fallback=int(tr_fallback), auto_multiseed_fallback=int(auto_multiseed_fallback),
                          omega_tie_rel=omega_tie_rel, n_threads=n_threads)

    def gauleg(self, a, b, n):
        x = np.zeros(n)
        w = np.zeros(n)
        self.lib.oracle_gauleg(a, b, n, _dp(x), _dp(w))
        return x, w

    def num_threads(self):
        return self.cfg.n_threads if self.cfg.n_threads > 0 else int(self.lib.oracle_num_threads())

    def omega(self, x, T_fm, mu_fm, xi=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.lib.oracle_omega(C.byref(self.cfg), _dp(x), T_fm, mu_fm, xi)

    def FJ(self, x, T_fm, mu_fm, xi=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        F = np.zeros(5)
        J = np.zeros((5, 5))
        self.lib.oracle_FJ(C.byref(self.cfg), _dp(x), T_fm, mu_fm, xi, _dp(F), _dp(J))
        return F, J

    def thermo(self, x, T_fm, mu_fm, xi=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros(17)
        self.lib.oracle_thermo(C.byref(self.cfg), _dp(x), T_fm, mu_fm, xi, _dp(o))
        return dict(omega=o[0], pressure=o[1], rho_norm=o[2], entropy=o[3], energy=o[4], rho=o[5:8].copy(),
                    masses=o[8:11].copy(), n_q=o[11:14].copy(), n_qbar=o[14:17].copy())

    def nlsolve_trace(self, x0, T_fm, mu_fm, xi=0.0, method="newton", cap=1100):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        zero = np.zeros(5)
        res = np.zeros(4)
        tr = np.zeros((cap, 5))
        n = self.lib.oracle_nlsolve_trace(C.byref(self.cfg), _dp(x0), T_fm, mu_fm, xi,
                                          0 if method == "newton" else 1, _dp(zero), _dp(res), _dp(tr), cap)
        return dict(zero=zero, iterations=int(res[0]), residual_norm=res[1], x_converged=bool(res[2]),
                    f_converged=bool(res[3]), trace=tr[:min(n, cap)].copy())

    def solve_points(self, T_fm, mu_fm, xi, seed_mode="multi", seeds=None, per_seed=False):
        T_fm = np.ascontiguousarray(np.atleast_1d(T_fm), dtype=np.float64)
        n = T_fm.size
        mu_fm = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(mu_fm), (n,)), dtype=np.float64)
        xi = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(xi), (n,)), dtype=np.float64)
        mode = {"explicit": 0, "auto": 1, "multi": 2}[seed_mode]
        n_seeds = 6 if mode == 2 else 1
        sp = None
        if mode == 0:
            seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(n, -1, 5)
            n_seeds = seeds.shape[1]
            sp = _dp(seeds)
        out = Result(n)
        ps = np.zeros((n, n_seeds, 8)) if per_seed else None
        cs = out.c_struct()
        self.lib.oracle_solve_points(C.byref(self.cfg), n, _dp(T_fm), _dp(mu_fm), _dp(xi), mode, n_seeds, sp,
                                     C.byref(cs), _dp(ps) if per_seed else None)
        if per_seed:
            out.per_seed = ps
        return out

    def scan_lines(self, muq_MeV, xi, T_MeV, tables=None, table_idx=None, init_x=None, init_T_MeV=0.0):
        """tables: list of (T_MeV[], mu_c_MeV[], T_CEP); table_idx[line] = index into tables or -1.
        init_x [n_lines, 5] (optional): converged solutions of the point just before T_MeV[0] (at init_T_MeV) — the march
        then starts in the middle of a line with the tracker state the full march would have there."""
        muq_MeV = np.ascontiguousarray(muq_MeV, dtype=np.float64)
        n_lines = muq_MeV.size
        xi = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(xi), (n_lines,)), dtype=np.float64)
        T_MeV = np.ascontiguousarray(T_MeV, dtype=np.float64)
        tables = tables or []
        keep = []
        ctabs = (Table * max(1, len(tables)))()
        for i, (tt, mm, tcep) in enumerate(tables):
            tt = np.ascontiguousarray(tt, dtype=np.float64)
            mm = np.ascontiguousarray(mm, dtype=np.float64)
            keep += [tt, mm]
            ctabs[i] = Table(_dp(tt), _dp(mm), tt.size, tcep)
        if table_idx is None:
            table_idx = np.full(n_lines, -1, dtype=np.int32)
        table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
        out = Result(n_lines * T_MeV.size)
        cs = out.c_struct()
        ix = None
        if init_x is not None:
            init_x = np.ascontiguousarray(init_x, dtype=np.float64).reshape(n_lines, 5)
            ix = _dp(init_x)
        self.lib.oracle_scan_lines_from(C.byref(self.cfg), n_lines, _dp(muq_MeV), _dp(xi), _ip(table_idx), T_MeV.size,
                                        _dp(T_MeV), len(tables), ctabs, ix, float(init_T_MeV), C.byref(cs))
        out.n_lines, out.n_T = n_lines, T_MeV.size
        return out


def _tmu_scan(self, T_MeV, xi, mu_MeV, tables=None, table_idx=None):
    """TmuScan semantics: one line per (xi, T), marching mu_MeV in the given order."""
    T_MeV = np.ascontiguousarray(T_MeV, dtype=np.float64)
    n_lines = T_MeV.size
    xi = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(xi), (n_lines,)), dtype=np.float64)
    mu_MeV = np.ascontiguousarray(mu_MeV, dtype=np.float64)
    tables = tables or []
    keep = []
    ctabs = (Table * max(1, len(tables)))()
    for i, (tt, mm, tcep) in enumerate(tables):
        tt = np.ascontiguousarray(tt, dtype=np.float64)
        mm = np.ascontiguousarray(mm, dtype=np.float64)
        keep += [tt, mm]
        ctabs[i] = Table(_dp(tt), _dp(mm), tt.size, tcep)
    if table_idx is None:
        table_idx = np.full(n_lines, -1, dtype=np.int32)
    table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
    out = Result(n_lines * mu_MeV.size)
    cs = out.c_struct()
    self.lib.oracle_tmu_scan(C.byref(self.cfg), n_lines, _dp(T_MeV), _dp(xi), _ip(table_idx), mu_MeV.size, _dp(mu_MeV),
                             len(tables), ctabs, C.byref(cs))
    out.n_lines, out.n_mu = n_lines, mu_MeV.size
    return out


Oracle.tmu_scan = _tmu_scan


def load_phase_tables(boundary_csv, cep_csv, xis):
    """Python restatement of load_phase_boundary (SeedStrategies.jl:388-436) for test use.

    Returns (tables, lookup) where lookup(xi) → table index or -1 (no rows and no CEP for that xi)."""
    import csv
    ceps = {}
    with open(cep_csv) as f:
        for row in csv.reader(f):
            if not row or row[0].startswith("xi"):
                continue
            ceps[float(row[0])] = float(row[1])
    rows = []
    with open(boundary_csv) as f:
        for row in csv.reader(f):
            if not row or row[0].startswith("xi"):
                continue
            rows.append((float(row[0]), float(row[1]), float(row[2])))
    tables, index = [], {}
    for xi in xis:
        tcep = float("nan")
        for k, v in ceps.items():
            if abs(k - xi) <= 1e-6:
                tcep = v
                break
        sel = sorted([(t, m) for (x, t, m) in rows if abs(x - xi) <= 1e-6])
        if not sel and tcep != tcep:
            index[xi] = -1
            continue
        index[xi] = len(tables)
        tables.append((np.array([s[0] for s in sel]), np.array([s[1] for s in sel]), tcep))
    return tables, index


def _oneloop_A(self, m, mu, T, Phi, Phib, nodes=None, weights=None):
    """OneLoopIntegrals.A(m, mu, T, Phi, Phibar, nodes, weights); default rule = gauleg(0, 10, 64)."""
    if nodes is None:
        nodes, weights = self.gauleg(0.0, 10.0, 64)
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    self.lib.oracle_oneloop_A.restype = C.c_double
    return self.lib.oracle_oneloop_A(C.byref(self.cfg), C.c_double(m), C.c_double(mu), C.c_double(T), C.c_double(Phi),
                                     C.c_double(Phib), C.c_int32(nodes.size), _dp(nodes), _dp(weights))


def _effective_couplings(self, G, K, G_u, G_s):
    out = np.zeros(12)
    self.lib.oracle_effective_couplings.restype = None
    self.lib.oracle_effective_couplings(C.c_double(G), C.c_double(K), C.c_double(G_u), C.c_double(G_s), _dp(out))
    return out


def _couplings_batch(self, T, mu, m_u, m_s, Phi, Phib, nodes=None, weights=None):
    """build_K_data for arrays of states: [n, 16] = A_u, A_s, G_u, G_s, K0+-, K123+-, K4567+-, K8+-, K08+-, detK+-."""
    if nodes is None:
        nodes, weights = self.gauleg(0.0, 10.0, 64)
    arrs = [np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64) for a in (T, mu, m_u, m_s, Phi, Phib)]
    n = arrs[0].size
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    aux = np.zeros((n, 16))
    self.lib.oracle_couplings_batch.restype = None
    self.lib.oracle_couplings_batch(C.byref(self.cfg), C.c_int64(n), *[_dp(a) for a in arrs], C.c_int32(nodes.size),
                                    _dp(nodes), _dp(weights), _dp(aux))
    return aux


Oracle.oneloop_A = _oneloop_A
Oracle.effective_couplings = _effective_couplings
Oracle.couplings_batch = _couplings_batch


def _dual_branch(self, T_MeV, xi, mu_MeV):
    """DualBranchScan semantics: Result of n_lines * 2 * n_mu entries, index (line * 2 + branch) * n_mu + imu."""
    T_MeV = np.ascontiguousarray(np.atleast_1d(T_MeV), dtype=np.float64)
    n_lines = T_MeV.size
    xi = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(xi), (n_lines,)), dtype=np.float64)
    mu_MeV = np.ascontiguousarray(mu_MeV, dtype=np.float64)
    out = Result(n_lines * 2 * mu_MeV.size)
    cs = out.c_struct()
    self.lib.oracle_dual_branch(C.byref(self.cfg), C.c_int64(n_lines), _dp(T_MeV), _dp(xi), C.c_int32(mu_MeV.size), _dp(mu_MeV),
                                C.byref(cs))
    out.n_lines, out.n_mu = n_lines, mu_MeV.size
    return out


Oracle.dual_branch = _dual_branch


TD_NAMES = ("v_n_sq", "dmuB_dT_sigma", "M_u", "M_d", "M_s", "dM_u_dT", "dM_d_dT", "dM_s_dT", "dM_u_dmuB", "dM_d_dmuB",
            "dM_s_dmuB", "s", "n_B", "P", "eps", "dP_dT", "dP_dmu", "dEps_dT", "dEps_dmu", "dn_dT", "dn_dmu", "dP_deps_n",
            "dP_dn_eps", "dM_u_dmu", "dM_d_dmu", "dM_s_dmu")


def _thermo_derivatives(self, T_fm, mu_fm, xi, x):
    """ThermoDerivatives.jl (bulk_viscosity_coefficients, thermo_derivatives, mass_derivatives order 1) at given converged
    states x [n, 5]: dict name -> array, by exact AD of Omega over (x, T, mu)."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 5)
    n = x.shape[0]
    T_fm = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(T_fm), (n,)), dtype=np.float64)
    mu_fm = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(mu_fm), (n,)), dtype=np.float64)
    xi = np.ascontiguousarray(np.broadcast_to(np.atleast_1d(xi), (n,)), dtype=np.float64)
    out = np.zeros((n, 32))
    self.lib.oracle_thermo_derivatives.restype = None
    self.lib.oracle_thermo_derivatives(C.byref(self.cfg), C.c_int64(n), _dp(T_fm), _dp(mu_fm), _dp(xi), _dp(x), _dp(out))
    return {name: out[:, i].copy() for i, name in enumerate(TD_NAMES)}


Oracle.thermo_derivatives = _thermo_derivatives
