"""Build and load libpnjl_b200.so (csrc/pnjl_kernels.cu → sm_100a) and wrap its C ABI (include/pnjl_b200.h).

There is no CPU path: if the shared library is missing, cannot be built, or no B200 is visible,
the calls raise `PnjlError` — they never fall back to anything else.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _abi
from .constants import DEFAULT, PNJLConstants

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libpnjl_b200.so")
SOURCES = [os.path.join(CSRC, f) for f in ("pnjl_kernels.cu", "pnjl_math.cuh", "pnjl_solver.cuh", "pnjl_march.cuh", "pnjl_lean.cuh")] + [
    os.path.join(HERE, "..", "include", "pnjl_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class PnjlError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """nvcc → csrc/_build/libpnjl_b200.so (in-tree, so it travels to the GPU box)."""
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in SOURCES)):
        return LIB_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("PNJL_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
    try:
        out = subprocess.run(cmd, check=True, capture_output=True, text=True)
    except (OSError, subprocess.CalledProcessError) as e:
        raise PnjlError("building libpnjl_b200.so failed: %s\n%s" % (e, getattr(e, "stderr", ""))) from e
    if verbose:
        print(out.stderr)
    return LIB_PATH


_LIB = None


def load():
    """dlopen the library (building it first when the sources are newer) and declare the prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # PNJL_LIB: developer knob — load an experimental build of the same sources (e.g. -DPNJL_PROFILE_PHASES) instead
    path = os.environ.get("PNJL_LIB") or build()
    try:
        L = C.CDLL(path)
    except OSError as e:
        raise PnjlError("cannot load %s: %s" % (path, e)) from e
    H = C.c_void_p
    dp, ip = _abi.c_double_p, _abi.c_int32_p
    L.pnjl_abi_version.restype = C.c_int
    L.pnjl_last_error.restype = C.c_char_p
    L.pnjl_default_config.argtypes = [C.POINTER(_abi.PnjlConfig)]
    L.pnjl_default_config.restype = None
    L.pnjl_create.argtypes = [C.POINTER(_abi.PnjlConfig), C.POINTER(H)]
    L.pnjl_destroy.argtypes = [H]
    L.pnjl_destroy.restype = None
    L.pnjl_alloc_pinned.argtypes = [C.c_uint64, C.POINTER(C.c_void_p)]
    L.pnjl_free_pinned.argtypes = [C.c_void_p]
    L.pnjl_gauleg.argtypes = [C.c_double, C.c_double, C.c_int32, dp, dp]
    L.pnjl_solve_points_host.argtypes = [H, C.c_int64, dp, dp, dp, C.c_int32, C.c_int32, dp, dp]
    L.pnjl_solve_points_device.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
    L.pnjl_set_boundaries.argtypes = [H, C.c_int32, C.POINTER(_abi.PnjlBoundary)]
    L.pnjl_scan_lines_host.argtypes = [H, C.c_int64, dp, dp, ip, C.c_int32, dp, dp]
    L.pnjl_scan_lines_device.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    L.pnjl_scan_lines_device_indexed.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.pnjl_ipc_alloc.argtypes = [C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
    L.pnjl_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.pnjl_ipc_close.argtypes = [C.c_void_p]
    L.pnjl_ipc_free.argtypes = [C.c_void_p]
    L.pnjl_tmu_scan_host.argtypes = [H, C.c_int64, dp, dp, ip, C.c_int32, dp, dp]
    L.pnjl_tmu_scan_device.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    L.pnjl_dual_branch_host.argtypes = [H, C.c_int64, dp, dp, C.c_int32, dp, dp]
    L.pnjl_dual_branch_device.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pnjl_set_oneloop_rule.argtypes = [H, C.c_int32, dp, dp]
    L.pnjl_effective_couplings_host.argtypes = [H, C.c_int64, dp, dp, dp, dp, dp, dp, dp]
    L.pnjl_effective_couplings_device.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pnjl_scan_lines_couplings_host.argtypes = [H, C.c_int64, dp, dp, ip, C.c_int32, dp, dp, dp]
    L.pnjl_eval_fj_host.argtypes = [H, C.c_int64, dp, dp, dp, dp, dp]
    L.pnjl_eval_state_host.argtypes = [H, C.c_int64, dp, dp, dp, dp, dp]
    L.pnjl_eval_derivs_host.argtypes = [H, C.c_int64, dp, dp, dp, dp, dp]
    L.pnjl_selftest_math.argtypes = [H, C.c_int64, dp, C.c_int32, dp]
    L.pnjl_get_stats.argtypes = [H, C.POINTER(_abi.PnjlStats)]
    L.pnjl_measure_fp64_peak.argtypes = [H, C.c_double, dp, dp]
    L.pnjl_set_option.argtypes = [H, C.c_char_p, C.c_int64]
    if L.pnjl_abi_version() != _abi.ABI_VERSION:
        raise PnjlError("ABI version mismatch between _abi.py and libpnjl_b200.so")
    bad = _abi.check_layout(L)
    if bad:
        raise PnjlError("struct layout mismatch between _abi.py and libpnjl_b200.so: " + "; ".join(bad))
    _LIB = L
    return L


EXPORTED_SYMBOLS = [
    "pnjl_default_config", "pnjl_alloc_pinned", "pnjl_free_pinned", "pnjl_abi_version", "pnjl_last_error", "pnjl_create", "pnjl_destroy", "pnjl_gauleg",
    "pnjl_solve_points_host", "pnjl_solve_points_device", "pnjl_set_boundaries", "pnjl_scan_lines_host",
    "pnjl_scan_lines_device", "pnjl_scan_lines_device_indexed", "pnjl_ipc_alloc", "pnjl_ipc_open", "pnjl_ipc_close",
    "pnjl_ipc_free", "pnjl_set_oneloop_rule", "pnjl_effective_couplings_host",
    "pnjl_effective_couplings_device", "pnjl_scan_lines_couplings_host", "pnjl_dual_branch_host", "pnjl_dual_branch_device", "pnjl_tmu_scan_host", "pnjl_tmu_scan_device", "pnjl_eval_fj_host", "pnjl_eval_state_host", "pnjl_selftest_math", "pnjl_get_stats", "pnjl_measure_fp64_peak",
    "pnjl_eval_derivs_host", "pnjl_set_option", "pnjl_sizeof_config", "pnjl_sizeof_boundary", "pnjl_sizeof_stats", "pnjl_config_field_offset"]


class PinnedArray:
    """numpy view of page-locked host memory from pnjl_alloc_pinned: result buffers the kernels write in place."""

    def __init__(self, shape):
        L = load()
        self.shape = tuple(int(x) for x in shape)
        n = int(np.prod(self.shape))
        self._p = C.c_void_p()
        rc = L.pnjl_alloc_pinned(C.c_uint64(8 * n), C.byref(self._p))
        if rc != 0:
            raise PnjlError("pnjl_alloc_pinned failed (%d): %s" % (rc, L.pnjl_last_error().decode()))
        buf = (C.c_double * n).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.float64).reshape(self.shape)

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self.array = None
            load().pnjl_free_pinned(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gauleg(a, b, n):
    """gauleg(a, b, n) of src/integration/GaussLegendre.jl:94-119 (host-side, no GPU needed)."""
    L = load()
    x = np.zeros(int(n))
    w = np.zeros(int(n))
    rc = L.pnjl_gauleg(float(a), float(b), int(n), _abi.dptr(x), _abi.dptr(w))
    if rc != 0:
        raise ValueError(L.pnjl_last_error().decode())   # ArgumentError in the reference
    return x, w


class Engine:
    """One handle of the library: constants + quadrature rule + solver options on one GPU."""

    def __init__(self, p_num=64, t_num=8, max_iter=1000, trust_region_fallback=True, auto_multiseed_fallback=True,
                 residual_norm_max=1e-6, omega_tie_rel=1e-12, device=-1, lanes_per_solve=0, nodes=None,
                 isospin_symmetric=True, predict_tol=1e-4, schedule=0, isotropic_collapse=True,
                 consts: PNJLConstants = DEFAULT):
        self.L = load()
        self.p_num, self.t_num = int(p_num), int(t_num)
        self.isotropic_collapse = bool(isotropic_collapse)
        self._keep = None
        k = consts
        cfg = _abi.PnjlConfig(
            hbarc=k.hbarc, Lambda=k.Lambda_inv_fm, m_ud0=k.m_ud0_inv_fm, m_s0=k.m_s0_inv_fm, G=k.G_fm2, K=k.K_fm5,
            T0=k.T0_inv_fm, a0=k.a0, a1=k.a1, a2=k.a2, b3=k.b3, rho0=k.rho0_fm3, Nc=k.N_color,
            p_num=self.p_num, t_num=self.t_num, p_nodes=None, p_w=None, c_nodes=None, c_w=None,
            xtol=1e-9, ftol=1e-9, residual_norm_max=residual_norm_max, phi_tol=1e-8, max_iter=int(max_iter),
            tr_fallback=int(trust_region_fallback), auto_multiseed_fallback=int(auto_multiseed_fallback),
            omega_tie_rel=omega_tie_rel, device=int(device), lanes_per_solve=int(lanes_per_solve),
            predict_tol=float(predict_tol), isospin_symmetric=int(isospin_symmetric), schedule=int(schedule),
            isotropic_collapse=int(isotropic_collapse))
        if nodes is not None:
            self._keep = [np.ascontiguousarray(a, dtype=np.float64) for a in nodes]
            cfg.p_nodes, cfg.p_w, cfg.c_nodes, cfg.c_w = [_abi.dptr(a) for a in self._keep]
        self.cfg = cfg
        self.consts = consts
        self.h = C.c_void_p()
        rc = self.L.pnjl_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise PnjlError("pnjl_create failed (%d): %s" % (rc, self.L.pnjl_last_error().decode()))
        self._tables_key = None

    def close(self):
        if getattr(self, "h", None):
            self.L.pnjl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise PnjlError("%s failed (%d): %s" % (what, rc, self.L.pnjl_last_error().decode()))

    def set_option(self, key, value):
        """Run-time option of the handle (pnjl_set_option): 'schedule', 'march_parts', 'march_quantum', 'isotropic_batch'."""
        self._check(self.L.pnjl_set_option(self.h, key.encode(), int(value)), "pnjl_set_option(%s)" % key)

    # ---- boundaries -------------------------------------------------------------------------------
    def set_boundaries(self, tables):
        """tables: list of (T_MeV[], mu_c_MeV[], T_CEP_MeV)."""
        keep = []
        arr = (_abi.PnjlBoundary * max(1, len(tables)))()
        for i, (tt, mm, tcep) in enumerate(tables):
            tt = _abi.as_f64(tt) if len(tt) else np.zeros(0)
            mm = _abi.as_f64(mm) if len(mm) else np.zeros(0)
            keep += [tt, mm]
            arr[i] = _abi.PnjlBoundary(_abi.dptr(tt), _abi.dptr(mm), int(tt.size), float(tcep))
        self._check(self.L.pnjl_set_boundaries(self.h, len(tables), arr), "pnjl_set_boundaries")

    # ---- host-buffer entry points -------------------------------------------------------------------
    def solve_points(self, T_fm, mu_fm, xi, seed_mode=_abi.SEED_MULTI, seeds=None, out=None):
        T_fm = _abi.as_f64(T_fm)
        n = T_fm.size
        mu_fm = _abi.as_f64(mu_fm, n)
        xi = _abi.as_f64(xi, n)
        n_seeds, sp = 6, None
        if seed_mode == _abi.SEED_EXPLICIT:
            seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(n, -1, 5)
            n_seeds = seeds.shape[1]
            sp = _abi.dptr(seeds)
        rec = _abi.check_records(out, n * _abi.REC_DOUBLES) if out is not None else np.empty((n, _abi.REC_DOUBLES))
        self._check(self.L.pnjl_solve_points_host(self.h, n, _abi.dptr(T_fm), _abi.dptr(mu_fm), _abi.dptr(xi),
                                                  int(seed_mode), int(n_seeds), sp, _abi.dptr(rec)),
                    "pnjl_solve_points_host")
        return rec

    def scan_lines(self, muq_MeV, xi, T_MeV, table_idx=None, out=None):
        muq_MeV = _abi.as_f64(muq_MeV)
        n_lines = muq_MeV.size
        xi = _abi.as_f64(xi, n_lines)
        T_MeV = _abi.as_f64(T_MeV)
        ti = None
        if table_idx is not None:
            table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
            ti = _abi.iptr(table_idx)
        rec = (_abi.check_records(out, n_lines * T_MeV.size * _abi.REC_DOUBLES) if out is not None
               else np.empty((n_lines, T_MeV.size, _abi.REC_DOUBLES)))
        self._check(self.L.pnjl_scan_lines_host(self.h, n_lines, _abi.dptr(muq_MeV), _abi.dptr(xi), ti,
                                                int(T_MeV.size), _abi.dptr(T_MeV), _abi.dptr(rec)),
                    "pnjl_scan_lines_host")
        return rec

    def scan_lines_couplings(self, muq_MeV, xi, T_MeV, table_idx=None):
        """scan_lines plus build_K_data of every point in the same call: (records [L][T][32], aux [L][T][16])."""
        muq_MeV = _abi.as_f64(muq_MeV)
        n_lines = muq_MeV.size
        xi = _abi.as_f64(xi, n_lines)
        T_MeV = _abi.as_f64(T_MeV)
        ti = None
        if table_idx is not None:
            table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
            ti = _abi.iptr(table_idx)
        rec = np.empty((n_lines, T_MeV.size, _abi.REC_DOUBLES))
        aux = np.empty((n_lines, T_MeV.size, _abi.AUX_DOUBLES))
        self._check(self.L.pnjl_scan_lines_couplings_host(self.h, n_lines, _abi.dptr(muq_MeV), _abi.dptr(xi), ti,
                                                          int(T_MeV.size), _abi.dptr(T_MeV), _abi.dptr(rec), _abi.dptr(aux)),
                    "pnjl_scan_lines_couplings_host")
        return rec, aux

    def set_oneloop_rule(self, nodes, weights):
        """Quadrature rule of the one-loop integral A (default gauleg(0, 10, 64) = DEFAULT_MOMENTUM_NODES/WEIGHTS)."""
        nodes, weights = _abi.as_f64(nodes), _abi.as_f64(weights)
        if nodes.size != weights.size:
            raise ValueError("nodes and weights differ in length")
        self._check(self.L.pnjl_set_oneloop_rule(self.h, int(nodes.size), _abi.dptr(nodes), _abi.dptr(weights)),
                    "pnjl_set_oneloop_rule")

    def effective_couplings(self, T_fm, mu_fm, m_u, m_s, Phi, Phibar):
        """A_u, A_s, G_u, G_s and the 12 effective couplings of n states: aux [n][16] (columns _abi.AUX_NAMES)."""
        T_fm = _abi.as_f64(T_fm)
        n = T_fm.size
        arrs = [T_fm] + [_abi.as_f64(a, n) for a in (mu_fm, m_u, m_s, Phi, Phibar)]
        aux = np.empty((n, _abi.AUX_DOUBLES))
        self._check(self.L.pnjl_effective_couplings_host(self.h, n, *[_abi.dptr(a) for a in arrs], _abi.dptr(aux)),
                    "pnjl_effective_couplings_host")
        return aux

    def effective_couplings_device(self, d_records, d_aux, stream=0):
        n = d_records.numel() // _abi.REC_DOUBLES
        self._check(self.L.pnjl_effective_couplings_device(self.h, n, d_records.data_ptr(), d_aux.data_ptr(),
                                                           C.c_void_p(stream)), "pnjl_effective_couplings_device")

    def tmu_scan(self, T_MeV, xi, mu_MeV, table_idx=None, out=None):
        """TmuScan semantics: line l = (xi[l], T_MeV[l]) marches mu_MeV; records [n_lines][n_mu][32]."""
        T_MeV = _abi.as_f64(T_MeV)
        n_lines = T_MeV.size
        xi = _abi.as_f64(xi, n_lines)
        mu_MeV = _abi.as_f64(mu_MeV)
        ti = None
        if table_idx is not None:
            table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
            ti = _abi.iptr(table_idx)
        rec = (_abi.check_records(out, n_lines * mu_MeV.size * _abi.REC_DOUBLES) if out is not None
               else np.empty((n_lines, mu_MeV.size, _abi.REC_DOUBLES)))
        self._check(self.L.pnjl_tmu_scan_host(self.h, n_lines, _abi.dptr(T_MeV), _abi.dptr(xi), ti, int(mu_MeV.size),
                                              _abi.dptr(mu_MeV), _abi.dptr(rec)), "pnjl_tmu_scan_host")
        return rec

    def dual_branch(self, T_MeV, xi, mu_MeV):
        """DualBranchScan: records [n_lines][2][n_mu][32] (branch 0 hadron / mu ascending, 1 quark / mu descending)."""
        T_MeV = _abi.as_f64(T_MeV)
        n_lines = T_MeV.size
        xi = _abi.as_f64(xi, n_lines)
        mu_MeV = _abi.as_f64(mu_MeV)
        rec = np.empty((n_lines, 2, mu_MeV.size, _abi.REC_DOUBLES))
        self._check(self.L.pnjl_dual_branch_host(self.h, n_lines, _abi.dptr(T_MeV), _abi.dptr(xi), int(mu_MeV.size),
                                                 _abi.dptr(mu_MeV), _abi.dptr(rec)), "pnjl_dual_branch_host")
        return rec

    def eval_fj(self, T_fm, mu_fm, xi, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 5)
        n = x.shape[0]
        T_fm, mu_fm, xi = _abi.as_f64(T_fm, n), _abi.as_f64(mu_fm, n), _abi.as_f64(xi, n)
        FJ = np.empty((n, 30))
        self._check(self.L.pnjl_eval_fj_host(self.h, n, _abi.dptr(T_fm), _abi.dptr(mu_fm), _abi.dptr(xi), _abi.dptr(x),
                                             _abi.dptr(FJ)), "pnjl_eval_fj_host")
        return FJ[:, :5].copy(), FJ[:, 5:].reshape(n, 5, 5).copy()

    def eval_state(self, T_fm, mu_fm, xi, x):
        """F, J and the thermodynamic functions at given states (no solve): dict of arrays."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 5)
        n = x.shape[0]
        T_fm, mu_fm, xi = _abi.as_f64(T_fm, n), _abi.as_f64(mu_fm, n), _abi.as_f64(xi, n)
        out = np.empty((n, 48))
        self._check(self.L.pnjl_eval_state_host(self.h, n, _abi.dptr(T_fm), _abi.dptr(mu_fm), _abi.dptr(xi), _abi.dptr(x),
                                                _abi.dptr(out)), "pnjl_eval_state_host")
        return dict(F=out[:, :5].copy(), J=out[:, 5:30].reshape(n, 5, 5).copy(), omega=out[:, 30].copy(),
                    pressure=out[:, 31].copy(), rho_norm=out[:, 32].copy(), entropy=out[:, 33].copy(),
                    energy=out[:, 34].copy(), rho=out[:, 35:38].copy(), n_q=out[:, 38:41].copy(),
                    n_qbar=out[:, 41:44].copy(), masses=out[:, 44:47].copy())

    def eval_derivs(self, T_fm, mu_fm, xi, x):
        """eval_state plus the closed-form partial derivatives in (T, mu) at fixed x (pnjl_eval_derivs_host): adds dF_dT [n, 5],
        dF_dmu [n, 5], s_T, s_mu, nB_T, nB_mu."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 5)
        n = x.shape[0]
        T_fm, mu_fm, xi = _abi.as_f64(T_fm, n), _abi.as_f64(mu_fm, n), _abi.as_f64(xi, n)
        out = np.empty((n, 64))
        self._check(self.L.pnjl_eval_derivs_host(self.h, n, _abi.dptr(T_fm), _abi.dptr(mu_fm), _abi.dptr(xi), _abi.dptr(x),
                                                 _abi.dptr(out)), "pnjl_eval_derivs_host")
        return dict(F=out[:, :5].copy(), J=out[:, 5:30].reshape(n, 5, 5).copy(), omega=out[:, 30].copy(),
                    pressure=out[:, 31].copy(), rho_norm=out[:, 32].copy(), entropy=out[:, 33].copy(),
                    energy=out[:, 34].copy(), rho=out[:, 35:38].copy(), n_q=out[:, 38:41].copy(),
                    n_qbar=out[:, 41:44].copy(), masses=out[:, 44:47].copy(), dF_dT=out[:, 48:53].copy(),
                    dF_dmu=out[:, 53:58].copy(), s_T=out[:, 58].copy(), s_mu=out[:, 59].copy(), nB_T=out[:, 60].copy(),
                    nB_mu=out[:, 61].copy())

    # ---- device-pointer entry points (torch tensors are only used for their data_ptr) --------------
    def solve_points_device(self, d_T, d_mu, d_xi, d_records, seed_mode=_abi.SEED_MULTI, d_seeds=None, n_seeds=6,
                            stream=0):
        n = d_T.numel()
        self._check(self.L.pnjl_solve_points_device(
            self.h, n, d_T.data_ptr(), d_mu.data_ptr(), d_xi.data_ptr(), int(seed_mode), int(n_seeds),
            d_seeds.data_ptr() if d_seeds is not None else None, d_records.data_ptr(), C.c_void_p(stream)),
            "pnjl_solve_points_device")

    def scan_lines_device(self, d_muq, d_xi, d_table_idx, d_T, d_records, stream=0):
        self._check(self.L.pnjl_scan_lines_device(
            self.h, d_muq.numel(), d_muq.data_ptr(), d_xi.data_ptr(),
            d_table_idx.data_ptr() if d_table_idx is not None else None, d_T.numel(), d_T.data_ptr(),
            d_records.data_ptr(), C.c_void_p(stream)), "pnjl_scan_lines_device")

    def scan_lines_device_indexed(self, d_muq, d_xi, d_table_idx, d_T, records_base_ptr, d_out_index, stream=0):
        """records_base_ptr: integer device address (may be a peer GPU's buffer); d_out_index: int64 tensor [n_lines]."""
        self._check(self.L.pnjl_scan_lines_device_indexed(
            self.h, d_muq.numel(), d_muq.data_ptr(), d_xi.data_ptr(),
            d_table_idx.data_ptr() if d_table_idx is not None else None, d_T.numel(), d_T.data_ptr(),
            C.c_void_p(int(records_base_ptr)), d_out_index.data_ptr(), C.c_void_p(stream)), "pnjl_scan_lines_device_indexed")

    def selftest_math(self, x, which):
        """which: 'exp' (x in [-708, 0]), 'rcp', 'rsqrt' — the kernels' branch-free primitives evaluated on the GPU."""
        x = _abi.as_f64(x)
        out = np.empty_like(x)
        self._check(self.L.pnjl_selftest_math(self.h, x.size, _abi.dptr(x), {"exp": 0, "rcp": 1, "rsqrt": 2}[which],
                                              _abi.dptr(out)), "pnjl_selftest_math")
        return out

    def stats(self):
        s = _abi.PnjlStats()
        self._check(self.L.pnjl_get_stats(self.h, C.byref(s)), "pnjl_get_stats")
        return {f[0]: getattr(s, f[0]) for f in _abi.PnjlStats._fields_}

    def measure_fp64_peak(self, seconds=1.0):
        """(burst, sustained) FP64 FMA throughput in TFLOP/s of a register-resident DFMA kernel."""
        burst = C.c_double()
        sus = C.c_double()
        self._check(self.L.pnjl_measure_fp64_peak(self.h, float(seconds), C.byref(burst), C.byref(sus)),
                    "pnjl_measure_fp64_peak")
        return burst.value, sus.value
