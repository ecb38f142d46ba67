"""Loader of the test-only host build of the product's math/solver headers (oracle/pnjl_analytic_cpu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from julia_relaxtime_b200 import _abi
from julia_relaxtime_b200.constants import DEFAULT

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "..", "oracle", "pnjl_analytic_cpu.cpp")
LIB = os.path.join(HERE, "..", "..", "oracle", "_build", "libpnjl_analytic_cpu.so")
CSRC = os.path.join(HERE, "..", "..", "julia_relaxtime_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("pnjl_math.cuh", "pnjl_solver.cuh", "pnjl_lean.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
                           "-Wno-unknown-pragmas", "-o", LIB, SRC])
    return LIB


class HostSim:
    def __init__(self, p_nodes, p_w, c_nodes, c_w, max_iter=1000, tr_fallback=True, auto_multiseed_fallback=True,
                 omega_tie_rel=1e-12, isospin_symmetric=True, predict_tol=1e-4, isotropic_collapse=True, consts=DEFAULT):
        self.lib = C.CDLL(build())
        self.keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (p_nodes, p_w, c_nodes, c_w)]
        k = consts
        self.cfg = _abi.PnjlConfig(
            hbarc=k.hbarc, Lambda=k.Lambda_inv_fm, m_ud0=k.m_ud0_inv_fm, m_s0=k.m_s0_inv_fm, G=k.G_fm2, K=k.K_fm5,
            T0=k.T0_inv_fm, a0=k.a0, a1=k.a1, a2=k.a2, b3=k.b3, rho0=k.rho0_fm3, Nc=k.N_color,
            p_num=len(self.keep[0]), t_num=len(self.keep[2]), p_nodes=_abi.dptr(self.keep[0]),
            p_w=_abi.dptr(self.keep[1]), c_nodes=_abi.dptr(self.keep[2]), c_w=_abi.dptr(self.keep[3]),
            xtol=1e-9, ftol=1e-9, residual_norm_max=1e-6, phi_tol=1e-8, max_iter=max_iter,
            tr_fallback=int(tr_fallback), auto_multiseed_fallback=int(auto_multiseed_fallback),
            omega_tie_rel=omega_tie_rel, device=-1, lanes_per_solve=0, predict_tol=predict_tol,
            isospin_symmetric=int(isospin_symmetric), schedule=0, isotropic_collapse=int(isotropic_collapse))

    def fj(self, x, T, mu, xi):
        x = _abi.as_f64(x)
        F = np.zeros(5)
        J = np.zeros((5, 5))
        self.lib.hostsim_fj(C.byref(self.cfg), _abi.dptr(x), C.c_double(T), C.c_double(mu), C.c_double(xi),
                            _abi.dptr(F), _abi.dptr(J))
        return F, J

    def fj_step(self, x, T, mu, xi, lean=False):
        """F and the Newton direction p = -J^{-1} F at x: redundant-per-lane version, or the lane-parallel one emulated lane by
        lane (csrc/pnjl_lean.cuh).  Returns (F, p, rc) with rc 1 ok, 0 singular, < 0 not on the lean path."""
        x = _abi.as_f64(x)
        F = np.zeros(5)
        p = np.zeros(5)
        fn = self.lib.hostsim_lean_fj_step if lean else self.lib.hostsim_fj_step
        fn.restype = C.c_int
        rc = fn(C.byref(self.cfg), _abi.dptr(x), C.c_double(T), C.c_double(mu), C.c_double(xi), _abi.dptr(F), _abi.dptr(p))
        return F, p, rc

    def derivs(self, x, T, mu, xi):
        """Closed-form partial derivatives in (T, mu) at fixed x: dF_dT[5], dF_dmu[5], s_T, s_mu, nB_T, nB_mu."""
        x = _abi.as_f64(x)
        o = np.zeros(16)
        self.lib.hostsim_derivs(C.byref(self.cfg), _abi.dptr(x), C.c_double(T), C.c_double(mu), C.c_double(xi), _abi.dptr(o))
        return dict(dF_dT=o[0:5].copy(), dF_dmu=o[5:10].copy(), s_T=o[10], s_mu=o[11], nB_T=o[12], nB_mu=o[13])

    def thermo(self, x, T, mu, xi):
        x = _abi.as_f64(x)
        o = np.zeros(17)
        self.lib.hostsim_thermo(C.byref(self.cfg), _abi.dptr(x), C.c_double(T), C.c_double(mu), C.c_double(xi),
                                _abi.dptr(o))
        return dict(omega=o[0], pressure=o[1], rho_norm=o[2], entropy=o[3], energy=o[4], rho=o[5:8].copy(),
                    masses=o[8:11].copy(), n_q=o[11:14].copy(), n_qbar=o[14:17].copy())

    def solve_points(self, T, mu, xi, seed_mode=_abi.SEED_MULTI, seeds=None):
        T = _abi.as_f64(T)
        n = T.size
        mu = _abi.as_f64(mu, n)
        xi = _abi.as_f64(xi, n)
        n_seeds = 6
        sp = None
        if seed_mode == _abi.SEED_EXPLICIT:
            seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(n, -1, 5)
            n_seeds = seeds.shape[1]
            sp = _abi.dptr(seeds)
        rec = np.zeros((n, _abi.REC_DOUBLES))
        self.lib.hostsim_solve_points(C.byref(self.cfg), C.c_int64(n), _abi.dptr(T), _abi.dptr(mu), _abi.dptr(xi),
                                      C.c_int32(seed_mode), C.c_int32(n_seeds), sp, _abi.dptr(rec))
        return rec

    def dual_branch(self, T_MeV, xi, mu_MeV):
        r = self.scan_lines(T_MeV, xi, mu_MeV, (), None, mode=2)
        return r.reshape(-1, 2, r.shape[1], _abi.REC_DOUBLES)

    def tmu_scan(self, T_MeV, xi, mu_MeV, tables=(), table_idx=None):
        return self.scan_lines(T_MeV, xi, mu_MeV, tables, table_idx, mode=1)

    def scan_lines(self, muq_MeV, xi, T_MeV, tables=(), table_idx=None, mode=0):
        muq_MeV = _abi.as_f64(muq_MeV)
        n_lines = muq_MeV.size
        xi = _abi.as_f64(xi, n_lines)
        T_MeV = _abi.as_f64(T_MeV)
        keep = []
        ctabs = (_abi.PnjlBoundary * max(1, len(tables)))()
        for i, (tt, mm, tcep) in enumerate(tables):
            tt = _abi.as_f64(tt)
            mm = _abi.as_f64(mm)
            keep += [tt, mm]
            ctabs[i] = _abi.PnjlBoundary(_abi.dptr(tt), _abi.dptr(mm), tt.size, tcep)
        if table_idx is None:
            table_idx = np.full(n_lines, -1, dtype=np.int32)
        table_idx = np.ascontiguousarray(table_idx, dtype=np.int32)
        rec = np.zeros((n_lines * (2 if mode == 2 else 1), T_MeV.size, _abi.REC_DOUBLES))
        self.lib.hostsim_scan_lines_mode(C.byref(self.cfg), C.c_int64(n_lines), _abi.dptr(muq_MeV), _abi.dptr(xi),
                                         _abi.iptr(table_idx), C.c_int32(T_MeV.size), _abi.dptr(T_MeV),
                                         C.c_int32(len(tables)), ctabs, _abi.dptr(rec), C.c_int32(mode))
        return rec

    def couplings(self, T, mu, m_u, m_s, Phi, Phib, nodes, weights):
        arrs = [_abi.as_f64(a) for a in (T, mu, m_u, m_s, Phi, Phib)]
        n = arrs[0].size
        nodes, weights = _abi.as_f64(nodes), _abi.as_f64(weights)
        aux = np.zeros((n, 16))
        self.lib.hostsim_couplings(C.byref(self.cfg), C.c_int64(n), *[_abi.dptr(a) for a in arrs], C.c_int32(nodes.size),
                                   _abi.dptr(nodes), _abi.dptr(weights), _abi.dptr(aux))
        return aux
