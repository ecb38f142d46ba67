"""`solve` / `solve_multi` / `SolverResult` — host-side mirror of src/pnjl/solver/ImplicitSolver.jl:178-328,
:532-559 on top of the C ABI.  Every numerical step runs in libpnjl_b200.so on the GPU; this file only
marshals arguments, chooses the seed exactly like the reference's `get_seed`, and rebuilds `SolverResult`.

The batched entry points (`solve_batch`, scan.run_scan) are what the hot path is for; the scalar
`solve()` exists so that reference call sites (`PNJL.solve(FixedMu(), T_fm, mu_fm; xi, seed_strategy,
p_num, t_num, iterations)`) keep working unchanged.
"""
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import _abi as A
from ._lib import Engine
from .seeds import (ContinuitySeed, DefaultSeed, FixedMu, MultiSeed, PhaseAwareContinuitySeed, SeedStrategy,
                    get_all_seeds, get_seed)

DEFAULT_MOMENTUM_COUNT = 64   # GaussLegendre.jl:121
DEFAULT_THETA_COUNT = 8       # GaussLegendre.jl:122


@dataclass
class SolverResult:
    """ImplicitSolver.jl:178-193 (same field names) + the densities and flags the scan needs."""
    mode: object
    converged: bool
    solution: List[float]
    x_state: Tuple[float, ...]
    mu_vec: Tuple[float, float, float]
    omega: float
    pressure: float
    rho_norm: float
    entropy: float
    energy: float
    masses: Tuple[float, float, float]
    iterations: int
    residual_norm: float
    xi: float
    # extras (not in the reference struct)
    n_quark: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    n_antiquark: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    status: int = 0

    @staticmethod
    def from_record(rec, mode=None):
        rec = np.asarray(rec)
        st = int(rec[A.REC_STATUS])
        mu = float(rec[A.REC_MU])
        x = tuple(float(v) for v in rec[A.REC_X:A.REC_X + 5])
        return SolverResult(mode if mode is not None else FixedMu(), bool(st & A.ST_CONVERGED), list(x), x,
                            (mu, mu, mu), float(rec[A.REC_OMEGA]), float(rec[A.REC_PRESSURE]),
                            float(rec[A.REC_RHO_NORM]), float(rec[A.REC_ENTROPY]), float(rec[A.REC_ENERGY]),
                            tuple(float(v) for v in rec[A.REC_MASS:A.REC_MASS + 3]), int(rec[A.REC_ITER]),
                            float(rec[A.REC_RESNORM]), float(rec[A.REC_XI]),
                            tuple(float(v) for v in rec[A.REC_NQ:A.REC_NQ + 3]),
                            tuple(float(v) for v in rec[A.REC_NQBAR:A.REC_NQBAR + 3]), st)


_ENGINES = {}


def _engine(p_num, t_num, iterations, trust_region_fallback, auto_multiseed_fallback, residual_norm_max):
    key = (int(p_num), int(t_num), int(iterations), bool(trust_region_fallback), bool(auto_multiseed_fallback),
           float(residual_norm_max))
    e = _ENGINES.get(key)
    if e is None:
        e = Engine(p_num=key[0], t_num=key[1], max_iter=key[2], trust_region_fallback=key[3],
                   auto_multiseed_fallback=key[4], residual_norm_max=key[5])
        _ENGINES[key] = e
    return e


def _check_kwargs(nlsolve_method, fallback_method, physicality_check):
    if nlsolve_method != "newton" or fallback_method != "trust_region":
        raise NotImplementedError("the accelerated path implements primary :newton with :trust_region fallback "
                                  "(the only combination the scan uses, ImplicitSolver.jl:216-219)")
    if physicality_check is not None:
        raise NotImplementedError("custom physicality_check callbacks cannot run on the device; the default "
                                  "criterion (ImplicitSolver.jl:50-60) is built in")


def solve(mode, T_fm, mu_fm, *, xi=0.0, seed_strategy: SeedStrategy = None, p_num=DEFAULT_MOMENTUM_COUNT,
          t_num=DEFAULT_THETA_COUNT, nlsolve_method="newton", trust_region_fallback=True,
          auto_multiseed_fallback=True, fallback_method="trust_region", physicality_check=None,
          residual_norm_max=1e-6, iterations=1000) -> SolverResult:
    """solve(::FixedMu, T_fm, μ_fm; ...) — ImplicitSolver.jl:211-328."""
    if not isinstance(mode, FixedMu):
        raise NotImplementedError("only FixedMu is on the accelerated path")
    _check_kwargs(nlsolve_method, fallback_method, physicality_check)
    s = seed_strategy if seed_strategy is not None else DefaultSeed()
    if isinstance(s, PhaseAwareContinuitySeed) and s.bootstrap_multiseed and s.previous_solution is None:
        return solve_multi(mode, T_fm, mu_fm, seed_strategy=s.bootstrap_strategy, xi=xi, p_num=p_num, t_num=t_num,
                           trust_region_fallback=trust_region_fallback, residual_norm_max=residual_norm_max,
                           iterations=iterations)                                      # :225-243
    if isinstance(s, MultiSeed):
        return solve_multi(mode, T_fm, mu_fm, seed_strategy=s, xi=xi, p_num=p_num, t_num=t_num,
                           trust_region_fallback=trust_region_fallback, residual_norm_max=residual_norm_max,
                           iterations=iterations)                                      # :246-259
    e = _engine(p_num, t_num, iterations, trust_region_fallback, auto_multiseed_fallback, residual_norm_max)
    x0 = get_seed(s, [float(T_fm), float(mu_fm)], mode)                                # :265-267
    rec = e.solve_points([T_fm], [mu_fm], [xi], A.SEED_EXPLICIT, np.asarray(x0, dtype=np.float64).reshape(1, 1, 5))
    return SolverResult.from_record(rec[0], mode)


def solve_multi(mode, T_fm, mu_fm, *, seed_strategy: MultiSeed = None, xi=0.0, p_num=DEFAULT_MOMENTUM_COUNT,
                t_num=DEFAULT_THETA_COUNT, nlsolve_method="newton", trust_region_fallback=True,
                residual_norm_max=1e-6, iterations=1000, **_ignored) -> SolverResult:
    """solve_multi(::FixedMu, ...) — ImplicitSolver.jl:532-559.  Raises RuntimeError when no candidate
    converges, like the reference's `error("All seeds failed ...")` (:553,:556)."""
    if not isinstance(mode, FixedMu):
        raise NotImplementedError("only FixedMu is on the accelerated path")
    s = seed_strategy if seed_strategy is not None else MultiSeed()
    e = _engine(p_num, t_num, iterations, trust_region_fallback, False, residual_norm_max)
    seeds = get_all_seeds(s, [float(T_fm), float(mu_fm)], mode)
    if len(seeds) > 6:
        raise NotImplementedError("at most 6 MultiSeed candidates per point")
    if len(seeds) == 1:   # solve_multi over one candidate: no further fallback (auto_multiseed_fallback=false)
        rec = e.solve_points([T_fm], [mu_fm], [xi], A.SEED_EXPLICIT, np.asarray(seeds).reshape(1, 1, 5))
        st = int(rec[0, A.REC_STATUS])
        if not st & A.ST_CONVERGED:
            raise RuntimeError("All seeds failed to converge to a physical solution")
        return SolverResult.from_record(rec[0], mode)
    rec = e.solve_points([T_fm], [mu_fm], [xi], A.SEED_EXPLICIT, np.asarray(seeds, dtype=np.float64).reshape(1, -1, 5))
    if int(rec[0, A.REC_STATUS]) & A.ST_ALL_SEEDS_FAILED:
        raise RuntimeError("All seeds failed to converge to a physical solution")
    return SolverResult.from_record(rec[0], mode)


def solve_batch(T_fm, mu_fm, xi=0.0, *, seed_strategy="multi", seeds=None, p_num=DEFAULT_MOMENTUM_COUNT,
                t_num=DEFAULT_THETA_COUNT, iterations=1000, trust_region_fallback=True,
                auto_multiseed_fallback=True, residual_norm_max=1e-6):
    """Batched independent solves: records [n][32].  seed_strategy: "multi" (MultiSeed()), "auto"
    (DefaultSeed(:auto)) or "explicit" (seeds[n][k][5])."""
    mode = {"multi": A.SEED_MULTI, "auto": A.SEED_AUTO, "explicit": A.SEED_EXPLICIT}[seed_strategy]
    e = _engine(p_num, t_num, iterations, trust_region_fallback, auto_multiseed_fallback, residual_norm_max)
    return e.solve_points(T_fm, mu_fm, xi, mode, seeds)
