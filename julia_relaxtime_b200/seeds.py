"""Seed strategies — host-side mirror of src/pnjl/solver/SeedStrategies.jl (FixedMu, 5-dim states).

The names, arguments and state transitions follow the reference so that code written against
`PNJL.DefaultSeed / MultiSeed / ContinuitySeed / PhaseAwareContinuitySeed` reads the same here.
The batched line scan runs the same tracker logic on the GPU (csrc/pnjl_solver.cuh); these
classes serve the point-at-a-time `solve()` API and the host tests.
"""
import math
from typing import List, Optional

from .boundary import PhaseBoundaryData, interpolate_mu_c, load_phase_boundary

# SeedStrategies.jl:56-91 (order: phi_u, phi_d, phi_s, Phi, Phibar)
HADRON_SEED_5 = [-1.84329, -1.84329, -2.22701, 1.0e-5, 4.0e-5]
MEDIUM_SEED_5 = [-1.3647, -1.3647, -2.14502, 0.10594, 0.15569]
HIGH_DENSITY_SEED_5 = [-0.21695, -0.21695, -2.01372, 0.18601, 0.22333]
HIGH_TEMP_SEED_5 = [-0.73192, -0.73192, -1.79539, 0.60532, 0.60532]
VERY_HIGH_TEMP_SEED_5 = [-0.30, -0.30, -0.90, 0.90, 0.90]
HT_GUESS_0p8_SEED_5 = [-0.50, -0.50, -1.20, 0.80, 0.80]
HT_GUESS_0p9_SEED_5 = [-0.30, -0.30, -0.90, 0.90, 0.90]
HT_GUESS_0p95_SEED_5 = [-0.20, -0.20, -0.70, 0.95, 0.95]
WEAK_CHIRAL_CONF_SEED_5 = [-0.50, -0.50, -1.20, 1e-3, 1e-3]
QUARK_SEED_5 = HIGH_TEMP_SEED_5

_HBARC_LITERAL = 197.327   # SeedStrategies.jl:132,216,582 use the literal, not the config value


class ConstraintMode:
    pass


class FixedMu(ConstraintMode):
    """ConstraintModes.jl:48 — theta = [T, mu], state = 5 unknowns."""

    def __repr__(self):
        return "FixedMu()"


def state_dim(mode):
    if isinstance(mode, FixedMu):
        return 5
    raise NotImplementedError("only FixedMu is on the accelerated path (SURVEY.md §8a)")


def auto_phase_hint(T_fm, mu_fm):
    """SeedStrategies.jl:131-136."""
    return "quark" if (T_fm * _HBARC_LITERAL > 150 or mu_fm * _HBARC_LITERAL > 300) else "hadron"


def extend_seed(base_seed, mode):
    """SeedStrategies.jl:155-157 (FixedMu)."""
    state_dim(mode)
    return [float(v) for v in base_seed[:5]]


class SeedStrategy:
    pass


class DefaultSeed(SeedStrategy):
    """SeedStrategies.jl:193-225."""

    def __init__(self, hadron_seed=None, quark_seed=None, phase_hint="auto"):
        self.hadron_seed = list(HADRON_SEED_5 if hadron_seed is None else hadron_seed)
        self.quark_seed = list(QUARK_SEED_5 if quark_seed is None else quark_seed)
        if phase_hint not in ("auto", "hadron", "quark"):
            raise ValueError("phase_hint must be auto, hadron or quark")
        self.phase_hint = phase_hint

    def __repr__(self):
        return "DefaultSeed(phase_hint=%s)" % self.phase_hint


def _default_get_seed(s: DefaultSeed, theta, mode):
    hint = s.phase_hint
    if hint == "auto" and len(theta) >= 2:
        hint = auto_phase_hint(theta[0], theta[1])
    elif hint == "auto":
        hint = "hadron"
    if hint == "quark":
        if len(theta) >= 1:
            base = VERY_HIGH_TEMP_SEED_5 if theta[0] * _HBARC_LITERAL >= 300.0 else s.quark_seed
        else:
            base = s.quark_seed
    else:
        base = s.hadron_seed
    return extend_seed(base, mode)


def default_omega_selector(results):
    """SeedStrategies.jl:236-240 made deterministic: argmin Omega over converged candidates; candidates within
    1e-12 * max(1, |Omega_min|) tie and the first (lowest seed index) wins."""
    conv = [r for r in results if r.converged]
    if not conv:
        return results[0]
    omin = min(r.omega for r in conv)
    tol = 1e-12 * max(1.0, abs(omin))
    for r in conv:
        if r.omega <= omin + tol:
            return r
    return conv[0]


class MultiSeed(SeedStrategy):
    """SeedStrategies.jl:251-284: six candidates, argmin-Omega selector."""

    def __init__(self, selector=default_omega_selector, candidates=None):
        self.candidates: List[SeedStrategy] = candidates if candidates is not None else [
            DefaultSeed(phase_hint="hadron"),
            DefaultSeed(phase_hint="quark"),
            DefaultSeed(WEAK_CHIRAL_CONF_SEED_5, WEAK_CHIRAL_CONF_SEED_5, "hadron"),
            DefaultSeed(HT_GUESS_0p8_SEED_5, HT_GUESS_0p8_SEED_5, "hadron"),
            DefaultSeed(HT_GUESS_0p9_SEED_5, HT_GUESS_0p9_SEED_5, "hadron"),
            DefaultSeed(HT_GUESS_0p95_SEED_5, HT_GUESS_0p95_SEED_5, "hadron"),
        ]
        self.selector = selector

    @property
    def is_builtin(self):
        return self.selector is default_omega_selector and len(self.candidates) == 6 and not hasattr(self, "_custom")

    def __repr__(self):
        return "MultiSeed(%d candidates)" % len(self.candidates)


def get_all_seeds(s: MultiSeed, theta, mode):
    """SeedStrategies.jl:282."""
    return [get_seed(c, theta, mode) for c in s.candidates]


class ContinuitySeed(SeedStrategy):
    """SeedStrategies.jl:303-345."""

    def __init__(self, fallback: Optional[SeedStrategy] = None):
        self.previous_solution: Optional[List[float]] = None
        self.fallback = fallback if fallback is not None else DefaultSeed()

    def __repr__(self):
        return "ContinuitySeed(has_previous=%s)" % (self.previous_solution is not None)


class PhaseAwareContinuitySeed(SeedStrategy):
    """SeedStrategies.jl:679-888.  previous_phase ∈ {hadron, quark, crossover, unknown}."""

    def __init__(self, xi=None, bootstrap_multiseed=False, bootstrap_strategy=None, **paths):
        self.boundary_data: Optional[PhaseBoundaryData] = None
        if xi is not None:
            try:
                self.boundary_data = load_phase_boundary(xi, **paths)
            except Exception:   # the reference warns and continues with plain continuity (:744-749)
                self.boundary_data = None
        self.previous_solution: Optional[List[float]] = None
        self.previous_phase = "unknown"
        self.hadron_seed = list(HADRON_SEED_5)
        self.quark_seed = list(QUARK_SEED_5)
        self.bootstrap_multiseed = bool(bootstrap_multiseed)
        self.bootstrap_strategy = bootstrap_strategy if bootstrap_strategy is not None else MultiSeed()
        self.fallback = DefaultSeed(phase_hint="auto")

    def __repr__(self):
        return "PhaseAwareContinuitySeed(data=%s, prev=%s, phase=%s, bootstrap_multiseed=%s)" % (
            self.boundary_data is not None, self.previous_solution is not None, self.previous_phase,
            self.bootstrap_multiseed)


def _get_current_phase(s: PhaseAwareContinuitySeed, T_MeV, mu_MeV):
    """SeedStrategies.jl:762-782."""
    if s.boundary_data is None:
        return "unknown"
    d = s.boundary_data
    if not math.isnan(d.T_CEP) and T_MeV > d.T_CEP:
        return "crossover"
    mu_c = interpolate_mu_c(d, T_MeV)
    if math.isnan(mu_c):
        return "unknown"
    return "hadron" if mu_MeV < mu_c else "quark"


def _is_phase_transition(prev_phase, curr_phase):
    """SeedStrategies.jl:789-793."""
    return (prev_phase == "hadron" and curr_phase == "quark") or (prev_phase == "quark" and curr_phase == "hadron")


def get_seed(s: SeedStrategy, theta, mode=FixedMu()):
    """get_seed(strategy, θ=[T_fm, μ_fm], mode) — one method per strategy like the reference's multiple dispatch."""
    if isinstance(s, DefaultSeed):
        return _default_get_seed(s, theta, mode)
    if isinstance(s, MultiSeed):
        return get_seed(s.candidates[0], theta, mode)                      # :271-275
    if isinstance(s, ContinuitySeed):
        if s.previous_solution is not None:                                # :314-325
            if len(s.previous_solution) == state_dim(mode):
                return list(s.previous_solution)
            if len(s.previous_solution) >= 5:
                return extend_seed(s.previous_solution, mode)
        return get_seed(s.fallback, theta, mode)
    if isinstance(s, PhaseAwareContinuitySeed):                            # :795-839
        T_fm = theta[0]
        mu_fm = theta[1] if len(theta) >= 2 else 0.0
        cur = _get_current_phase(s, T_fm * _HBARC_LITERAL, mu_fm * _HBARC_LITERAL)
        if s.previous_solution is None:
            if cur == "hadron":
                return extend_seed(s.hadron_seed, mode)
            if cur == "quark":
                return extend_seed(s.quark_seed, mode)
            return get_seed(s.fallback, theta, mode)
        if _is_phase_transition(s.previous_phase, cur):
            return extend_seed(s.hadron_seed if cur == "hadron" else s.quark_seed, mode)
        if len(s.previous_solution) == state_dim(mode):
            return list(s.previous_solution)
        if len(s.previous_solution) >= 5:
            return extend_seed(s.previous_solution, mode)
        return get_seed(s.fallback, theta, mode)
    raise TypeError("unknown seed strategy %r" % (s,))


def update_(s, solution, T_MeV=None, mu_MeV=None):
    """`update!` (:327-330, :851-865).  With (T_MeV, mu_MeV) the phase tag is refreshed, without it is kept."""
    s.previous_solution = [float(v) for v in solution]
    if isinstance(s, PhaseAwareContinuitySeed) and T_MeV is not None and mu_MeV is not None:
        s.previous_phase = _get_current_phase(s, T_MeV, mu_MeV)
    return s


def reset_(s):
    """`reset!` (:337-340, :874-878)."""
    s.previous_solution = None
    if isinstance(s, PhaseAwareContinuitySeed):
        s.previous_phase = "unknown"
    return s


def set_phase_(s: PhaseAwareContinuitySeed, phase):
    """`set_phase!` (:885-888)."""
    s.previous_phase = phase
    return s
