// pnjl_oracle.cpp — CPU ORACLE for the batched PNJL gap-equation scan.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (julia_relaxtime_b200/) never links, imports or calls anything in oracle/.
//
// What it is: a plain C++17 / FP64 restatement of the reference algorithm of
// w5851/Julia_RelaxTime for the path  PNJL.solve(FixedMu(), T, mu; xi, seed_strategy, ...)
// as driven by scripts/relaxtime/run_gap_transport_scan.jl.  It follows the reference
// the way the reference computes: Omega is written once, generically over the number
// type, and F = grad_x P and J = Hess_x P come from forward-mode dual numbers
// (Dual<5,double> and Dual<5,Dual<5,double>>), exactly like ForwardDiff.gradient nested
// inside NLsolve's autodiff=:forward Jacobian.  The GPU product uses hand-derived analytic
// derivatives instead, so the two are independent derivations of the same numbers.
//
// Third-party arithmetic that is NOT under /root/reference (Manifest.toml pins):
//   NLsolve 4.5.1      newton_ / trust_region_ / dogleg! / assess_convergence — restated
//                      below from the package's published algorithm.
//   ForwardDiff 1.3.0  dual-number rules (value-part comparisons, max/abs rules).
//   FastGaussQuadrature 1.1.0  gausslegendre(n) — restated as Newton on the Legendre
//                      three-term recurrence (long double), agrees with numpy to 1e-15.
//
// Parity pinning: tests/test_oracle_golden.py checks this file against the reference's
// committed golden CSV (406 rows, data/outputs/results/relaxtime/gap_transport_scan_xi-0p6to0p6.csv,
// copied to tests/golden/ by tests/golden/make_fixtures.py).  MultiSeed tie-breaks between
// same-branch candidates are round-off-defined in the reference; here the
// rule is made deterministic (lowest seed index within omega_tie_rel) and shared with the GPU.
//
// Reference citations are given per function as  file:line  relative to /root/reference.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kPolyakovEps = 1e-16;  // Integrals.jl:152

// ---------------------------------------------------------------------------------------
// Forward-mode dual numbers (restates ForwardDiff.Dual semantics used on this path)
// ---------------------------------------------------------------------------------------
template <int N, class T>
struct Dual {
    T v;
    T d[N];
};

inline double prim(double x) { return x; }
template <int N, class T>
inline double prim(const Dual<N, T>& x) { return prim(x.v); }

template <class T> struct Zero { static T make() { return T(0); } };
template <int N, class T> struct Zero<Dual<N, T>> {
    static Dual<N, T> make() {
        Dual<N, T> r; r.v = Zero<T>::make();
        for (int i = 0; i < N; ++i) r.d[i] = Zero<T>::make();
        return r;
    }
};
template <class T> inline T from_double(double c) { T r = Zero<T>::make(); return r + c; }

#define DUAL_TMPL template <int N, class T>
#define DUALT Dual<N, T>

DUAL_TMPL inline DUALT operator-(const DUALT& a) { DUALT r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
DUAL_TMPL inline DUALT operator+(const DUALT& a, const DUALT& b) { DUALT r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
DUAL_TMPL inline DUALT operator-(const DUALT& a, const DUALT& b) { DUALT r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
DUAL_TMPL inline DUALT operator*(const DUALT& a, const DUALT& b) { DUALT r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
DUAL_TMPL inline DUALT operator/(const DUALT& a, const DUALT& b) {
    DUALT r; r.v = a.v / b.v;
    T inv = 1.0 / b.v;
    for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
DUAL_TMPL inline DUALT operator+(const DUALT& a, double c) { DUALT r = a; r.v = a.v + c; return r; }
DUAL_TMPL inline DUALT operator+(double c, const DUALT& a) { DUALT r = a; r.v = c + a.v; return r; }
DUAL_TMPL inline DUALT operator-(const DUALT& a, double c) { DUALT r = a; r.v = a.v - c; return r; }
DUAL_TMPL inline DUALT operator-(double c, const DUALT& a) { DUALT r; r.v = c - a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
DUAL_TMPL inline DUALT operator*(const DUALT& a, double c) { DUALT r; r.v = a.v * c; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c; return r; }
DUAL_TMPL inline DUALT operator*(double c, const DUALT& a) { DUALT r; r.v = c * a.v; for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
DUAL_TMPL inline DUALT operator/(const DUALT& a, double c) { DUALT r; r.v = a.v / c; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / c; return r; }
DUAL_TMPL inline DUALT operator/(double c, const DUALT& a) {
    DUALT r; r.v = c / a.v;
    T k = -(r.v / a.v);
    for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i];
    return r;
}

inline double d_exp(double x) { return std::exp(x); }
inline double d_log(double x) { return std::log(x); }
inline double d_sqrt(double x) { return std::sqrt(x); }
inline double d_abs(double x) { return std::fabs(x); }
DUAL_TMPL inline DUALT d_exp(const DUALT& a) { DUALT r; r.v = d_exp(a.v); for (int i = 0; i < N; ++i) r.d[i] = r.v * a.d[i]; return r; }
DUAL_TMPL inline DUALT d_log(const DUALT& a) { DUALT r; r.v = d_log(a.v); T inv = 1.0 / a.v; for (int i = 0; i < N; ++i) r.d[i] = inv * a.d[i]; return r; }
DUAL_TMPL inline DUALT d_sqrt(const DUALT& a) { DUALT r; r.v = d_sqrt(a.v); T k = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = k * a.d[i]; return r; }
DUAL_TMPL inline DUALT d_abs(const DUALT& a) { return prim(a) < 0 ? -a : (prim(a) > 0 ? a : a * 0.0); }

// max(x, c) with a Float64 constant: DiffRules gives d/dx = (c > x) ? 0 : 1.
inline double d_maxc(double x, double c) { return c > x ? c : x; }
DUAL_TMPL inline DUALT d_maxc(const DUALT& a, double c) {
    if (c > prim(a)) return from_double<DUALT>(c);
    return a;
}

// ---------------------------------------------------------------------------------------
// Model constants — src/Constants_PNJL.jl:82-103, config/pnjl/default.toml
// ---------------------------------------------------------------------------------------
struct Consts {
    double hbarc, Lambda, m_ud0, m_s0, G, K, T0, a0, a1, a2, b3, rho0;
    int Nc;
};

// Mesh — Integrals.jl:87-96 (build_nodes): column-major (p fastest), coef = w_p * (2 w_c) * p^2 / (2π)^2
struct Mesh {
    int n = 0;
    std::vector<double> p, c, coef;
};

// ---------------------------------------------------------------------------------------
// Omega and friends, generic over the number type
// ---------------------------------------------------------------------------------------

// Integrals.jl:159-162
template <class R>
inline R safe_log(const R& x) {
    if (prim(x) <= 0) return from_double<R>(std::log(kPolyakovEps));
    if (prim(x) < kPolyakovEps) return from_double<R>(std::log(kPolyakovEps));
    return d_log(x);
}

// Integrals.jl:123-132
template <class R>
inline R vacuum_integral(const R& mass, const Consts& k) {
    const double L = k.Lambda;
    R mass_abs = d_abs(mass);
    R mass_safe = mass_abs + 1e-12;
    R m2 = mass_safe * mass_safe;
    R sqrt_term = d_sqrt(L * L + m2);
    R poly_part = L * sqrt_term * (2 * (L * L) + m2);
    R log_term = (m2 * m2) * d_log((L + sqrt_term) / mass_safe);
    return (poly_part - log_term) / (16 * (kPi * kPi));
}

// Integrals.jl:193-238.  E, mu, T, Phi, Phib may be double or dual in any mix in the
// reference; here callers promote everything that is differentiated to the same R and
// leave the rest double (operators above cover R∘double).
template <class RE, class RM, class RT, class RP>
inline auto calculate_log_term(const RE& E, const RM& mu, const RT& T, const RP& Phi, const RP& Phib)
    -> decltype((E - mu) * (1.0 / T) * Phi) {
    using R = decltype((E - mu) * (1.0 / T) * Phi);
    auto invT = 1.0 / T;
    auto a = -(E - mu) * invT;
    auto b = -(E + mu) * invT;
    R log_f_plus, log_f_minus;
    if (prim(a) > 0) {
        auto m_a = 3.0 * a;
        auto exp_a_m = d_exp(-2.0 * a);
        auto exp_2a_m = d_exp(-a);
        auto exp_neg_m = d_exp(-m_a);
        R term_a = exp_neg_m + 3.0 * Phi * exp_a_m + 3.0 * Phib * exp_2a_m + 1.0;
        log_f_plus = m_a + d_log(d_maxc(term_a, kPolyakovEps));
    } else {
        auto exp_a = d_exp(a);
        auto exp_2a = exp_a * exp_a;
        auto exp_3a = exp_a * exp_2a;
        R f_plus = 1.0 + 3.0 * Phi * exp_a + 3.0 * Phib * exp_2a + exp_3a;
        log_f_plus = d_log(d_maxc(f_plus, kPolyakovEps));
    }
    if (prim(b) > 0) {
        auto m_b = 3.0 * b;
        auto exp_b_m = d_exp(-2.0 * b);
        auto exp_2b_m = d_exp(-b);
        auto exp_neg_m = d_exp(-m_b);
        R term_b = exp_neg_m + 3.0 * Phib * exp_b_m + 3.0 * Phi * exp_2b_m + 1.0;
        log_f_minus = m_b + d_log(d_maxc(term_b, kPolyakovEps));
    } else {
        auto exp_b = d_exp(b);
        auto exp_2b = exp_b * exp_b;
        auto exp_3b = exp_b * exp_2b;
        R f_minus = 1.0 + 3.0 * Phib * exp_b + 3.0 * Phi * exp_2b + exp_3b;
        log_f_minus = d_log(d_maxc(f_minus, kPolyakovEps));
    }
    return log_f_plus + log_f_minus;
}

// Thermodynamics.jl:181-195 (calculate_omega) with :81-88 (masses), :112-114 (chi),
// :124-130 (U), Integrals.jl:139-145 (vacuum sum), :248-259 (thermal log sum).
// RX: type of the state x;  RM: type of mu;  RT: type of T.  At most one of them is dual.
template <class RX, class RM, class RT>
auto calculate_omega(const RX x[5], const RM mu[3], const RT& T, const Mesh& mesh, double xi, const Consts& k)
    -> decltype(x[0] * mu[0] * T) {
    using R = decltype(x[0] * mu[0] * T);
    const RX &pu = x[0], &pd = x[1], &ps = x[2], &Phi = x[3], &Phib = x[4];

    // chi = 2G Σφ² − 4K φuφdφs
    RX chi = 2 * k.G * ((pu * pu + pd * pd) + ps * ps) - 4 * k.K * ((pu * pd) * ps);

    // U(T, Φ, Φ̄)
    auto T_ratio = k.T0 / T;
    auto Ta = k.a0 + k.a1 * T_ratio + k.a2 * (T_ratio * T_ratio);
    auto Tb = k.b3 * (T_ratio * T_ratio * T_ratio);
    RX value = 1 - 6 * Phib * Phi + 4 * (Phib * Phib * Phib + Phi * Phi * Phi) - 3 * ((Phib * Phi) * (Phib * Phi));
    auto T2 = T * T;
    auto U = (T2 * T2) * (-0.5 * Ta * Phib * Phi + Tb * safe_log(value));

    // masses
    RX m[3];
    m[0] = k.m_ud0 - 4 * k.G * pu + 2 * k.K * pd * ps;
    m[1] = k.m_ud0 - 4 * k.G * pd + 2 * k.K * pu * ps;
    m[2] = k.m_s0 - 4 * k.G * ps + 2 * k.K * pu * pd;

    // vacuum: −2 Nc Σ I(Λ, M_i)
    RX vac_total = Zero<RX>::make();
    for (int i = 0; i < 3; ++i) vac_total = vac_total + vacuum_integral(m[i], k);
    RX energy_sum = (-2.0 * k.Nc) * vac_total;

    // thermal: −2T Σ_i Σ_k coef_k · logterm(E_ik, μ_i)
    R total = Zero<R>::make();
    for (int i = 0; i < 3; ++i) {
        RX m2 = m[i] * m[i];
        for (int idx = 0; idx < mesh.n; ++idx) {
            const double p = mesh.p[idx], t = mesh.c[idx];
            const double pt = p * t;
            RX E = d_sqrt(p * p + m2 + xi * (pt * pt));
            total = total + calculate_log_term(E, mu[i], T, Phi, Phib) * mesh.coef[idx];
        }
    }
    R log_sum = -2 * T * total;
    return chi + U + energy_sum + log_sum;
}

using D5 = Dual<5, double>;
using D55 = Dual<5, D5>;
using D3 = Dual<3, double>;
using D1 = Dual<1, double>;

struct Problem {
    const Consts* k;
    const Mesh* mesh;
    double T, mu, xi;
};

inline void masses_of(const double x[5], const Consts& k, double m[3]) {
    m[0] = k.m_ud0 - 4 * k.G * x[0] + 2 * k.K * x[1] * x[2];
    m[1] = k.m_ud0 - 4 * k.G * x[1] + 2 * k.K * x[0] * x[2];
    m[2] = k.m_s0 - 4 * k.G * x[2] + 2 * k.K * x[0] * x[1];
}

double eval_omega(const Problem& pb, const double x[5]) {
    double mu[3] = {pb.mu, pb.mu, pb.mu};
    return calculate_omega(x, mu, pb.T, *pb.mesh, pb.xi, *pb.k);
}

// Conditions.jl:72-81 — F = ∇ₓ P = −∇ₓ Ω via ForwardDiff.gradient
void eval_F(const Problem& pb, const double x[5], double F[5]) {
    D5 xs[5];
    for (int i = 0; i < 5; ++i) {
        xs[i].v = x[i];
        for (int j = 0; j < 5; ++j) xs[i].d[j] = (i == j) ? 1.0 : 0.0;
    }
    double mu[3] = {pb.mu, pb.mu, pb.mu};
    D5 om = calculate_omega(xs, mu, pb.T, *pb.mesh, pb.xi, *pb.k);
    for (int i = 0; i < 5; ++i) F[i] = -om.d[i];
}

// NLsolve autodiff=:forward over Conditions.jl:205-213 ⇒ nested duals; J = Hess_x P.
// F is the value part of the outer dual and is bit-identical to eval_F.
void eval_FJ(const Problem& pb, const double x[5], double F[5], double J[25]) {
    D55 xs[5];
    for (int i = 0; i < 5; ++i) {
        xs[i] = Zero<D55>::make();
        xs[i].v.v = x[i];
        xs[i].v.d[i] = 1.0;   // inner: gradient direction
        xs[i].d[i].v = 1.0;   // outer: jacobian direction
    }
    double mu[3] = {pb.mu, pb.mu, pb.mu};
    D55 om = calculate_omega(xs, mu, pb.T, *pb.mesh, pb.xi, *pb.k);
    for (int i = 0; i < 5; ++i) {
        F[i] = -om.v.d[i];
        for (int j = 0; j < 5; ++j) J[i * 5 + j] = -om.d[j].d[i];  // row-major J[i][j] = ∂F_i/∂x_j
    }
}

struct Thermo {
    double omega, pressure, rho_norm, entropy, energy;
    double rho[3];
    double masses[3];
};

// Thermodynamics.jl:215-220 (rho), :233-244 (thermo); ImplicitSolver.jl:271-277 (postprocess)
Thermo eval_thermo(const Problem& pb, const double x[5]) {
    Thermo th;
    const Consts& k = *pb.k;
    {
        D3 mu[3];
        for (int i = 0; i < 3; ++i) {
            mu[i].v = pb.mu;
            for (int j = 0; j < 3; ++j) mu[i].d[j] = (i == j) ? 1.0 : 0.0;
        }
        D3 om = calculate_omega(x, mu, pb.T, *pb.mesh, pb.xi, k);
        for (int i = 0; i < 3; ++i) th.rho[i] = -om.d[i];
    }
    th.rho_norm = ((th.rho[0] + th.rho[1]) + th.rho[2]) / (3.0 * k.rho0);
    {
        D1 T; T.v = pb.T; T.d[0] = 1.0;
        double mu[3] = {pb.mu, pb.mu, pb.mu};
        D1 om = calculate_omega(x, mu, T, *pb.mesh, pb.xi, k);
        th.entropy = -om.d[0];
    }
    th.pressure = -eval_omega(pb, x);
    th.energy = -th.pressure + ((pb.mu * th.rho[0] + pb.mu * th.rho[1]) + pb.mu * th.rho[2]) + pb.T * th.entropy;
    th.omega = -th.pressure;
    masses_of(x, k, th.masses);
    return th;
}

// QuarkDistribution.jl:14-31 / :34-51 (clamp(exp, 1e-200, 1e200)); @fastmath not modelled.
inline double quark_distribution(double E, double mu, double T, double Phi, double Phib) {
    double beta = 1 / T;
    double e1 = std::exp(-(E - mu) * beta);
    e1 = std::min(std::max(e1, 1e-200), 1e200);
    double e2 = e1 * e1, e3 = e2 * e1;
    double num = Phi * e1 + 2 * Phib * e2 + e3;
    double den = 1 + 3 * Phi * e1 + 3 * Phib * e2 + e3;
    return num / den;
}
inline double antiquark_distribution(double E, double mu, double T, double Phi, double Phib) {
    double beta = 1 / T;
    double e1 = std::exp(-(E + mu) * beta);
    e1 = std::min(std::max(e1, 1e-200), 1e200);
    double e2 = e1 * e1, e3 = e2 * e1;
    double num = Phib * e1 + 2 * Phi * e2 + e3;
    double den = 1 + 3 * Phib * e1 + 3 * Phi * e2 + e3;
    return num / den;
}

// Thermodynamics.jl:255-281 with QuarkDistribution_Aniso.jl:96-120
void number_densities(const Problem& pb, const double x[5], double nq[3], double nqb[3]) {
    const Consts& k = *pb.k;
    double m[3];
    masses_of(x, k, m);
    const double pref = 2 * k.Nc;
    for (int i = 0; i < 3; ++i) {
        double tq = 0, taq = 0;
        for (int idx = 0; idx < pb.mesh->n; ++idx) {
            double p = pb.mesh->p[idx], c = pb.mesh->c[idx], w = pb.mesh->coef[idx];
            double pc = p * c;
            double E = std::sqrt(p * p + m[i] * m[i] + pb.xi * (pc * pc));
            tq += w * pref * quark_distribution(E, pb.mu, pb.T, x[3], x[4]);
            taq += w * pref * antiquark_distribution(E, pb.mu, pb.T, x[3], x[4]);
        }
        nq[i] = tq;
        nqb[i] = taq;
    }
}

// ---------------------------------------------------------------------------------------
// Dense 5x5 helpers (Julia `A \ b` on a square Matrix{Float64} = LAPACK getrf + getrs)
// ---------------------------------------------------------------------------------------
constexpr int NX = 5;

// Returns false on an exactly-zero pivot (LinearAlgebra.SingularException).
bool lu_solve(const double A_in[25], const double b_in[5], double x[5], int n = NX) {
    double A[25];
    double b[5];
    std::memcpy(A, A_in, sizeof(double) * n * n);
    std::memcpy(b, b_in, sizeof(double) * n);
    for (int kcol = 0; kcol < n; ++kcol) {
        int piv = kcol;
        double best = std::fabs(A[kcol * n + kcol]);
        for (int i = kcol + 1; i < n; ++i) {
            double v = std::fabs(A[i * n + kcol]);
            if (v > best) { best = v; piv = i; }
        }
        if (!(best > 0.0) || !std::isfinite(best)) {
            if (best == 0.0) return false;
        }
        if (piv != kcol) {
            for (int j = 0; j < n; ++j) std::swap(A[kcol * n + j], A[piv * n + j]);
            std::swap(b[kcol], b[piv]);
        }
        double inv = 1.0 / A[kcol * n + kcol];
        for (int i = kcol + 1; i < n; ++i) {
            double l = A[i * n + kcol] * inv;
            A[i * n + kcol] = l;
            for (int j = kcol + 1; j < n; ++j) A[i * n + j] -= l * A[kcol * n + j];
            b[i] -= l * b[kcol];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * x[j];
        x[i] = s / A[i * n + i];
    }
    return true;
}

inline double norm_inf(const double* v, int n = NX) {
    double m = 0;
    for (int i = 0; i < n; ++i) {
        double a = std::fabs(v[i]);
        if (a > m || std::isnan(a)) m = a;
    }
    return m;
}
inline bool any_nan(const double* v, int n = NX) {
    for (int i = 0; i < n; ++i) if (std::isnan(v[i])) return true;
    return false;
}
inline bool all_finite(const double* v, int n = NX) {
    for (int i = 0; i < n; ++i) if (!std::isfinite(v[i])) return false;
    return true;
}

struct NLResult {
    double zero[5];
    int iterations = 0;
    double residual_norm = std::numeric_limits<double>::quiet_NaN();
    bool x_converged = false, f_converged = false;
    bool threw = false;  // IsFiniteException on the initial residual
    int n_fj = 0;        // number of Ω/Jacobian-class evaluations (for FLOP accounting)
};

struct Trace {
    double* x = nullptr;  // [cap][5]
    int cap = 0;
    int n = 0;
    void push(const double* xv) {
        if (x && n < cap) std::memcpy(x + 5 * n, xv, 5 * sizeof(double));
        ++n;
    }
};

// NLsolve 4.5.1 src/solvers/newton.jl (newton_, LineSearches.Static()), src/utils.jl
// (assess_convergence: sup-norms).  Call site: ImplicitSolver.jl:112.
NLResult nl_newton(const Problem& pb, const double x0[5], double xtol, double ftol, int iterations, Trace* tr) {
    NLResult r;
    double x[5], xold[5], f[5], J[25], p[5];
    std::memcpy(x, x0, sizeof x);
    eval_FJ(pb, x, f, J);
    r.n_fj++;
    if (tr) tr->push(x);
    if (!all_finite(f)) {  // check_isfinite → IsFiniteException
        r.threw = true;
        std::memcpy(r.zero, x, sizeof x);
        r.residual_norm = norm_inf(f);
        return r;
    }
    int it = 0;
    bool x_conv = false;
    bool f_conv = norm_inf(f) <= ftol;
    bool stopped = any_nan(x) || any_nan(f);
    bool converged = x_conv || f_conv;
    while (!stopped && !converged && it < iterations) {
        ++it;
        if (it > 1) {
            double fj[5];
            eval_FJ(pb, x, fj, J);
        }
        if (lu_solve(J, f, p)) {
            for (int i = 0; i < 5; ++i) p[i] = -p[i];
        } else {
            // singular: p solves −(JᵀJ + λI) p = Jᵀ f,  λ = 1e6·sqrt(n·eps)·‖JᵀJ‖₁
            double JtJ[25], Jtf[5];
            for (int i = 0; i < 5; ++i) {
                for (int j = 0; j < 5; ++j) {
                    double s = 0;
                    for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * J[q * 5 + j];
                    JtJ[i * 5 + j] = s;
                }
                double s = 0;
                for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * f[q];
                Jtf[i] = s;
            }
            double n1 = 0;
            for (int j = 0; j < 5; ++j) {
                double s = 0;
                for (int i = 0; i < 5; ++i) s += std::fabs(JtJ[i * 5 + j]);
                n1 = std::max(n1, s);
            }
            double lambda = 1e6 * std::sqrt(5 * std::numeric_limits<double>::epsilon()) * n1;
            for (int i = 0; i < 25; ++i) JtJ[i] = -JtJ[i];
            for (int i = 0; i < 5; ++i) JtJ[i * 5 + i] -= lambda;
            if (!lu_solve(JtJ, Jtf, p)) {
                for (int i = 0; i < 5; ++i) p[i] = std::numeric_limits<double>::quiet_NaN();
            }
        }
        std::memcpy(xold, x, sizeof x);
        for (int i = 0; i < 5; ++i) x[i] = x[i] + p[i];
        eval_F(pb, x, f);
        r.n_fj++;
        if (tr) tr->push(x);
        double dx = 0;
        for (int i = 0; i < 5; ++i) {
            double a = std::fabs(x[i] - xold[i]);
            if (a > dx || std::isnan(a)) dx = a;
        }
        x_conv = dx <= xtol;
        f_conv = norm_inf(f) <= ftol;
        stopped = any_nan(x) || any_nan(f);
        converged = x_conv || f_conv;
    }
    std::memcpy(r.zero, x, sizeof x);
    r.iterations = it;
    r.residual_norm = norm_inf(f);
    r.x_converged = x_conv;
    r.f_converged = f_conv;
    return r;
}

inline double wnorm(const double* d, const double* v) {
    double s = 0;
    for (int i = 0; i < 5; ++i) { double t = d[i] * v[i]; s += t * t; }
    return std::sqrt(s);
}

// NLsolve 4.5.1 src/solvers/trust_region.jl (dogleg!).  Singular J ⇒ SVD pseudo-inverse in the
// package; here a singular J yields a NaN Gauss-Newton step which forces the Cauchy branch
// (exact singularity never occurs on this path; noted as a deviation).
void dogleg(double p[5], const double r[5], const double d[5], const double J[25], double delta) {
    double p_i[5], p_c[5], g[5];
    bool ok = lu_solve(J, r, p_i);
    for (int i = 0; i < 5; ++i) p_i[i] = ok ? -p_i[i] : std::numeric_limits<double>::infinity();
    if (wnorm(d, p_i) <= delta) {
        std::memcpy(p, p_i, sizeof p_i);
        return;
    }
    for (int i = 0; i < 5; ++i) {
        double s = 0;
        for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * r[q];
        g[i] = s / (d[i] * d[i]);
    }
    double Jg2 = 0;
    for (int q = 0; q < 5; ++q) {
        double s = 0;
        for (int i = 0; i < 5; ++i) s += J[q * 5 + i] * g[i];
        Jg2 += s * s;
    }
    double wg = wnorm(d, g);
    double coef = -(wg * wg) / Jg2;
    for (int i = 0; i < 5; ++i) p_c[i] = coef * g[i];
    if (wnorm(d, p_c) >= delta) {
        double s = -delta / wg;
        for (int i = 0; i < 5; ++i) p[i] = g[i] * s;
        return;
    }
    double p_diff[5];
    for (int i = 0; i < 5; ++i) p_diff[i] = p_i[i] - p_c[i];
    double wd = 0;
    for (int i = 0; i < 5; ++i) wd += (d[i] * p_c[i]) * (d[i] * p_diff[i]);
    double b = 2 * wd;
    double a = wnorm(d, p_diff);
    a = a * a;
    double wc = wnorm(d, p_c);
    double tau = (-b + std::sqrt(b * b - 4 * a * (wc * wc - delta * delta))) / (2 * a);
    for (int i = 0; i < 5; ++i) p[i] = p_c[i] + tau * p_diff[i];
}

// NLsolve 4.5.1 trust_region_ (factor = 1.0, autoscale = true).  Call site: ImplicitSolver.jl:144.
NLResult nl_trust_region(const Problem& pb, const double x0[5], double xtol, double ftol, int iterations, Trace* tr) {
    NLResult res;
    double x[5], xold[5], r[5], fv[5], J[25], d[5], p[5], r_predict[5];
    std::memcpy(x, x0, sizeof x);
    eval_FJ(pb, x, fv, J);
    res.n_fj++;
    if (tr) tr->push(x);
    std::memcpy(r, fv, sizeof r);
    if (!all_finite(r)) {
        res.threw = true;
        std::memcpy(res.zero, x, sizeof x);
        res.residual_norm = norm_inf(r);
        return res;
    }
    int it = 0;
    bool x_conv = false;
    bool f_conv = norm_inf(fv) <= ftol;
    bool stopped = any_nan(x) || any_nan(fv);
    bool converged = x_conv || f_conv;
    if (converged) {
        std::memcpy(res.zero, x, sizeof x);
        res.iterations = 0;
        res.residual_norm = norm_inf(r);
        res.x_converged = x_conv;
        res.f_converged = f_conv;
        return res;
    }
    for (int j = 0; j < 5; ++j) {
        double s = 0;
        for (int i = 0; i < 5; ++i) s += J[i * 5 + j] * J[i * 5 + j];
        d[j] = std::sqrt(s);
        if (d[j] == 0.0) d[j] = 1.0;
    }
    const double factor = 1.0;
    double delta = factor * wnorm(d, x);
    if (delta == 0.0) delta = factor;
    const double eta = 1e-4;
    while (!stopped && !converged && it < iterations) {
        ++it;
        dogleg(p, r, d, J, delta);
        std::memcpy(xold, x, sizeof x);
        for (int i = 0; i < 5; ++i) x[i] += p[i];
        eval_F(pb, x, fv);
        res.n_fj++;
        if (tr) tr->push(x);
        for (int q = 0; q < 5; ++q) {
            double s = 0;
            for (int i = 0; i < 5; ++i) s += J[q * 5 + i] * p[i];
            r_predict[q] = s + r[q];
        }
        double sr = 0, sf = 0, sp = 0;
        for (int i = 0; i < 5; ++i) { sr += r[i] * r[i]; sf += fv[i] * fv[i]; sp += r_predict[i] * r_predict[i]; }
        double rho = (sr - sf) / (sr - sp);
        if (rho > eta) {
            std::memcpy(r, fv, sizeof r);
            double fj[5];
            eval_FJ(pb, x, fj, J);
            for (int j = 0; j < 5; ++j) {
                double s = 0;
                for (int i = 0; i < 5; ++i) s += J[i * 5 + j] * J[i * 5 + j];
                d[j] = std::max(0.1 * d[j], std::sqrt(s));
            }
            double dx = 0;
            for (int i = 0; i < 5; ++i) {
                double a = std::fabs(x[i] - xold[i]);
                if (a > dx || std::isnan(a)) dx = a;
            }
            x_conv = dx <= xtol;
            f_conv = norm_inf(r) <= ftol;
            converged = x_conv || f_conv;
        } else {
            for (int i = 0; i < 5; ++i) x[i] -= p[i];
            x_conv = false;
            converged = false;
        }
        if (rho < 0.1) {
            delta = delta / 2;
        } else if (rho >= 0.9) {
            delta = 2 * wnorm(d, p);
        } else if (rho >= 0.5) {
            delta = std::max(delta, 2 * wnorm(d, p));
        }
        stopped = any_nan(x) || any_nan(fv);
    }
    std::memcpy(res.zero, x, sizeof x);
    res.iterations = it;
    res.residual_norm = norm_inf(r);
    res.x_converged = x_conv;
    res.f_converged = f_conv;
    return res;
}

// ---------------------------------------------------------------------------------------
// Solve cascade — ImplicitSolver.jl
// ---------------------------------------------------------------------------------------
struct SolverOpts {
    double xtol = 1e-9, ftol = 1e-9, residual_norm_max = 1e-6, phi_tol = 1e-8;
    int max_iter = 1000;
    bool tr_fallback = true;
    bool auto_multiseed_fallback = true;
    double omega_tie_rel = 1e-12;
};

enum StatusBits : int32_t {
    ST_CONVERGED = 1,
    ST_USED_TR = 2,        // returned candidate came from the trust-region solve
    ST_TR_ATTEMPTED = 4,
    ST_USED_MULTISEED = 8, // returned candidate came out of solve_multi
    ST_SEED_SHIFT = 4,     // bits 4..6: seed index chosen by solve_multi
    ST_PHASE_SWITCH = 128, // PhaseAwareContinuitySeed re-seeded at a hadron<->quark flip
    ST_NONFINITE = 256,    // initial residual non-finite (IsFiniteException)
    ST_ALL_SEEDS_FAILED = 512,
    ST_PROMOTED = 1024,     // TmuScan: residual <= 1e-4 force-marked converged (TmuScan.jl:428-458)
    ST_REFINED = 2048,      // TmuScan: re-solved from a near-converged state (TmuScan.jl:391-408)
    ST_CAND_SHIFT = 12,     // bits 12..13: index of the TmuScan seed candidate that succeeded
    ST_NO_RESULT = 16384,   // TmuScan: every candidate failed (the CSV row is all NaN)
};

struct Candidate {
    bool phys = false;
    double x[5];
    Thermo th;
};

struct PointResult {
    bool converged = false;
    double x[5];
    Thermo th;
    int iterations = 0;
    double residual_norm = std::numeric_limits<double>::quiet_NaN();
    int32_t status = 0;
    int n_fj = 0;  // total Ω-gradient/Jacobian-class evaluations spent on this point
    int n_thermo = 0;
};

// ImplicitSolver.jl:50-60
bool is_physical(const double x[5], const double m[3], double phi_tol) {
    double Phi = x[3], Phib = x[4];
    if (!(std::isfinite(Phi) && std::isfinite(Phib) && (-phi_tol <= Phi && Phi <= 1 + phi_tol) &&
          (-phi_tol <= Phib && Phib <= 1 + phi_tol)))
        return false;
    for (int i = 0; i < 3; ++i)
        if (!std::isfinite(m[i]) || m[i] <= 0.0) return false;
    return true;
}

// ImplicitSolver.jl:66-70 (+ :62-64)
Candidate postprocess(const Problem& pb, const double x[5], const SolverOpts& o) {
    Candidate c;
    std::memcpy(c.x, x, sizeof c.x);
    c.th = eval_thermo(pb, x);
    bool fin = std::isfinite(c.th.omega) && std::isfinite(c.th.pressure) && std::isfinite(c.th.rho_norm) &&
               std::isfinite(c.th.entropy) && std::isfinite(c.th.energy);
    c.phys = is_physical(x, c.th.masses, o.phi_tol) && fin;
    return c;
}

// ImplicitSolver.jl:211-328 with a fixed seed x0 (the part after get_seed), i.e.
// _nlsolve_with_tr_fallback (:103-151) + _choose_candidate (:72-101) + `converged` (:287).
PointResult solve_single(const Problem& pb, const double x0[5], const SolverOpts& o, Trace* tr = nullptr) {
    PointResult out;
    NLResult pr = nl_newton(pb, x0, o.xtol, o.ftol, o.max_iter, tr);
    out.n_fj += pr.n_fj;
    if (pr.threw) {
        // IsFiniteException propagates out of solve(); solve_multi catches it per seed.
        out.status |= ST_NONFINITE;
        std::memcpy(out.x, pr.zero, sizeof out.x);
        out.th = Thermo{};
        masses_of(out.x, *pb.k, out.th.masses);
        out.th.omega = out.th.pressure = out.th.rho_norm = out.th.entropy = out.th.energy =
            std::numeric_limits<double>::quiet_NaN();
        out.residual_norm = pr.residual_norm;
        return out;
    }
    Candidate pc = postprocess(pb, pr.zero, o);
    out.n_thermo++;
    bool need_fallback = o.tr_fallback && (!pr.f_converged || !std::isfinite(pr.residual_norm) ||
                                           pr.residual_norm > o.residual_norm_max || !pc.phys);
    const NLResult* res = &pr;
    const Candidate* cand = &pc;
    NLResult fr;
    Candidate fc;
    bool used_tr = false;
    if (need_fallback) {
        out.status |= ST_TR_ATTEMPTED;
        fr = nl_trust_region(pb, x0, o.xtol, o.ftol, o.max_iter, nullptr);
        out.n_fj += fr.n_fj;
        if (!fr.threw) {
            fc = postprocess(pb, fr.zero, o);
            out.n_thermo++;
            const double rmax = o.residual_norm_max;
            bool pg = pr.f_converged && std::isfinite(pr.residual_norm) && pr.residual_norm <= rmax && pc.phys;
            bool fg = fr.f_converged && std::isfinite(fr.residual_norm) && fr.residual_norm <= rmax && fc.phys;
            bool take_f;
            if (fg && !pg) take_f = true;
            else if (pg && !fg) take_f = false;
            else if (fg && pg) {
                if (fc.th.omega < pc.th.omega) take_f = true;
                else if (fc.th.omega > pc.th.omega) take_f = false;
                else take_f = fr.residual_norm < pr.residual_norm;
            } else if (fr.f_converged && !pr.f_converged) take_f = true;
            else if (pr.f_converged && !fr.f_converged) take_f = false;
            else if (std::isfinite(fr.residual_norm) && std::isfinite(pr.residual_norm))
                take_f = fr.residual_norm < pr.residual_norm;
            else take_f = false;
            if (take_f) { res = &fr; cand = &fc; used_tr = true; }
        }
    }
    out.converged = res->f_converged && cand->phys && std::isfinite(res->residual_norm) &&
                    res->residual_norm <= o.residual_norm_max;
    std::memcpy(out.x, cand->x, sizeof out.x);
    out.th = cand->th;
    out.iterations = res->iterations;
    out.residual_norm = res->residual_norm;
    if (out.converged) out.status |= ST_CONVERGED;
    if (used_tr) out.status |= ST_USED_TR;
    return out;
}

// Seeds — SeedStrategies.jl:56-91
const double HADRON_SEED[5] = {-1.84329, -1.84329, -2.22701, 1.0e-5, 4.0e-5};
const double HIGH_TEMP_SEED[5] = {-0.73192, -0.73192, -1.79539, 0.60532, 0.60532};  // == QUARK_SEED_5
const double VERY_HIGH_TEMP_SEED[5] = {-0.30, -0.30, -0.90, 0.90, 0.90};
const double HT_0p8_SEED[5] = {-0.50, -0.50, -1.20, 0.80, 0.80};
const double HT_0p9_SEED[5] = {-0.30, -0.30, -0.90, 0.90, 0.90};
const double HT_0p95_SEED[5] = {-0.20, -0.20, -0.70, 0.95, 0.95};
const double WEAK_CHIRAL_CONF_SEED[5] = {-0.50, -0.50, -1.20, 1e-3, 1e-3};

// SeedStrategies.jl:193-225 — DefaultSeed(phase_hint) get_seed.  hint: 0 hadron, 1 quark, 2 auto.
void default_seed(int hint, double T_fm, double mu_fm, double out[5]) {
    if (hint == 2) {
        const double hc = 197.327;  // literal in SeedStrategies.jl:132
        double T_mev = T_fm * hc, mu_mev = mu_fm * hc;
        hint = (T_mev > 150 || mu_mev > 300) ? 1 : 0;
    }
    const double* base;
    if (hint == 1) {
        double T_mev = T_fm * 197.327;  // SeedStrategies.jl:216
        base = (T_mev >= 300.0) ? VERY_HIGH_TEMP_SEED : HIGH_TEMP_SEED;
    } else {
        base = HADRON_SEED;
    }
    std::memcpy(out, base, 5 * sizeof(double));
}

// SeedStrategies.jl:251-284 — the six MultiSeed candidates, in order.
void multiseed_seeds(double T_fm, double mu_fm, double seeds[6][5]) {
    default_seed(0, T_fm, mu_fm, seeds[0]);
    default_seed(1, T_fm, mu_fm, seeds[1]);
    std::memcpy(seeds[2], WEAK_CHIRAL_CONF_SEED, sizeof(double) * 5);
    std::memcpy(seeds[3], HT_0p8_SEED, sizeof(double) * 5);
    std::memcpy(seeds[4], HT_0p9_SEED, sizeof(double) * 5);
    std::memcpy(seeds[5], HT_0p95_SEED, sizeof(double) * 5);
}

// ImplicitSolver.jl:532-559 (solve_multi) + SeedStrategies.jl:236-240 (argmin Ω).
// Deterministic tie rule: among converged candidates take the smallest Ω;
// candidates within omega_tie_rel·max(1,|Ω_min|) of it tie, lowest seed index wins.
// Returns false when no seed converged (the reference throws `error(...)`).
bool solve_multi(const Problem& pb, const double (*seeds)[5], int n_seeds, const SolverOpts& o, PointResult& best,
                 double* per_seed /* optional [n_seeds][8]: x[5], omega, converged, iterations */ = nullptr) {
    std::vector<PointResult> rs(n_seeds);
    int total_fj = 0, total_th = 0;
    double omin = std::numeric_limits<double>::infinity();
    bool any = false;
    for (int s = 0; s < n_seeds; ++s) {
        rs[s] = solve_single(pb, seeds[s], o);
        total_fj += rs[s].n_fj;
        total_th += rs[s].n_thermo;
        if (per_seed) {
            for (int i = 0; i < 5; ++i) per_seed[s * 8 + i] = rs[s].x[i];
            per_seed[s * 8 + 5] = rs[s].th.omega;
            per_seed[s * 8 + 6] = rs[s].converged ? 1.0 : 0.0;
            per_seed[s * 8 + 7] = rs[s].iterations;
        }
        if (rs[s].converged) {
            any = true;
            if (rs[s].th.omega < omin) omin = rs[s].th.omega;
        }
    }
    if (!any) {
        best = rs[0];
        best.converged = false;
        best.status = (best.status & ~ST_CONVERGED) | ST_ALL_SEEDS_FAILED | ST_USED_MULTISEED;
        best.n_fj = total_fj;
        best.n_thermo = total_th;
        return false;
    }
    double tol = o.omega_tie_rel * std::max(1.0, std::fabs(omin));
    for (int s = 0; s < n_seeds; ++s) {
        if (rs[s].converged && rs[s].th.omega <= omin + tol) {
            best = rs[s];
            best.status |= ST_USED_MULTISEED | (s << ST_SEED_SHIFT);
            best.n_fj = total_fj;
            best.n_thermo = total_th;
            return true;
        }
    }
    return false;  // unreachable
}

bool solve_multiseed(const Problem& pb, const SolverOpts& o, PointResult& best, double* per_seed = nullptr) {
    double seeds[6][5];
    multiseed_seeds(pb.T, pb.mu, seeds);
    return solve_multi(pb, seeds, 6, o, best, per_seed);
}

// ImplicitSolver.jl:211-328 — solve() for a seed x0 obtained from a non-MultiSeed strategy,
// including the automatic MultiSeed fallback (:306-327; exceptions there return `single`).
PointResult solve_with_fallback(const Problem& pb, const double x0[5], const SolverOpts& o) {
    PointResult single = solve_single(pb, x0, o);
    if (single.converged || !o.auto_multiseed_fallback || (single.status & ST_NONFINITE)) return single;
    PointResult multi;
    if (solve_multiseed(pb, o, multi)) {
        multi.n_fj += single.n_fj;
        multi.n_thermo += single.n_thermo;
        multi.status |= (single.status & ST_TR_ATTEMPTED);
        return multi;
    }
    single.n_fj += multi.n_fj;
    single.n_thermo += multi.n_thermo;
    single.status |= ST_ALL_SEEDS_FAILED;
    return single;
}

// ---------------------------------------------------------------------------------------
// PhaseAwareContinuitySeed — SeedStrategies.jl:365-475 (table), :679-888 (tracker)
// ---------------------------------------------------------------------------------------
struct PhaseTable {
    std::vector<double> T, mu;  // MeV, sorted by T
    double T_CEP = std::numeric_limits<double>::quiet_NaN();
};

enum Phase { PH_UNKNOWN = 0, PH_HADRON = 1, PH_QUARK = 2, PH_CROSSOVER = 3 };

// SeedStrategies.jl:446-475
double interpolate_mu_c(const PhaseTable& t, double T_MeV) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (!std::isnan(t.T_CEP) && T_MeV > t.T_CEP) return nan;
    if (t.T.empty()) return nan;
    const size_t n = t.T.size();
    if (T_MeV <= t.T[0]) return t.mu[0];
    if (T_MeV >= t.T[n - 1]) return t.mu[n - 1];
    for (size_t i = 0; i + 1 < n; ++i) {
        if (t.T[i] <= T_MeV && T_MeV <= t.T[i + 1]) {
            double w = (T_MeV - t.T[i]) / (t.T[i + 1] - t.T[i]);
            return t.mu[i] + w * (t.mu[i + 1] - t.mu[i]);
        }
    }
    return nan;
}

// SeedStrategies.jl:762-782
Phase current_phase(const PhaseTable& t, double T_MeV, double mu_MeV) {
    if (!std::isnan(t.T_CEP) && T_MeV > t.T_CEP) return PH_CROSSOVER;
    double mu_c = interpolate_mu_c(t, T_MeV);
    if (std::isnan(mu_c)) return PH_UNKNOWN;
    return mu_MeV < mu_c ? PH_HADRON : PH_QUARK;
}

struct Tracker {
    const PhaseTable* table;
    bool has_prev = false;
    double prev[5];
    Phase prev_phase = PH_UNKNOWN;
};

// SeedStrategies.jl:795-839.  Returns true if a phase-switch re-seed happened.
bool tracker_get_seed(const Tracker& tk, double T_fm, double mu_fm, double out[5]) {
    const double hc = 197.327;  // SeedStrategies.jl:582
    double T_MeV = T_fm * hc, mu_MeV = mu_fm * hc;
    Phase cur = current_phase(*tk.table, T_MeV, mu_MeV);
    if (!tk.has_prev) {
        if (cur == PH_HADRON) std::memcpy(out, HADRON_SEED, 40);
        else if (cur == PH_QUARK) std::memcpy(out, HIGH_TEMP_SEED, 40);
        else default_seed(2, T_fm, mu_fm, out);
        return false;
    }
    bool flip = (tk.prev_phase == PH_HADRON && cur == PH_QUARK) || (tk.prev_phase == PH_QUARK && cur == PH_HADRON);
    if (flip) {
        std::memcpy(out, cur == PH_HADRON ? HADRON_SEED : HIGH_TEMP_SEED, 40);
        return true;
    }
    std::memcpy(out, tk.prev, 40);
    return false;
}

// SeedStrategies.jl:851-856
void tracker_update(Tracker& tk, const double x[5], double T_MeV, double mu_MeV) {
    std::memcpy(tk.prev, x, 40);
    tk.has_prev = true;
    tk.prev_phase = current_phase(*tk.table, T_MeV, mu_MeV);
}

// Integrals.jl:67-96
Mesh build_mesh(int p_num, int t_num, const double* p_nodes, const double* p_w, const double* c_nodes,
                const double* c_w_raw) {
    Mesh m;
    m.n = p_num * t_num;
    m.p.resize(m.n);
    m.c.resize(m.n);
    m.coef.resize(m.n);
    const double two_pi = 2 * kPi;
    for (int j = 0; j < t_num; ++j) {
        double cw = c_w_raw[j] * 2.0;
        for (int i = 0; i < p_num; ++i) {
            int idx = j * p_num + i;  // column-major (p_num, t_num), p fastest
            m.p[idx] = p_nodes[i];
            m.c[idx] = c_nodes[j];
            double pi2 = p_nodes[i] * p_nodes[i];
            m.coef[idx] = (p_w[i] * cw) * pi2 / (two_pi * two_pi);
        }
    }
    return m;
}

}  // namespace

// =======================================================================================
// Flat C API (ctypes)
// =======================================================================================
extern "C" {

struct oracle_config {
    double hbarc, Lambda, m_ud0, m_s0, G, K, T0, a0, a1, a2, b3, rho0;
    int32_t Nc;
    int32_t p_num, t_num;
    const double* p_nodes;   // [p_num]  gauleg(0, 10, p_num)
    const double* p_w;       // [p_num]
    const double* c_nodes;   // [t_num]  gauleg(0, 1, t_num)
    const double* c_w;       // [t_num]  raw weights (doubled internally, Integrals.jl:71-72)
    double xtol, ftol, residual_norm_max, phi_tol;
    int32_t max_iter;
    int32_t tr_fallback, auto_multiseed_fallback;
    double omega_tie_rel;
    int32_t n_threads;  // 0 ⇒ OpenMP default
};

struct oracle_out {  // SoA, caller-allocated, n entries each (x: [5][n], etc.)
    double* x;
    double* mass;
    double* omega;
    double* pressure;
    double* rho_norm;
    double* entropy;
    double* energy;
    double* n_q;
    double* n_qbar;
    double* residual_norm;
    int32_t* iterations;
    int32_t* status;
    int32_t* n_fj;  // Ω/Jacobian-class evaluations spent (FLOP accounting); may be NULL
};

struct oracle_table {  // one per distinct xi
    const double* T_MeV;
    const double* mu_c_MeV;
    int32_t n;
    double T_CEP;
};

static Consts consts_of(const oracle_config* c) {
    Consts k;
    k.hbarc = c->hbarc; k.Lambda = c->Lambda; k.m_ud0 = c->m_ud0; k.m_s0 = c->m_s0; k.G = c->G; k.K = c->K;
    k.T0 = c->T0; k.a0 = c->a0; k.a1 = c->a1; k.a2 = c->a2; k.b3 = c->b3; k.rho0 = c->rho0; k.Nc = c->Nc;
    return k;
}
static SolverOpts opts_of(const oracle_config* c) {
    SolverOpts o;
    o.xtol = c->xtol; o.ftol = c->ftol; o.residual_norm_max = c->residual_norm_max; o.phi_tol = c->phi_tol;
    o.max_iter = c->max_iter; o.tr_fallback = c->tr_fallback != 0;
    o.auto_multiseed_fallback = c->auto_multiseed_fallback != 0; o.omega_tie_rel = c->omega_tie_rel;
    return o;
}
static Mesh mesh_of(const oracle_config* c) {
    return build_mesh(c->p_num, c->t_num, c->p_nodes, c->p_w, c->c_nodes, c->c_w);
}

static void store(const oracle_out* o, int64_t n, int64_t i, const Problem& pb, const PointResult& r) {
    for (int q = 0; q < 5; ++q) o->x[q * n + i] = r.x[q];
    for (int q = 0; q < 3; ++q) o->mass[q * n + i] = r.th.masses[q];
    o->omega[i] = r.th.omega;
    o->pressure[i] = r.th.pressure;
    o->rho_norm[i] = r.th.rho_norm;
    o->entropy[i] = r.th.entropy;
    o->energy[i] = r.th.energy;
    double nq[3], nqb[3];
    number_densities(pb, r.x, nq, nqb);
    for (int q = 0; q < 3; ++q) { o->n_q[q * n + i] = nq[q]; o->n_qbar[q * n + i] = nqb[q]; }
    o->residual_norm[i] = r.residual_norm;
    o->iterations[i] = r.iterations;
    o->status[i] = r.status;
    if (o->n_fj) o->n_fj[i] = r.n_fj;
}

// gausslegendre(n) on [-1,1] — FastGaussQuadrature 1.1.0 stand-in (GaussLegendre.jl:33-42).
void oracle_gausslegendre(int32_t n, double* x, double* w) {
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
        long double pp = 0;
        for (int iter = 0; iter < 100; ++iter) {
            long double p1 = 1, p2 = 0;
            for (int j = 1; j <= n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1);
            long double z1 = z;
            z = z1 - p1 / pp;
            if (fabsl(z - z1) < 1e-19L) break;
        }
        x[i] = (double)(-z);
        x[n - 1 - i] = (double)z;
        long double ww = 2 / ((1 - z * z) * pp * pp);
        w[i] = (double)ww;
        w[n - 1 - i] = (double)ww;
    }
    if (n % 2 == 1) x[n / 2] = 0.0;
}

// gauleg(a, b, n) — GaussLegendre.jl:94-119
void oracle_gauleg(double a, double b, int32_t n, double* x, double* w) {
    oracle_gausslegendre(n, x, w);
    double scale = (b - a) / 2.0, shift = (b + a) / 2.0;
    for (int i = 0; i < n; ++i) {
        x[i] = scale * x[i] + shift;
        w[i] = scale * w[i];
    }
}

double oracle_omega(const oracle_config* c, const double* x, double T_fm, double mu_fm, double xi) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    Problem pb{&k, &m, T_fm, mu_fm, xi};
    return eval_omega(pb, x);
}

void oracle_FJ(const oracle_config* c, const double* x, double T_fm, double mu_fm, double xi, double* F, double* J) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    Problem pb{&k, &m, T_fm, mu_fm, xi};
    eval_FJ(pb, x, F, J);
}

// out[16]: omega, pressure, rho_norm, entropy, energy, rho[3], masses[3], phys(0/1), n_q..., see Python
void oracle_thermo(const oracle_config* c, const double* x, double T_fm, double mu_fm, double xi, double* out) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    Problem pb{&k, &m, T_fm, mu_fm, xi};
    Thermo th = eval_thermo(pb, x);
    out[0] = th.omega; out[1] = th.pressure; out[2] = th.rho_norm; out[3] = th.entropy; out[4] = th.energy;
    for (int i = 0; i < 3; ++i) { out[5 + i] = th.rho[i]; out[8 + i] = th.masses[i]; }
    double nq[3], nqb[3];
    number_densities(pb, x, nq, nqb);
    for (int i = 0; i < 3; ++i) { out[11 + i] = nq[i]; out[14 + i] = nqb[i]; }
}

// One nlsolve call with a trace of iterates.  method: 0 newton, 1 trust_region.
// res[4] = iterations, residual_norm, x_converged, f_converged.  Returns number of trace rows.
int32_t oracle_nlsolve_trace(const oracle_config* c, const double* x0, double T_fm, double mu_fm, double xi,
                             int32_t method, double* zero, double* res, double* trace_x, int32_t trace_cap) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    SolverOpts o = opts_of(c);
    Problem pb{&k, &m, T_fm, mu_fm, xi};
    Trace tr;
    tr.x = trace_x;
    tr.cap = trace_cap;
    NLResult r = method == 0 ? nl_newton(pb, x0, o.xtol, o.ftol, o.max_iter, &tr)
                             : nl_trust_region(pb, x0, o.xtol, o.ftol, o.max_iter, &tr);
    std::memcpy(zero, r.zero, 40);
    res[0] = r.iterations; res[1] = r.residual_norm; res[2] = r.x_converged; res[3] = r.f_converged;
    return tr.n;
}

// Independent points.
//   seed_mode 0: explicit seeds[n][n_seeds][5]; n_seeds == 1 ⇒ solve(DefaultSeed(seed,seed,:hadron))
//                with the auto-MultiSeed fallback per config; n_seeds > 1 ⇒ solve_multi over them.
//   seed_mode 1: DefaultSeed(:auto) (SeedStrategies.jl:193-225) + fallback per config.
//   seed_mode 2: MultiSeed() (six built-in candidates).
// per_seed (optional, seed_mode 2 / explicit multi): [n][n_seeds][8].
int32_t oracle_solve_points(const oracle_config* c, int64_t n, const double* T_fm, const double* mu_fm,
                            const double* xi, int32_t seed_mode, int32_t n_seeds, const double* seeds,
                            const oracle_out* out, double* per_seed) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    SolverOpts o = opts_of(c);
#ifdef _OPENMP
    int nt = c->n_threads > 0 ? c->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t i = 0; i < n; ++i) {
        Problem pb{&k, &m, T_fm[i], mu_fm[i], xi[i]};
        PointResult r;
        if (seed_mode == 0 && n_seeds == 1) {
            r = solve_with_fallback(pb, seeds + i * 5, o);
        } else if (seed_mode == 0) {
            solve_multi(pb, reinterpret_cast<const double(*)[5]>(seeds + i * n_seeds * 5), n_seeds, o, r,
                        per_seed ? per_seed + i * n_seeds * 8 : nullptr);
        } else if (seed_mode == 1) {
            double x0[5];
            default_seed(2, pb.T, pb.mu, x0);
            r = solve_with_fallback(pb, x0, o);
        } else {
            solve_multiseed(pb, o, r, per_seed ? per_seed + i * 6 * 8 : nullptr);
        }
        store(out, n, i, pb, r);
    }
    return 0;
}

// Continuity lines in run_gap_transport_scan.jl order (:407-443): for each line (xi, muq) march
// T ascending; MultiSeed while the tracker has no previous (converged) solution, then
// PhaseAwareContinuitySeed.  Units follow the script: T_fm = T_MeV/ħc, muq_fm = muq_MeV/ħc (:425-427);
// tracker update! gets (T_MeV, muq_MeV) (:441).  Output index = line * n_T + iT.
int32_t oracle_scan_lines_from(const oracle_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                               const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                               const oracle_table* tables, const double* init_x, double init_T_MeV, const oracle_out* out);

int32_t oracle_scan_lines(const oracle_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                          const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                          const oracle_table* tables, const oracle_out* out) {
    return oracle_scan_lines_from(c, n_lines, muq_MeV, xi, table_idx, n_T, T_MeV, n_tables, tables, nullptr, 0.0, out);
}

// The same march started in the middle of a line: init_x [n_lines][5] is the converged solution of the point just before
// T_MeV[0] (at temperature init_T_MeV), handed to the tracker with update! (SeedStrategies.jl:851-856) exactly as the full
// march would have left it.  Lets a timing sample cover a window of the T grid with the seeds — and therefore the evaluation
// counts — of the complete line (bench.py's CPU arm).  init_x == NULL: a fresh line.
int32_t oracle_scan_lines_from(const oracle_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                               const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                               const oracle_table* tables, const double* init_x, double init_T_MeV, const oracle_out* out) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    SolverOpts o = opts_of(c);
    std::vector<PhaseTable> pts(n_tables + 1);  // last = empty table
    for (int t = 0; t < n_tables; ++t) {
        pts[t].T.assign(tables[t].T_MeV, tables[t].T_MeV + tables[t].n);
        pts[t].mu.assign(tables[t].mu_c_MeV, tables[t].mu_c_MeV + tables[t].n);
        pts[t].T_CEP = tables[t].T_CEP;
    }
    const int64_t n = n_lines * (int64_t)n_T;
#ifdef _OPENMP
    int nt = c->n_threads > 0 ? c->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t l = 0; l < n_lines; ++l) {
        Tracker tk;
        int ti = table_idx ? table_idx[l] : -1;
        tk.table = (ti >= 0 && ti < n_tables) ? &pts[ti] : &pts[n_tables];
        if (init_x) tracker_update(tk, init_x + 5 * l, init_T_MeV, muq_MeV[l]);
        for (int it = 0; it < n_T; ++it) {
            double T_fm = T_MeV[it] / k.hbarc;
            double mu_fm = muq_MeV[l] / k.hbarc;
            Problem pb{&k, &m, T_fm, mu_fm, xi[l]};
            PointResult r;
            if (!tk.has_prev) {
                solve_multiseed(pb, o, r);
            } else {
                double x0[5];
                bool sw = tracker_get_seed(tk, T_fm, mu_fm, x0);
                r = solve_with_fallback(pb, x0, o);
                if (sw) r.status |= ST_PHASE_SWITCH;
            }
            if (r.converged) tracker_update(tk, r.x, T_MeV[it], muq_MeV[l]);
            store(out, n, l * n_T + it, pb, r);
        }
    }
    return 0;
}

// TmuScan.run_tmu_scan (src/pnjl/scans/TmuScan.jl:120-234): for each (xi, T) line march mu in the given order.
// Tracker reset per line (:169-172); candidates (:269-300): phase-aware seed, continuation cache, then quark/hadron
// defaults ordered by (T > 150 || mu > 300); every candidate goes through solve() with its automatic fallbacks
// (:349-368); success = converged or residual <= 1e-4 (:411-419); refine (:391-408) and force-promote (:422-458).
// Output index = line * n_mu + imu; rows where every candidate failed carry ST_NO_RESULT and NaNs.
int32_t oracle_tmu_scan(const oracle_config* c, int64_t n_lines, const double* T_MeV, const double* xi,
                        const int32_t* table_idx, int32_t n_mu, const double* mu_MeV, int32_t n_tables,
                        const oracle_table* tables, const oracle_out* out) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    SolverOpts o = opts_of(c);
    std::vector<PhaseTable> pts(n_tables + 1);
    for (int t = 0; t < n_tables; ++t) {
        pts[t].T.assign(tables[t].T_MeV, tables[t].T_MeV + tables[t].n);
        pts[t].mu.assign(tables[t].mu_c_MeV, tables[t].mu_c_MeV + tables[t].n);
        pts[t].T_CEP = tables[t].T_CEP;
    }
    const int64_t n = n_lines * (int64_t)n_mu;
    const double kAcceptable = 1e-4;  // TmuScan.jl:60
#ifdef _OPENMP
    int nt = c->n_threads > 0 ? c->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t l = 0; l < n_lines; ++l) {
        Tracker tk;
        int ti = table_idx ? table_idx[l] : -1;
        tk.table = (ti >= 0 && ti < n_tables) ? &pts[ti] : &pts[n_tables];
        bool have_cache = false;
        double cache[5];
        const double T_fm = T_MeV[l] / k.hbarc;
        for (int im = 0; im < n_mu; ++im) {
            const double mu_fm = mu_MeV[im] / k.hbarc;
            Problem pb{&k, &m, T_fm, mu_fm, xi[l]};
            // candidate slots are numbered 0 phase-aware, 1 continuation cache, 2 and 3 the defaults, whether or not
            // the cache slot is filled (the number is reported in the status word)
            double cand[4][5];
            bool have[4] = {true, have_cache, true, true};
            tracker_get_seed(tk, T_fm, mu_fm, cand[0]);
            if (have_cache) std::memcpy(cand[1], cache, 40);
            if (T_MeV[l] > 150 || mu_MeV[im] > 300) {
                std::memcpy(cand[2], HIGH_TEMP_SEED, 40);
                std::memcpy(cand[3], HADRON_SEED, 40);
            } else {
                std::memcpy(cand[2], HADRON_SEED, 40);
                std::memcpy(cand[3], HIGH_TEMP_SEED, 40);
            }
            PointResult r;
            bool success = false;
            int n_fj = 0;
            for (int ci = 0; ci < 4 && !success; ++ci) {
                if (!have[ci]) continue;
                PointResult a = solve_with_fallback(pb, cand[ci], o);
                n_fj += a.n_fj;
                if (a.status & ST_NONFINITE) continue;   // IsFiniteException -> caught in _solve_point -> next candidate
                bool ok = a.converged || (std::isfinite(a.residual_norm) && a.residual_norm <= kAcceptable);
                if (!ok) continue;
                if (!a.converged) {
                    PointResult b = solve_with_fallback(pb, a.x, o);
                    n_fj += b.n_fj;
                    if (!(b.status & ST_NONFINITE) && b.converged) { a = b; a.status |= ST_REFINED; }
                }
                if (!a.converged) { a.converged = true; a.status |= ST_PROMOTED | ST_CONVERGED; }
                a.status |= (ci << ST_CAND_SHIFT);
                r = a;
                success = true;
            }
            if (!success) {
                r = PointResult{};
                for (int q = 0; q < 5; ++q) r.x[q] = std::numeric_limits<double>::quiet_NaN();
                r.th = Thermo{};
                const double nan = std::numeric_limits<double>::quiet_NaN();
                r.th.omega = r.th.pressure = r.th.rho_norm = r.th.entropy = r.th.energy = nan;
                for (int q = 0; q < 3; ++q) r.th.masses[q] = r.th.rho[q] = nan;
                r.iterations = -1;
                r.residual_norm = nan;
                r.status = ST_NO_RESULT;
            } else {
                tracker_update(tk, r.x, T_MeV[l], mu_MeV[im]);
                std::memcpy(cache, r.x, 40);
                have_cache = true;
            }
            r.n_fj = n_fj;
            if (success) store(out, n, l * n_mu + im, pb, r);
            else {
                const int64_t i = l * n_mu + im;
                for (int q = 0; q < 5; ++q) out->x[q * n + i] = r.x[q];
                for (int q = 0; q < 3; ++q) { out->mass[q * n + i] = r.th.masses[q]; out->n_q[q * n + i] = r.x[0]; out->n_qbar[q * n + i] = r.x[0]; }
                out->omega[i] = out->pressure[i] = out->rho_norm[i] = out->entropy[i] = out->energy[i] = r.x[0];
                out->residual_norm[i] = r.x[0];
                out->iterations[i] = -1;
                out->status[i] = r.status;
                if (out->n_fj) out->n_fj[i] = n_fj;
            }
        }
    }
    return 0;
}

// DualBranchScan.run_dual_branch_scan (src/pnjl/scans/DualBranchScan.jl:104-182).  Output index = (line * 2 + branch) * n_mu
// + imu, branch 0 = hadron (mu ascending, ContinuitySeed(fallback = DefaultSeed(:hadron))), 1 = quark (mu descending,
// fallback DefaultSeed(:quark)); every point is solve() from the fixed seed with its automatic fallbacks (:334-351); a
// branch ends at its first non-converged or jumping point (:420-429) and the rest carries ST_NO_RESULT / NaN.
int32_t oracle_dual_branch(const oracle_config* c, int64_t n_lines, const double* T_MeV, const double* xi, int32_t n_mu,
                           const double* mu_MeV, const oracle_out* out) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
    SolverOpts o = opts_of(c);
    const int64_t n = n_lines * 2 * (int64_t)n_mu;
    const double nan = std::numeric_limits<double>::quiet_NaN();
#ifdef _OPENMP
    int nt = c->n_threads > 0 ? c->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t task = 0; task < 2 * n_lines; ++task) {
        const int64_t l = task / 2;
        const int branch = (int)(task % 2);
        const double T_fm = T_MeV[l] / k.hbarc;
        bool has_prev = false, alive = true;
        double prev[5] = {0, 0, 0, 0, 0}, prev_Mu = 0;
        for (int kk = 0; kk < n_mu; ++kk) {
            const int im = branch == 0 ? kk : n_mu - 1 - kk;
            const double mu_fm = mu_MeV[im] / k.hbarc;
            Problem pb{&k, &m, T_fm, mu_fm, xi[l]};
            const int64_t i = task * n_mu + im;
            bool keep = false;
            PointResult r;
            int n_fj = 0;
            if (alive) {
                double x0[5];
                if (has_prev) std::memcpy(x0, prev, 40);
                else default_seed(branch == 0 ? 0 : 1, T_fm, mu_fm, x0);
                r = solve_with_fallback(pb, x0, o);
                n_fj = r.n_fj;
                if (r.converged && !(r.status & ST_NONFINITE)) {
                    const bool jump = has_prev && (std::fabs(prev[0] - r.x[0]) > 0.5 ||
                                                   std::fabs(prev_Mu - r.th.masses[0]) * 197.327 > 50.0);
                    if (!jump) {
                        keep = true;
                        std::memcpy(prev, r.x, 40);
                        prev_Mu = r.th.masses[0];
                        has_prev = true;
                    }
                }
                if (!keep) alive = false;
            }
            if (keep) {
                store(out, n, i, pb, r);
            } else {
                for (int q = 0; q < 5; ++q) out->x[q * n + i] = nan;
                for (int q = 0; q < 3; ++q) { out->mass[q * n + i] = nan; out->n_q[q * n + i] = nan; out->n_qbar[q * n + i] = nan; }
                out->omega[i] = out->pressure[i] = out->rho_norm[i] = out->entropy[i] = out->energy[i] = nan;
                out->residual_norm[i] = nan;
                out->iterations[i] = -1;
                out->status[i] = ST_NO_RESULT;
                if (out->n_fj) out->n_fj[i] = n_fj;
            }
        }
    }
    return 0;
}

// ---- thermodynamic derivatives ---------------------------------------------------------------------------------------
// src/pnjl/derivatives/ThermoDerivatives.jl differentiates through the solve with ImplicitDifferentiation.jl
// (dx/dtheta = -J^-1 dF/dtheta, :80-109) and ForwardDiff on calculate_thermo / calculate_rho.  Restated with one nested
// dual evaluation of Omega over the seven variables (x[5], T, mu_q): value, gradient and Hessian of P = -Omega.
using D7 = Dual<7, double>;
using D77 = Dual<7, D7>;

// out[57] = P, grad P [7], Hess P [7][7] (row-major), variable order phi_u, phi_d, phi_s, Phi, Phibar, T, mu_q
static void pressure_hessian7(const Problem& pb, const double x[5], double* out) {
    D77 v[7];
    const double base[7] = {x[0], x[1], x[2], x[3], x[4], pb.T, pb.mu};
    for (int i = 0; i < 7; ++i) {
        v[i] = Zero<D77>::make();
        v[i].v.v = base[i];
        v[i].v.d[i] = 1.0;
        v[i].d[i].v = 1.0;
    }
    D77 mu[3] = {v[6], v[6], v[6]};
    D77 om = calculate_omega(v, mu, v[5], *pb.mesh, pb.xi, *pb.k);
    out[0] = -om.v.v;
    for (int i = 0; i < 7; ++i) {
        out[1 + i] = -om.v.d[i];
        for (int j = 0; j < 7; ++j) out[8 + 7 * i + j] = -om.d[j].d[i];
    }
}

static bool solve5_multi(const double J[25], const double* rhs, int nrhs, double* sol) {
    // Gaussian elimination with partial pivoting, nrhs right-hand sides stored as rhs[r * 5 + i]
    double A[5][5 + 4];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 5; ++j) A[i][j] = J[i * 5 + j];
        for (int r = 0; r < nrhs; ++r) A[i][5 + r] = rhs[r * 5 + i];
    }
    for (int c = 0; c < 5; ++c) {
        int piv = c;
        for (int i = c + 1; i < 5; ++i) if (std::fabs(A[i][c]) > std::fabs(A[piv][c])) piv = i;
        if (A[piv][c] == 0.0) return false;
        if (piv != c) for (int j = 0; j < 5 + nrhs; ++j) std::swap(A[piv][j], A[c][j]);
        for (int i = c + 1; i < 5; ++i) {
            const double f = A[i][c] / A[c][c];
            for (int j = c; j < 5 + nrhs; ++j) A[i][j] -= f * A[c][j];
        }
    }
    for (int r = 0; r < nrhs; ++r)
        for (int i = 4; i >= 0; --i) {
            double acc = A[i][5 + r];
            for (int j = i + 1; j < 5; ++j) acc -= A[i][j] * sol[r * 5 + j];
            sol[r * 5 + i] = acc / A[i][i];
        }
    return true;
}

// For n given states x (normally converged solutions at (T, mu, xi)):
//   out[n][32] = [0] v_n_sq, [1] dmuB_dT_sigma, [2..4] masses (compute_masses_from_state, :111-121, with ITS bare masses
//   5.5/197.327 and 140.0/197.327), [5..7] dM_dT, [8..10] dM_dmuB, [11] s, [12] n_B       bulk_viscosity_coefficients :342-467
//   [13] P, [14] eps, [15] dP_dT, [16] dP_dmu, [17] dEps_dT, [18] dEps_dmu, [19] dn_dT, [20] dn_dmu,
//   [21] dP_deps_n, [22] dP_dn_eps, [23..25] dM_dmu (per mu_q)                            thermo_derivatives :186-250, mass_derivatives :132-150
void oracle_thermo_derivatives(const oracle_config* c, int64_t n, const double* T_fm, const double* mu_fm, const double* xi,
                               const double* x, double* out) {
    Consts k = consts_of(c);
    Mesh m = mesh_of(c);
#ifdef _OPENMP
    int nt = c->n_threads > 0 ? c->n_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#endif
    for (int64_t p = 0; p < n; ++p) {
        Problem pb{&k, &m, T_fm[p], mu_fm[p], xi[p]};
        const double* xs = x + 5 * p;
        double h[57];
        pressure_hessian7(pb, xs, h);
        const double P = h[0];
        const double* g = h + 1;
        auto H = [&](int i, int j) { return h[8 + 7 * i + j]; };
        double J[25], rhs[10], dx[10];
        for (int i = 0; i < 5; ++i) {
            for (int j = 0; j < 5; ++j) J[i * 5 + j] = H(i, j);
            rhs[i] = -H(i, 5);        // -dF/dT
            rhs[5 + i] = -H(i, 6);    // -dF/dmu
        }
        double* o = out + 32 * p;
        for (int q = 0; q < 32; ++q) o[q] = std::numeric_limits<double>::quiet_NaN();
        if (!solve5_multi(J, rhs, 2, dx)) continue;
        const double* dxT = dx;
        const double* dxM = dx + 5;
        const double T = pb.T, mu = pb.mu;
        const double s = g[5], nB = g[6] / 3.0;
        auto dot5 = [](const double* a, const double* b) { double r = 0; for (int i = 0; i < 5; ++i) r += a[i] * b[i]; return r; };
        double ds_dx[5], dn_dx[5], F[5];
        for (int i = 0; i < 5; ++i) { ds_dx[i] = H(5, i); dn_dx[i] = H(6, i) / 3.0; F[i] = g[i]; }
        const double ds_dT = H(5, 5) + dot5(ds_dx, dxT);
        const double ds_dmu = H(5, 6) + dot5(ds_dx, dxM);
        const double dn_dT = H(6, 5) / 3.0 + dot5(dn_dx, dxT);
        const double dn_dmu = H(6, 6) / 3.0 + dot5(dn_dx, dxM);
        // dM/dx  (compute_masses_from_state differs from the solver's masses only in the bare masses)
        const double G4 = -4 * k.G, K2 = 2 * k.K;
        const double dM[3][5] = {{G4, K2 * xs[2], K2 * xs[1], 0, 0}, {K2 * xs[2], G4, K2 * xs[0], 0, 0}, {K2 * xs[1], K2 * xs[0], G4, 0, 0}};
        const double mu0 = 0.0055 / 0.197327, ms0 = 0.140 / 0.197327;
        o[2] = mu0 + G4 * xs[0] + K2 * xs[1] * xs[2];
        o[3] = mu0 + G4 * xs[1] + K2 * xs[0] * xs[2];
        o[4] = ms0 + G4 * xs[2] + K2 * xs[0] * xs[1];
        for (int i = 0; i < 3; ++i) {
            o[5 + i] = dot5(dM[i], dxT);
            o[23 + i] = dot5(dM[i], dxM);
            o[8 + i] = o[23 + i] / 3.0;
        }
        const double ds_dmuB = ds_dmu / 3.0, dn_dmuB = dn_dmu / 3.0;
        o[0] = (s * dn_dmuB - nB * dn_dT) / (T * (ds_dT * dn_dmuB - ds_dmuB * dn_dT));
        o[1] = -(nB * ds_dT - s * dn_dT) / (nB * ds_dmuB - s * dn_dmuB);
        o[11] = s;
        o[12] = nB;
        // totals along the solution
        const double P_T = s + dot5(F, dxT), P_mu = g[6] + dot5(F, dxM);
        const double eps = -P + mu * g[6] + T * s;
        const double E_T = -P_T + mu * 3.0 * dn_dT + s + T * ds_dT;
        const double E_mu = -P_mu + g[6] + mu * 3.0 * dn_dmu + T * ds_dmu;
        o[13] = P; o[14] = eps; o[15] = P_T; o[16] = P_mu; o[17] = E_T; o[18] = E_mu; o[19] = dn_dT; o[20] = dn_dmu;
        const double den_e = E_T * dn_dmu - E_mu * dn_dT, den_n = dn_dT * E_mu - dn_dmu * E_T;
        o[21] = den_e == 0 ? std::numeric_limits<double>::quiet_NaN() : (P_T * dn_dmu - P_mu * dn_dT) / den_e;
        o[22] = den_n == 0 ? std::numeric_limits<double>::quiet_NaN() : (P_T * E_mu - P_mu * E_T) / den_n;
    }
}

int32_t oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- one-loop integral A and effective couplings -------------------------------------------------------------------
// Restated literally from the reference (clamped exponential, un-rescaled ratio):
//   quark_distribution / antiquark_distribution          src/QuarkDistribution.jl:14-51
//   const_integral_term_A, A                              src/relaxtime/OneLoopIntegrals.jl:507-543
//   calculate_G_from_A, calculate_effective_couplings     src/relaxtime/EffectiveCouplings.jl:56-60, 232-279
static double ref_clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double ref_quark_distribution(double E, double mu, double T, double Phi, double Phib) {
    const double beta = 1.0 / T;
    const double e1 = ref_clamp(std::exp(-(E - mu) * beta), 1e-200, 1e200);
    const double e2 = e1 * e1, e3 = e2 * e1;
    return (Phi * e1 + 2 * Phib * e2 + e3) / (1 + 3 * Phi * e1 + 3 * Phib * e2 + e3);
}
static double ref_antiquark_distribution(double E, double mu, double T, double Phi, double Phib) {
    const double beta = 1.0 / T;
    const double e1 = ref_clamp(std::exp(-(E + mu) * beta), 1e-200, 1e200);
    const double e2 = e1 * e1, e3 = e2 * e1;
    return (Phib * e1 + 2 * Phi * e2 + e3) / (1 + 3 * Phib * e1 + 3 * Phi * e2 + e3);
}
static double ref_const_integral_term_A(double Lam, double m) {
    const double mp = std::max(m, 0.0);
    if (mp < 1e-14) return (Lam * Lam) / 2.0;
    const double term1 = Lam * std::sqrt(Lam * Lam + mp * mp);
    const double term2 = mp * mp * std::log((Lam + std::sqrt(Lam * Lam + mp * mp)) / mp);
    return (term1 - term2) / 2.0;
}

double oracle_oneloop_A(const oracle_config* c, double m, double mu, double T, double Phi, double Phib, int32_t n,
                        const double* nodes, const double* weights) {
    double integral = -ref_const_integral_term_A(c->Lambda, m);
    for (int i = 0; i < n; ++i) {
        const double p = nodes[i], w = weights[i];
        const double E = std::sqrt(p * p + m * m);
        const double dq = ref_quark_distribution(E, mu, T, Phi, Phib);
        const double da = ref_antiquark_distribution(E, mu, T, Phi, Phib);
        integral += w * (p * p) / E * (dq + da);
    }
    return 4.0 * integral;
}

// out[12] = K0+, K0-, K123+, K123-, K4567+, K4567-, K8+, K8-, K08+, K08-, detK+, detK-
void oracle_effective_couplings(double G, double K, double G_u, double G_s, double* out) {
    const double term_0 = (1.0 / 3.0) * K * (2.0 * G_u + G_s);
    out[0] = G - term_0; out[1] = G + term_0;
    const double term_123 = 0.5 * K * G_s;
    out[2] = G + term_123; out[3] = G - term_123;
    const double term_4567 = 0.5 * K * G_u;
    out[4] = G + term_4567; out[5] = G - term_4567;
    const double term_8 = (1.0 / 6.0) * K * (4.0 * G_u - G_s);
    out[6] = G + term_8; out[7] = G - term_8;
    const double term_08 = (1.0 / 6.0) * std::sqrt(2.0) * K * (G_u - G_s);
    out[8] = term_08; out[9] = -term_08;
    out[10] = out[0] * out[6] - out[8] * out[8];
    out[11] = out[1] * out[7] - out[9] * out[9];
}

// build_K_data (run_gap_transport_scan.jl:297-305) for n states: aux [n][16] = A_u, A_s, G_u, G_s, then the 12 couplings.
void oracle_couplings_batch(const oracle_config* c, int64_t n, const double* T, const double* mu, const double* m_u,
                            const double* m_s, const double* Phi, const double* Phib, int32_t n_rule, const double* nodes,
                            const double* weights, double* aux) {
    const double pi = 3.14159265358979323846;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double* a = aux + 16 * i;
        a[0] = oracle_oneloop_A(c, m_u[i], mu[i], T[i], Phi[i], Phib[i], n_rule, nodes, weights);
        a[1] = oracle_oneloop_A(c, m_s[i], mu[i], T[i], Phi[i], Phib[i], n_rule, nodes, weights);
        a[2] = -c->Nc / (4.0 * pi * pi) * (m_u[i] * a[0]);
        a[3] = -c->Nc / (4.0 * pi * pi) * (m_s[i] * a[1]);
        oracle_effective_couplings(c->G, c->K, a[2], a[3], a + 4);
    }
}

}  // extern "C"
