// pnjl_solver.cuh — the per-point solve cascade, written once over an evaluation policy `Ev`:
//
//     Ev::fj(T, mu, xi, x, F, J)      one Omega-gradient/Jacobian quadrature pass (group-collective on the GPU)
//     Ev::thermo(T, mu, xi, x, th)    one thermo quadrature pass
//
// On the GPU every lane of the group that owns a point runs this code redundantly on identical
// values (the state is 5 doubles; the quadrature passes are where the lanes split the work), so all
// branches below are group-uniform.
//
// What it restates (reference paths relative to the reference repo; NLsolve is un-vendored —
// Manifest.toml pins NLsolve 4.5.1 — and is restated from its published algorithm):
//   newton()            NLsolve newton_ with LineSearches.Static()            call site ImplicitSolver.jl:112
//   trust_region()      NLsolve trust_region_ (factor 1, autoscale) + dogleg! call site ImplicitSolver.jl:144
//   solve_single()      _nlsolve_with_tr_fallback + _choose_candidate + converged flag   ImplicitSolver.jl:72-151, :287
//   solve_multi()       solve_multi + default_omega_selector                  ImplicitSolver.jl:532-559, SeedStrategies.jl:236-240
//   solve_with_fallback() the auto MultiSeed fallback of solve()              ImplicitSolver.jl:306-327
//   default_seed(), multiseed_seeds()                                         SeedStrategies.jl:56-91, :193-284
//   PhaseTables / tracker_*                                                   SeedStrategies.jl:446-475, :762-856
//   scan_line()         the T-march of one (xi, muB) line                     scripts/relaxtime/run_gap_transport_scan.jl:407-443
#pragma once

#include "pnjl_math.cuh"
#include "../../include/pnjl_b200.h"

namespace pnjl {

struct NLRes {
    double x[5];
    double res;
    int it;
    bool xc, fc, threw;
    bool have_th;   // th holds the thermo pass at x (it came for free with a fused final pass)
    Thermo th;
};

struct PointRes {
    double x[5];
    Thermo th;
    double res;
    int it;
    int status;
    int n_fj, n_th;
    bool converged;
};

// First-order boundary tables mu_c(T), one per xi (PhaseBoundaryData, SeedStrategies.jl:365-371), flattened: table t owns the
// rows [start[t], start[t + 1]) of T / mu.  Any number of tables and rows (device: the arrays live in global memory).
struct PhaseTables {
    int n_tables;
    const int* start;      // [n_tables + 1]
    const double* T_CEP;   // [n_tables]  NaN: unknown
    const double* T;       // [rows] ascending within a table
    const double* mu;      // [rows]
};

enum { PH_UNKNOWN = 0, PH_HADRON = 1, PH_QUARK = 2, PH_CROSSOVER = 3 };

PNJL_HD void copy5(double* d, const double* s) {
#pragma unroll
    for (int i = 0; i < 5; ++i) d[i] = s[i];
}

// ---- seeds ---------------------------------------------------------------------------------------
PNJL_HD void seed_const(int which, double out[5]) {
    // 0 HADRON, 1 HIGH_TEMP(=QUARK), 2 VERY_HIGH_TEMP(=HT_0p9), 3 WEAK_CHIRAL_CONF, 4 HT_0p8, 5 HT_0p95
    switch (which) {
        case 0: out[0] = -1.84329; out[1] = -1.84329; out[2] = -2.22701; out[3] = 1.0e-5; out[4] = 4.0e-5; break;
        case 1: out[0] = -0.73192; out[1] = -0.73192; out[2] = -1.79539; out[3] = 0.60532; out[4] = 0.60532; break;
        case 2: out[0] = -0.30; out[1] = -0.30; out[2] = -0.90; out[3] = 0.90; out[4] = 0.90; break;
        case 3: out[0] = -0.50; out[1] = -0.50; out[2] = -1.20; out[3] = 1e-3; out[4] = 1e-3; break;
        case 4: out[0] = -0.50; out[1] = -0.50; out[2] = -1.20; out[3] = 0.80; out[4] = 0.80; break;
        default: out[0] = -0.20; out[1] = -0.20; out[2] = -0.70; out[3] = 0.95; out[4] = 0.95; break;
    }
}

// DefaultSeed get_seed: hint 0 hadron, 1 quark, 2 auto.  The literal 197.327 is the reference's.
PNJL_HD void default_seed(int hint, double T_fm, double mu_fm, double out[5]) {
    const double T_mev = T_fm * 197.327, mu_mev = mu_fm * 197.327;
    if (hint == 2) hint = (T_mev > 150.0 || mu_mev > 300.0) ? 1 : 0;
    if (hint == 1) seed_const(T_mev >= 300.0 ? 2 : 1, out);
    else seed_const(0, out);
}

// The s-th MultiSeed candidate (order of SeedStrategies.jl:257-268).
PNJL_HD void multiseed_seed(int s, double T_fm, double mu_fm, double out[5]) {
    switch (s) {
        case 0: default_seed(0, T_fm, mu_fm, out); break;
        case 1: default_seed(1, T_fm, mu_fm, out); break;
        case 2: seed_const(3, out); break;
        case 3: seed_const(4, out); break;
        case 4: seed_const(2, out); break;
        default: seed_const(5, out); break;
    }
}

// ---- phase table ---------------------------------------------------------------------------------
// _get_current_phase (SeedStrategies.jl:762-782) over interpolate_mu_c (:446-475).  The reference scans the rows for the
// first i with T[i] <= T_MeV <= T[i+1]; for ascending rows and T[0] < T_MeV < T[n-1] that is i = j - 1 with j the first row
// index >= 1 whose T[j] >= T_MeV, found here by bisection (same segment, same arithmetic, any table length).
PNJL_HD int current_phase(const PhaseTables* pt, int ti, double T_MeV, double mu_MeV) {
    if (ti < 0 || ti >= pt->n_tables) return PH_UNKNOWN;  // empty table, NaN CEP -> :unknown
    const double tcep = pt->T_CEP[ti];
    if (tcep == tcep && T_MeV > tcep) return PH_CROSSOVER;
    const int n = pt->start[ti + 1] - pt->start[ti];
    if (n == 0) return PH_UNKNOWN;
    const double* Ts = pt->T + pt->start[ti];
    const double* ms = pt->mu + pt->start[ti];
    double mu_c;
    if (T_MeV <= Ts[0]) mu_c = ms[0];
    else if (T_MeV >= Ts[n - 1]) mu_c = ms[n - 1];
    else if (!(T_MeV == T_MeV)) mu_c = NAN;
    else {
        int lo = 1, hi = n - 1;          // invariant: Ts[hi] >= T_MeV; the answer j is in [lo, hi]
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (Ts[mid] >= T_MeV) hi = mid;
            else lo = mid + 1;
        }
        const int i = lo - 1;
        const double w = (T_MeV - Ts[i]) / (Ts[i + 1] - Ts[i]);
        mu_c = ms[i] + w * (ms[i + 1] - ms[i]);
    }
    if (mu_c != mu_c) return PH_UNKNOWN;
    return mu_MeV < mu_c ? PH_HADRON : PH_QUARK;
}

struct Tracker {
    double prev[5];
    int prev_phase;
    bool has_prev;
};

template <class Ev>
struct Solver {
    const Model& m;
    const SolverParams& sp;
    Ev& ev;
    double T, mu, xi;
    int n_fj, n_th, n_ft;   // full FJ passes, thermo-only passes, fused final passes (F + thermo)
    int its_hint;           // Newton iterations the previous point of this line needed (0: unknown).  Along a continuity line
                            // the residual sequence repeats almost exactly from point to point, so this predicts which pass
                            // will be the last one; it only chooses the KIND of pass (fused final or not), never the iterates.

    PNJL_HD Solver(const Model& m_, const SolverParams& sp_, Ev& ev_)
        : m(m_), sp(sp_), ev(ev_), T(0), mu(0), xi(0), n_fj(0), n_th(0), n_ft(0), its_hint(0) {}

    PNJL_HD void set_point(double T_, double mu_, double xi_) { T = T_; mu = mu_; xi = xi_; }

    PNJL_HD void FJ(const double x[5], double F[5], double J[25]) { ev.fj(T, mu, xi, x, F, J); ++n_fj; }

    // One quadrature pass at x fused with the Newton direction: F(x) and p = -J(x)^{-1} F(x); J never leaves the
    // evaluator's registers.  Returns false when J(x) is exactly singular (p is then unusable).
    PNJL_HD bool FJ_step(const double x[5], double F[5], double p[5]) {
        ++n_fj;
        return ev.fj_step(T, mu, xi, x, F, p);
    }

    // NLsolve newton_: full steps, stop on ||F||inf <= ftol or ||dx||inf <= xtol, NaN guard, iteration cap.
    PNJL_HD_NOINL void newton(const double x0[5], NLRes& r) {
        double x[5], xold[5], f[5], p[5];
        copy5(x, x0);
        bool nonsing = FJ_step(x, f, p);   // F(x0) and the first Newton direction from J(x0)
        r.threw = !all_finite5(f);
        r.have_th = false;
        int it = 0;
        bool xc = false;
        double res = norm_inf5(f);
        bool fc = res <= sp.ftol;
        bool stopped = any_nan5(x) || any_nan5(f);
        if (!r.threw) {
            while (!stopped && !(xc || fc) && it < sp.max_iter) {
                ++it;
                if (!nonsing) {
                    // exactly singular J: NLsolve's regularised step needs J itself (rare; one extra pass, not
                    // counted as a reference-side evaluation)
                    double J[25], f2[5];
                    ev.fj(T, mu, xi, x, f2, J);
                    singular_step(J, f, p);
                }
                if (sp.isospin && x[0] == x[1]) p[1] = p[0];   // keep the exact u<->d symmetry of the equations
                copy5(xold, x);
#pragma unroll
                for (int i = 0; i < 5; ++i) x[i] = x[i] + p[i];
                // NLsolve evaluates only F here and J at the top of the next iteration.  When the previous residual
                // says this pass will most likely end the solve, run it as a fused final pass (F + thermo sums);
                // otherwise fuse F with the next direction.  Either way the iterates are the same.
                // Prediction (predict_tol > 0): certain when the step is at most xtol (the x-criterion stops the solve
                // whatever F is); otherwise "as many iterations as the previous point of the line", and the residual
                // rule ||F|| <= predict_tol when there is no history or the solve outlasts it.
                double pmax = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) pmax = fmax(pmax, fabs(p[i]));
                const bool by_history = its_hint > 0 && it <= its_hint;
                const bool predict = sp.predict_tol > 0.0 &&
                                     (pmax <= sp.xtol || (by_history ? it == its_hint : res <= sp.predict_tol));
                bool fused = predict && ev.f_thermo(T, mu, xi, x, f, r.th);
                if (fused) ++n_ft;
                else nonsing = FJ_step(x, f, p);
                double dx = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const double a = fabs(x[i] - xold[i]);
                    if (a > dx || a != a) dx = a;
                }
                xc = dx <= sp.xtol;
                res = norm_inf5(f);
                fc = res <= sp.ftol;
                stopped = any_nan5(x) || any_nan5(f);
                r.have_th = fused;
                if (fused && !stopped && !(xc || fc) && it < sp.max_iter) {
                    nonsing = FJ_step(x, f, p);   // mispredicted: the solve goes on and needs J(x)
                    r.have_th = false;
                }
            }
        }
        copy5(r.x, x);
        r.it = it;
        r.res = norm_inf5(f);
        r.xc = xc;
        r.fc = fc;
    }

    // NLsolve's singular-Jacobian step: -(J'J + lambda I) p = J' f, lambda = 1e6 sqrt(n eps) ||J'J||_1.
    PNJL_HD_NOINL void singular_step(const double J[25], const double f[5], double p[5]) {
        double JtJ[25], Jtf[5];
        for (int i = 0; i < 5; ++i) {
            for (int j = 0; j < 5; ++j) {
                double s = 0.0;
                for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * J[q * 5 + j];
                JtJ[i * 5 + j] = s;
            }
            double s = 0.0;
            for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * f[q];
            Jtf[i] = s;
        }
        double n1 = 0.0;
        for (int j = 0; j < 5; ++j) {
            double s = 0.0;
            for (int i = 0; i < 5; ++i) s += fabs(JtJ[i * 5 + j]);
            n1 = s > n1 ? s : n1;
        }
        const double lambda = 1e6 * sqrt(5 * 2.220446049250313e-16) * n1;
        for (int i = 0; i < 25; ++i) JtJ[i] = -JtJ[i];
        for (int i = 0; i < 5; ++i) JtJ[i * 5 + i] -= lambda;
        if (!lu_solve5(JtJ, Jtf, p)) {
            for (int i = 0; i < 5; ++i) p[i] = NAN;
        }
    }

    // NLsolve dogleg!
    PNJL_HD_NOINL void dogleg(double p[5], const double r[5], const double d[5], const double J[25], double delta) {
        double p_i[5], p_c[5], g[5];
        const bool ok = lu_solve5(J, r, p_i);
#pragma unroll
        for (int i = 0; i < 5; ++i) p_i[i] = ok ? -p_i[i] : INFINITY;
        if (wnorm5(d, p_i) <= delta) {
            copy5(p, p_i);
            return;
        }
        for (int i = 0; i < 5; ++i) {
            double s = 0.0;
            for (int q = 0; q < 5; ++q) s += J[q * 5 + i] * r[q];
            g[i] = s / (d[i] * d[i]);
        }
        double Jg2 = 0.0;
        for (int q = 0; q < 5; ++q) {
            double s = 0.0;
            for (int i = 0; i < 5; ++i) s += J[q * 5 + i] * g[i];
            Jg2 += s * s;
        }
        const double wg = wnorm5(d, g);
        const double coef = -(wg * wg) / Jg2;
        for (int i = 0; i < 5; ++i) p_c[i] = coef * g[i];
        if (wnorm5(d, p_c) >= delta) {
            const double s = -delta / wg;
            for (int i = 0; i < 5; ++i) p[i] = g[i] * s;
            return;
        }
        double p_diff[5];
        for (int i = 0; i < 5; ++i) p_diff[i] = p_i[i] - p_c[i];
        double wd = 0.0;
        for (int i = 0; i < 5; ++i) wd += (d[i] * p_c[i]) * (d[i] * p_diff[i]);
        const double b = 2 * wd;
        double a = wnorm5(d, p_diff);
        a = a * a;
        const double wc = wnorm5(d, p_c);
        const double tau = (-b + sqrt(b * b - 4 * a * (wc * wc - delta * delta))) / (2 * a);
        for (int i = 0; i < 5; ++i) p[i] = p_c[i] + tau * p_diff[i];
    }

    // NLsolve trust_region_ (factor = 1, autoscale = true).  Every trial counts as an iteration.
    PNJL_HD_NOINL void trust_region(const double x0[5], NLRes& res) {
        double x[5], xold[5], r[5], fv[5], J[25], Jn[25], d[5], p[5];
        copy5(x, x0);
        FJ(x, fv, J);
        copy5(r, fv);
        res.threw = !all_finite5(r);
        int it = 0;
        bool xc = false;
        bool fc = norm_inf5(fv) <= sp.ftol;
        bool stopped = any_nan5(x) || any_nan5(fv);
        bool converged = xc || fc;
        if (!res.threw && !converged) {
            for (int j = 0; j < 5; ++j) {
                double s = 0.0;
                for (int i = 0; i < 5; ++i) s += J[i * 5 + j] * J[i * 5 + j];
                d[j] = sqrt(s);
                if (d[j] == 0.0) d[j] = 1.0;
            }
            double delta = wnorm5(d, x);
            if (delta == 0.0) delta = 1.0;
            const double eta = 1e-4;
            while (!stopped && !converged && it < sp.max_iter) {
                ++it;
                dogleg(p, r, d, J, delta);
                if (sp.isospin && x[0] == x[1]) p[1] = p[0];
                copy5(xold, x);
                for (int i = 0; i < 5; ++i) x[i] += p[i];
                FJ(x, fv, Jn);  // trial residual; Jn is kept only if the step is accepted
                double sr = 0.0, sf = 0.0, spred = 0.0;
                for (int q = 0; q < 5; ++q) {
                    double s = 0.0;
                    for (int i = 0; i < 5; ++i) s += J[q * 5 + i] * p[i];
                    const double rp = s + r[q];
                    sr += r[q] * r[q];
                    sf += fv[q] * fv[q];
                    spred += rp * rp;
                }
                const double rho = (sr - sf) / (sr - spred);
                if (rho > eta) {
                    copy5(r, fv);
                    for (int i = 0; i < 25; ++i) J[i] = Jn[i];
                    for (int j = 0; j < 5; ++j) {
                        double s = 0.0;
                        for (int i = 0; i < 5; ++i) s += J[i * 5 + j] * J[i * 5 + j];
                        const double nj = sqrt(s);
                        d[j] = (0.1 * d[j] > nj) ? 0.1 * d[j] : nj;
                    }
                    double dx = 0.0;
                    for (int i = 0; i < 5; ++i) {
                        const double a = fabs(x[i] - xold[i]);
                        if (a > dx || a != a) dx = a;
                    }
                    xc = dx <= sp.xtol;
                    fc = norm_inf5(r) <= sp.ftol;
                    converged = xc || fc;
                } else {
                    for (int i = 0; i < 5; ++i) x[i] -= p[i];
                    xc = false;
                    converged = false;
                }
                if (rho < 0.1) delta = delta / 2;
                else if (rho >= 0.9) delta = 2 * wnorm5(d, p);
                else if (rho >= 0.5) { const double t = 2 * wnorm5(d, p); delta = delta > t ? delta : t; }
                stopped = any_nan5(x) || any_nan5(fv);
            }
        }
        res.have_th = false;
        copy5(res.x, x);
        res.it = it;
        res.res = norm_inf5(r);
        res.xc = xc;
        res.fc = fc;
    }

    PNJL_HD bool physical(const double x[5], const Thermo& th) const {
        const double P = x[3], Pb = x[4], tol = sp.phi_tol;
        bool ok = finite_d(P) && finite_d(Pb) && (-tol <= P && P <= 1 + tol) && (-tol <= Pb && Pb <= 1 + tol);
#pragma unroll
        for (int i = 0; i < 3; ++i) ok = ok && finite_d(th.M[i]) && th.M[i] > 0.0;
        ok = ok && finite_d(th.omega) && finite_d(th.pressure) && finite_d(th.rho_norm) && finite_d(th.entropy) &&
             finite_d(th.energy);
        return ok;
    }

    PNJL_HD void thermo_at(const double x[5], Thermo& th) { ev.thermo(T, mu, xi, x, th); ++n_th; }

    // solve() for one given seed, without the MultiSeed fallback.
    PNJL_HD_NOINL void solve_single(const double x0[5], PointRes& out) {
        NLRes pr;
        newton(x0, pr);
        out.status = 0;
        if (pr.threw) {
            copy5(out.x, pr.x);
            nan_thermo(out);
            out.res = pr.res;
            out.it = 0;
            out.converged = false;
            out.status = PNJL_ST_NONFINITE;
            return;
        }
        Thermo pth;
        if (pr.have_th) pth = pr.th;
        else thermo_at(pr.x, pth);
        const bool pphys = physical(pr.x, pth);
        const double rmax = sp.residual_norm_max;
        const bool pgood = pr.fc && finite_d(pr.res) && pr.res <= rmax && pphys;
        const bool need_fb = sp.tr_fallback && (!pr.fc || !finite_d(pr.res) || pr.res > rmax || !pphys);
        bool take_f = false;
        NLRes fr;
        Thermo fth;
        bool fphys = false;
        if (need_fb) {
            out.status |= PNJL_ST_TR_ATTEMPTED;
            trust_region(x0, fr);
            if (!fr.threw) {
                thermo_at(fr.x, fth);
                fphys = physical(fr.x, fth);
                const bool fgood = fr.fc && finite_d(fr.res) && fr.res <= rmax && fphys;
                if (fgood && !pgood) take_f = true;
                else if (pgood && !fgood) take_f = false;
                else if (fgood && pgood) {
                    if (fth.omega < pth.omega) take_f = true;
                    else if (fth.omega > pth.omega) take_f = false;
                    else take_f = fr.res < pr.res;
                } else if (fr.fc && !pr.fc) take_f = true;
                else if (pr.fc && !fr.fc) take_f = false;
                else if (finite_d(fr.res) && finite_d(pr.res)) take_f = fr.res < pr.res;
                else take_f = false;
            }
        }
        if (take_f) {
            copy5(out.x, fr.x);
            out.th = fth;
            out.it = fr.it;
            out.res = fr.res;
            out.converged = fr.fc && fphys && finite_d(fr.res) && fr.res <= rmax;
            out.status |= PNJL_ST_USED_TR;
        } else {
            copy5(out.x, pr.x);
            out.th = pth;
            out.it = pr.it;
            out.res = pr.res;
            out.converged = pgood;
        }
        if (out.converged) out.status |= PNJL_ST_CONVERGED;
    }

    PNJL_HD void nan_thermo(PointRes& out) const {
        masses_of(m, out.x, out.th.M);
        out.th.omega = out.th.pressure = out.th.rho_norm = out.th.entropy = out.th.energy = NAN;
#pragma unroll
        for (int i = 0; i < 3; ++i) out.th.rho[i] = out.th.nq[i] = out.th.nqb[i] = NAN;
    }

    // solve_multi over the six built-in candidates (or n_seeds explicit ones), serial in this group.
    // Selection rule shared with the oracle: smallest Omega among converged candidates; candidates within
    // omega_tie_rel * max(1, |Omega_min|) tie and the lowest seed index wins.
    PNJL_HD_NOINL bool solve_multi(const double* explicit_seeds, int n_seeds, PointRes& best) {
        double cx[6][5], com[6], cres[6];
        int cit[6], cst[6];
        bool cconv[6];
        if (n_seeds > 6) n_seeds = 6;
        double omin = INFINITY;
        bool any = false;
        PointRes r;
        for (int s = 0; s < n_seeds; ++s) {
            double x0[5];
            if (explicit_seeds) copy5(x0, explicit_seeds + 5 * s);
            else multiseed_seed(s, T, mu, x0);
            solve_single(x0, r);
            copy5(cx[s], r.x);
            com[s] = r.th.omega; cres[s] = r.res; cit[s] = r.it; cst[s] = r.status; cconv[s] = r.converged;
            if (r.converged) {
                any = true;
                if (r.th.omega < omin) omin = r.th.omega;
            }
            if (s == 0) best = r;
        }
        if (!any) {
            best.converged = false;
            best.status = (best.status & ~PNJL_ST_CONVERGED) | PNJL_ST_ALL_SEEDS_FAILED | PNJL_ST_USED_MULTISEED;
            return false;
        }
        const double tol = sp.omega_tie_rel * (fabs(omin) > 1.0 ? fabs(omin) : 1.0);
        int pick = 0;
        for (int s = n_seeds - 1; s >= 0; --s)
            if (cconv[s] && com[s] <= omin + tol) pick = s;
        copy5(best.x, cx[pick]);
        thermo_at(best.x, best.th);  // recomputed for the winner (same values as during its solve)
        --n_th;                      // bookkeeping: not an extra reference-side evaluation
        best.res = cres[pick];
        best.it = cit[pick];
        best.converged = true;
        best.status = cst[pick] | PNJL_ST_USED_MULTISEED | (pick << PNJL_ST_SEED_SHIFT);
        return true;
    }

    // solve() with a seed from a non-MultiSeed strategy, including the automatic MultiSeed fallback.
    PNJL_HD_NOINL void solve_with_fallback(const double x0[5], PointRes& out) {
        solve_single(x0, out);
        if (out.converged || !sp.auto_multiseed_fallback || (out.status & PNJL_ST_NONFINITE)) return;
        PointRes multi;
        if (solve_multi(nullptr, 6, multi)) {
            multi.status |= (out.status & PNJL_ST_TR_ATTEMPTED);
            out = multi;
        } else {
            out.status |= PNJL_ST_ALL_SEEDS_FAILED;
        }
    }

    // PhaseAwareContinuitySeed get_seed; returns true when it re-seeded at a hadron<->quark flip.
    PNJL_HD bool tracker_seed(const PhaseTables* pt, int ti, const Tracker& tk, double out[5]) const {
        const double T_MeV = T * 197.327, mu_MeV = mu * 197.327;  // SeedStrategies.jl:582,800-801
        const int cur = current_phase(pt, ti, T_MeV, mu_MeV);
        if (!tk.has_prev) {
            if (cur == PH_HADRON) seed_const(0, out);
            else if (cur == PH_QUARK) seed_const(1, out);
            else default_seed(2, T, mu, out);
            return false;
        }
        const bool flip = (tk.prev_phase == PH_HADRON && cur == PH_QUARK) || (tk.prev_phase == PH_QUARK && cur == PH_HADRON);
        if (flip) {
            seed_const(cur == PH_HADRON ? 0 : 1, out);
            return true;
        }
        copy5(out, tk.prev);
        return false;
    }
};

// Record writer shared by host-sim and device code: fills a 32-double record from a PointRes.
PNJL_HD void fill_record(const PointRes& r, double T, double mu, double xi, int n_fj, int n_th, int n_ft,
                         double rec[PNJL_REC_DOUBLES]) {
#pragma unroll
    for (int i = 0; i < 5; ++i) rec[PNJL_REC_X + i] = r.x[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        rec[PNJL_REC_MASS + i] = r.th.M[i];
        rec[PNJL_REC_NQ + i] = r.th.nq[i];
        rec[PNJL_REC_NQBAR + i] = r.th.nqb[i];
        rec[PNJL_REC_RHO + i] = r.th.rho[i];
    }
    rec[PNJL_REC_OMEGA] = r.th.omega;
    rec[PNJL_REC_PRESSURE] = r.th.pressure;
    rec[PNJL_REC_RHO_NORM] = r.th.rho_norm;
    rec[PNJL_REC_ENTROPY] = r.th.entropy;
    rec[PNJL_REC_ENERGY] = r.th.energy;
    rec[PNJL_REC_RESNORM] = r.res;
    rec[PNJL_REC_ITER] = (double)r.it;
    // flag the high-Omega root the reference's physicality filter lets through (see PNJL_ST_MASS_INVERSION); values unchanged
    rec[PNJL_REC_STATUS] = (double)(r.status | ((r.converged && r.th.M[2] <= r.th.M[0]) ? PNJL_ST_MASS_INVERSION : 0));
    rec[PNJL_REC_NEVAL] = (double)n_fj;
    rec[PNJL_REC_NTHERMO] = (double)n_th;
    rec[PNJL_REC_T] = T;
    rec[PNJL_REC_MU] = mu;
    rec[PNJL_REC_XI] = xi;
    rec[PNJL_REC_NFUSED] = (double)n_ft;
    rec[31] = 0.0;
}

// One (xi, muq) line of run_gap_transport_scan.jl:407-443: march T ascending; MultiSeed while the tracker has
// no previous converged solution, afterwards PhaseAwareContinuitySeed (+ solve()'s own fallbacks).
// Units as in the script: T_fm = T_MeV / hbarc, mu_fm = muq_MeV / hbarc (:425-427); the tracker's update!
// receives (T_MeV, muq_MeV) as given (:441), its get_seed converts back from fm with 197.327.
// sink(iT, result, T_fm, mu_fm, n_fj, n_th, n_ft) consumes each point.
template <class Ev, class Sink>
PNJL_HD_NOINL void scan_line(Solver<Ev>& sv, const PhaseTables* pt, int ti, double muq_MeV, double xi, int n_T,
                       const double* T_MeV, Sink& sink) {
    Tracker tk;
    tk.has_prev = false;
    tk.prev_phase = PH_UNKNOWN;
    const double mu_fm = muq_MeV / sv.m.hbarc;
    PointRes r;
    for (int it = 0; it < n_T; ++it) {
        const double Tm = T_MeV[it];
        const double T_fm = Tm / sv.m.hbarc;
        sv.set_point(T_fm, mu_fm, xi);
        sv.n_fj = 0;
        sv.n_th = 0;
        sv.n_ft = 0;
        if (!tk.has_prev) {
            sv.its_hint = 0;
            sv.solve_multi(nullptr, 6, r);
        } else {
            double x0[5];
            const bool sw = sv.tracker_seed(pt, ti, tk, x0);
            if (sw) sv.its_hint = 0;               // re-seeded at a phase flip: no history for this point
            sv.solve_with_fallback(x0, r);
            if (sw) r.status |= PNJL_ST_PHASE_SWITCH;
        }
        // history for the next point: plain Newton successes only (fallback results took another route)
        sv.its_hint = (r.converged && !(r.status & (PNJL_ST_USED_TR | PNJL_ST_TR_ATTEMPTED | PNJL_ST_USED_MULTISEED))) ? r.it : 0;
        if (r.converged) {
            copy5(tk.prev, r.x);
            tk.has_prev = true;
            tk.prev_phase = current_phase(pt, ti, Tm, muq_MeV);
        }
        sink(it, r, T_fm, mu_fm, sv.n_fj, sv.n_th, sv.n_ft);
    }
}

// The same march in resumable form: everything a (xi, muq) line carries from one T to the next — the tracker of
// PhaseAwareContinuitySeed (SeedStrategies.jl:851-856: previous solution + its phase), the index of the next T and the
// iteration-count history — lives in a LineState, so that a line can be advanced a few points at a time by whichever
// warp (team) is free and parked in between (k_march: time-sliced lines from a global queue).  Marching a line in
// slices gives bit-identical records to scan_line() in one go.
struct LineState {
    double prev[5];     // tracker: previous converged solution
    int it_next;        // next T index to solve
    int prev_phase;     // tracker: phase of the previous converged point
    int has_prev;       // tracker holds a solution
    int its_hint;       // Solver::its_hint carried across points
    int pad[4];
};

template <class Ev, class Sink>
PNJL_HD_NOINL void scan_line_slice(Solver<Ev>& sv, const PhaseTables* pt, int ti, double muq_MeV, double xi, int n_T,
                                   const double* T_MeV, LineState& st, int max_points, Sink& sink) {
    Tracker tk;
    copy5(tk.prev, st.prev);
    tk.has_prev = st.has_prev != 0;
    tk.prev_phase = st.prev_phase;
    sv.its_hint = st.its_hint;
    const double mu_fm = muq_MeV / sv.m.hbarc;
    const int it_end = (st.it_next + max_points < n_T) ? st.it_next + max_points : n_T;
    PointRes r;
    for (int it = st.it_next; it < it_end; ++it) {
        const double Tm = T_MeV[it];
        const double T_fm = Tm / sv.m.hbarc;
        sv.set_point(T_fm, mu_fm, xi);
        sv.n_fj = 0;
        sv.n_th = 0;
        sv.n_ft = 0;
        if (!tk.has_prev) {
            sv.its_hint = 0;
            sv.solve_multi(nullptr, 6, r);
        } else {
            double x0[5];
            const bool sw = sv.tracker_seed(pt, ti, tk, x0);
            if (sw) sv.its_hint = 0;
            sv.solve_with_fallback(x0, r);
            if (sw) r.status |= PNJL_ST_PHASE_SWITCH;
        }
        sv.its_hint = (r.converged && !(r.status & (PNJL_ST_USED_TR | PNJL_ST_TR_ATTEMPTED | PNJL_ST_USED_MULTISEED))) ? r.it : 0;
        if (r.converged) {
            copy5(tk.prev, r.x);
            tk.has_prev = true;
            tk.prev_phase = current_phase(pt, ti, Tm, muq_MeV);
        }
        sink(it, r, T_fm, mu_fm, sv.n_fj, sv.n_th, sv.n_ft);
    }
    copy5(st.prev, tk.prev);
    st.has_prev = tk.has_prev ? 1 : 0;
    st.prev_phase = tk.prev_phase;
    st.its_hint = sv.its_hint;
    st.it_next = it_end;
}

// One (xi, T) line of TmuScan.run_tmu_scan (src/pnjl/scans/TmuScan.jl:120-234): march mu in the given order.
// The tracker starts empty on every line (:169-172).  Seed candidates per point (:269-300): the phase-aware seed,
// the continuation cache (last success on this line), then the quark/hadron defaults ordered by
// (T > 150 MeV || mu > 300 MeV).  Every candidate runs through solve() with its automatic fallbacks (:349-368); a
// candidate succeeds when converged or when its residual is <= 1e-4 (:411-419); a near-converged success is re-solved
// from its own state once (:391-408) and otherwise force-marked converged (:422-458).  Rows where every candidate
// failed get PNJL_ST_NO_RESULT and NaNs (the reference writes an all-NaN CSV row).
template <class Ev, class Sink>
PNJL_HD_NOINL void scan_tmu_line(Solver<Ev>& sv, const PhaseTables* pt, int ti, double T_MeV, double xi, int n_mu,
                                 const double* mu_MeV, Sink& sink) {
    Tracker tk;
    tk.has_prev = false;
    tk.prev_phase = PH_UNKNOWN;
    bool have_cache = false;
    double cache[5];
    const double T_fm = T_MeV / sv.m.hbarc;
    const double kAcceptable = 1e-4;
    PointRes r, a;
    int line_hint = 0;
    for (int im = 0; im < n_mu; ++im) {
        const double mu_fm = mu_MeV[im] / sv.m.hbarc;
        sv.set_point(T_fm, mu_fm, xi);
        sv.n_fj = 0;
        sv.n_th = 0;
        sv.n_ft = 0;
        const bool quark_first = (T_MeV > 150.0) || (mu_MeV[im] > 300.0);
        bool success = false;
        for (int ci = 0; ci < 4 && !success; ++ci) {
            double x0[5];
            bool sw = false;
            if (ci == 0) sw = sv.tracker_seed(pt, ti, tk, x0);
            else if (ci == 1) { if (!have_cache) continue; copy5(x0, cache); }
            else seed_const(((ci == 2) == quark_first) ? 1 : 0, x0);
            // pass-kind prediction from the line's history only for the continuity candidate (see Solver::its_hint)
            sv.its_hint = (ci == 0 && tk.has_prev && !sw) ? line_hint : 0;
            sv.solve_with_fallback(x0, a);
            if (a.status & PNJL_ST_NONFINITE) continue;
            const bool ok = a.converged || (finite_d(a.res) && a.res <= kAcceptable);
            if (!ok) continue;
            int extra = 0;
            if (!a.converged) {
                double xs[5];
                copy5(xs, a.x);
                sv.solve_with_fallback(xs, r);
                if (!(r.status & PNJL_ST_NONFINITE) && r.converged) extra = PNJL_ST_REFINED;
                else r = a;
            } else {
                r = a;
            }
            r.status |= extra;
            if (!r.converged) {
                r.converged = true;
                r.status |= PNJL_ST_PROMOTED | PNJL_ST_CONVERGED;
            }
            r.status |= (ci << PNJL_ST_CAND_SHIFT);
            success = true;
        }
        if (!success) {
#pragma unroll
            for (int q = 0; q < 5; ++q) r.x[q] = NAN;
            sv.nan_thermo(r);
#pragma unroll
            for (int q = 0; q < 3; ++q) r.th.M[q] = NAN;
            r.res = NAN;
            r.it = -1;
            r.converged = false;
            r.status = PNJL_ST_NO_RESULT;
        } else {
            copy5(tk.prev, r.x);
            tk.has_prev = true;
            tk.prev_phase = current_phase(pt, ti, T_MeV, mu_MeV[im]);
            copy5(cache, r.x);
            have_cache = true;
        }
        line_hint = (success && !(r.status & (PNJL_ST_USED_TR | PNJL_ST_TR_ATTEMPTED | PNJL_ST_USED_MULTISEED | PNJL_ST_REFINED |
                                              PNJL_ST_PROMOTED | PNJL_ST_CAND_MASK))) ? r.it : 0;
        sink(im, r, T_fm, mu_fm, sv.n_fj, sv.n_th, sv.n_ft);
    }
}

// One branch of DualBranchScan.run_dual_branch_scan (src/pnjl/scans/DualBranchScan.jl:104-182) for one (T, xi) line:
//   branch 0 "hadron": mu ascending,  ContinuitySeed(fallback = DefaultSeed(phase_hint = :hadron))   :121-146
//   branch 1 "quark":  mu descending, ContinuitySeed(fallback = DefaultSeed(phase_hint = :quark))    :148-174
// Every point is solve(FixedMu(), T, mu; seed_strategy = fixed seed) with its automatic fallbacks (_solve_point :334-351;
// any exception = no result).  The branch stops at the first point that does not converge or that jumps
// (|d phi_u| > 0.5 or |d M_u| > 50 MeV against the last accepted point, _is_solution_jump :420-429); that point and
// everything after it carry PNJL_ST_NO_RESULT and NaNs ("nothing" in the reference's branch vectors).
template <class Ev, class Sink>
PNJL_HD_NOINL void scan_branch_line(Solver<Ev>& sv, double T_MeV, double xi, int n_mu, const double* mu_MeV, int branch,
                                    Sink& sink) {
    const double T_fm = T_MeV / sv.m.hbarc;
    bool has_prev = false, alive = true;
    double prev[5] = {0, 0, 0, 0, 0};
    double prev_Mu = 0.0;
    PointRes r;
    int line_hint = 0;
    for (int k = 0; k < n_mu; ++k) {
        const int im = branch == 0 ? k : n_mu - 1 - k;
        const double mu_fm = mu_MeV[im] / sv.m.hbarc;
        sv.set_point(T_fm, mu_fm, xi);
        sv.n_fj = 0;
        sv.n_th = 0;
        sv.n_ft = 0;
        bool keep = false;
        if (alive) {
            double x0[5];
            if (has_prev) copy5(x0, prev);
            else default_seed(branch == 0 ? 0 : 1, T_fm, mu_fm, x0);
            sv.its_hint = has_prev ? line_hint : 0;
            sv.solve_with_fallback(x0, r);
            line_hint = (r.converged && !(r.status & (PNJL_ST_USED_TR | PNJL_ST_TR_ATTEMPTED | PNJL_ST_USED_MULTISEED))) ? r.it : 0;
            if (r.converged && !(r.status & PNJL_ST_NONFINITE)) {
                const bool jump = has_prev && (fabs(prev[0] - r.x[0]) > 0.5 || fabs(prev_Mu - r.th.M[0]) * 197.327 > 50.0);
                if (!jump) {
                    keep = true;
                    copy5(prev, r.x);
                    prev_Mu = r.th.M[0];
                    has_prev = true;
                }
            }
            if (!keep) alive = false;
        }
        if (!keep) {
#pragma unroll
            for (int q = 0; q < 5; ++q) r.x[q] = NAN;
            sv.nan_thermo(r);
#pragma unroll
            for (int q = 0; q < 3; ++q) r.th.M[q] = NAN;
            r.res = NAN;
            r.it = -1;
            r.converged = false;
            r.status = PNJL_ST_NO_RESULT;
        }
        sink(im, r, T_fm, mu_fm, sv.n_fj, sv.n_th, sv.n_ft);
    }
}

}  // namespace pnjl
