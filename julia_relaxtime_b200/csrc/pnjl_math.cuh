// pnjl_math.cuh — per-node integrand with analytic first/second derivatives, closed-form parts of
// the thermodynamic potential, and the small dense algebra of the Newton / dogleg steps.
//
// Everything here is `PNJL_HD` (host + device) so the very same formulas can be exercised on the
// CPU by the test-only harness tests/hostsim/ (never part of the product library) and compared with
// the oracle's dual-number derivatives.
//
// Reference formulas (paths relative to the reference repo):
//   Omega = chi + U + vacuum + thermal                    src/pnjl/core/Thermodynamics.jl:181-195
//   masses                                                 Thermodynamics.jl:81-88
//   chi                                                    Thermodynamics.jl:112-114
//   U(T, Phi, Phibar) with floored log                     Thermodynamics.jl:124-130, Integrals.jl:159-162
//   vacuum integral I(Lambda, M), M -> |M| + 1e-12         src/pnjl/core/Integrals.jl:123-145
//   log term ln f+ + ln f- with the a>0 rescaling + floors Integrals.jl:193-238
//   E = sqrt(p^2 + M^2 + xi (p cos)^2)                     Integrals.jl:178-180
//   f+- distributions (explicit densities)                 src/QuarkDistribution.jl:14-51
// The reference differentiates these with ForwardDiff; here the derivatives are written out.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PNJL_HD __host__ __device__ __forceinline__
#define PNJL_HD_NOINL __host__ __device__ __noinline__
#else
#define PNJL_HD inline
#define PNJL_HD_NOINL
#endif

namespace pnjl {

constexpr double kPi = 3.14159265358979323846;
constexpr double kPolyakovEps = 1e-16;  // Integrals.jl:152

struct Model {
    double hbarc, Lambda, m_ud0, m_s0, G, K, T0, a0, a1, a2, b3, rho0;
    double Nc;
};

struct SolverParams {
    double xtol, ftol, residual_norm_max, phi_tol, omega_tie_rel;
    int max_iter, tr_fallback, auto_multiseed_fallback, pad;
};

// ------------------------------------------------------------------------------------------------
// fast FP64 primitives (device: MUFU seed + Newton-Raphson; host: libm so that hostsim is portable)
// ------------------------------------------------------------------------------------------------
PNJL_HD double f_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
PNJL_HD double f_rcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}
PNJL_HD double f_exp(double x) { return exp(x); }
PNJL_HD double f_log(double x) { return log(x); }
PNJL_HD double f_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return fma(a, b, c);
#else
    return a * b + c;
#endif
}

// ------------------------------------------------------------------------------------------------
// One Polyakov-loop species term L = ln(1 + 3 P1 y + 3 P2 y^2 + y^3), y = e^a, evaluated in the
// scale-free form of Integrals.jl:203-235: with w = e^{-|a|}
//   a <= 0:  D = 1 + 3 P1 w + 3 P2 w^2 + w^3          (y, y^2, y^3)/f = (w, w^2, w^3)/D
//   a >  0:  D = w^3 + 3 P1 w^2 + 3 P2 w + 1          (y, y^2, y^3)/f = (w^2, w, 1)/D,  L = 3a + ln D
// and the reference's floor max(D, 1e-16): when active, ln D is constant, i.e. every derivative of
// ln D vanishes (ForwardDiff's rule for max) and only the explicit 3a term remains.
// Outputs: r1 = y/f, r2 = y^2/f, r3 = y^3/f  (zeroed under the floor, r3 -> 1 on the a>0 side).
// ------------------------------------------------------------------------------------------------
struct Species {
    double r1, r2, r3;
    double D;    // floored denominator (for ln D in the value pass)
    bool pos;    // a > 0
};

PNJL_HD Species species_eval(double a, double P1x3, double P2x3) {
    Species s;
    const double w = f_exp(-fabs(a));
    const double w2 = w * w;
    const double w3 = w2 * w;
    s.pos = a > 0.0;
    const double t0 = s.pos ? w3 : 1.0;
    const double t1 = s.pos ? w2 : w;
    const double t2 = s.pos ? w : w2;
    const double t3 = s.pos ? 1.0 : w3;
    double D = f_fma(P1x3, t1, t0);
    D = f_fma(P2x3, t2, D);
    D += t3;
    const bool floored = D < kPolyakovEps;
    const double inv = f_rcp(D);
    s.r1 = floored ? 0.0 : t1 * inv;
    s.r2 = floored ? 0.0 : t2 * inv;
    s.r3 = floored ? (s.pos ? 1.0 : 0.0) : t3 * inv;
    s.D = floored ? kPolyakovEps : D;
    return s;
}

// ------------------------------------------------------------------------------------------------
// Accumulators of one Omega-gradient/Jacobian quadrature pass ("FJ pass").
// Per flavour i (5 each):  S1 = sum c (n+ + n-) / E
//                          S2a = sum c Q / E^2          Q = (q+ - 3 n+^2) + (q- - 3 n-^2)
//                          S2b = sum c (n+ + n-) k^2 / E^3
//                          S3 = sum c [r1+ (1 - 3 n+) + r2- (2 - 3 n-)] / E      (d/dPhi    of dL/dE)
//                          S4 = sum c [r2+ (2 - 3 n+) + r1- (1 - 3 n-)] / E      (d/dPhibar of dL/dE)
// Flavour-summed (5):      GP = sum c (r1+ + r2-),  GPb = sum c (r2+ + r1-)
//                          HPP = sum c (r1+^2 + r2-^2), HPPb = sum c (r1+ r2+ + r2- r1-), HPbPb = sum c (r2+^2 + r1-^2)
// with n = P1 r1 + 2 P2 r2 + r3 (occupation), q = P1 r1 + 4 P2 r2 + 3 r3.
// ------------------------------------------------------------------------------------------------
constexpr int kFJAcc = 20;
enum { ACC_S1 = 0, ACC_S2A = 3, ACC_S2B = 6, ACC_S3 = 9, ACC_S4 = 12, ACC_GP = 15, ACC_GPB = 16, ACC_HPP = 17,
       ACC_HPPB = 18, ACC_HPBPB = 19 };

struct PointCtx {
    double T, mu, xi, invT;
    double Phi, Phib, Phi3, Phib3;  // Phi, Phibar and 3x
    double M[3], M2[3];
};

PNJL_HD void masses_of(const Model& m, const double x[5], double M[3]) {
    M[0] = m.m_ud0 - 4 * m.G * x[0] + 2 * m.K * x[1] * x[2];
    M[1] = m.m_ud0 - 4 * m.G * x[1] + 2 * m.K * x[0] * x[2];
    M[2] = m.m_s0 - 4 * m.G * x[2] + 2 * m.K * x[0] * x[1];
}

PNJL_HD void make_ctx(const Model& m, double T, double mu, double xi, const double x[5], PointCtx& c) {
    c.T = T; c.mu = mu; c.xi = xi; c.invT = 1.0 / T;
    c.Phi = x[3]; c.Phib = x[4]; c.Phi3 = 3.0 * x[3]; c.Phib3 = 3.0 * x[4];
    masses_of(m, x, c.M);
    for (int i = 0; i < 3; ++i) c.M2[i] = c.M[i] * c.M[i];
}

// One node x one flavour of the FJ pass.  k2 = p^2 + xi (p cos)^2, coef = quadrature coefficient.
template <int FL>
PNJL_HD void fj_node(const PointCtx& c, double k2, double coef, double acc[kFJAcc]) {
    const double E2 = k2 + c.M2[FL];
    const double rE = f_rsqrt(E2);
    const double E = E2 * rE;
    const double a = (c.mu - E) * c.invT;   // -(E - mu)/T
    const double b = -(E + c.mu) * c.invT;  // -(E + mu)/T
    const Species sp = species_eval(a, c.Phi3, c.Phib3);   // quark:      P1 = Phi,    P2 = Phibar
    const Species sm = species_eval(b, c.Phib3, c.Phi3);   // antiquark:  P1 = Phibar, P2 = Phi
    const double np = f_fma(c.Phi, sp.r1, f_fma(2.0 * c.Phib, sp.r2, sp.r3));
    const double nm = f_fma(c.Phib, sm.r1, f_fma(2.0 * c.Phi, sm.r2, sm.r3));
    const double qp = f_fma(c.Phi, sp.r1, f_fma(4.0 * c.Phib, sp.r2, 3.0 * sp.r3));
    const double qm = f_fma(c.Phib, sm.r1, f_fma(4.0 * c.Phi, sm.r2, 3.0 * sm.r3));
    const double nsum = np + nm;
    const double Q = f_fma(-3.0 * np, np, qp) + f_fma(-3.0 * nm, nm, qm);
    const double crE = coef * rE;
    const double crE2 = crE * rE;
    acc[ACC_S1 + FL] = f_fma(crE, nsum, acc[ACC_S1 + FL]);
    acc[ACC_S2A + FL] = f_fma(crE2, Q, acc[ACC_S2A + FL]);
    acc[ACC_S2B + FL] = f_fma(crE2 * rE, nsum * k2, acc[ACC_S2B + FL]);
    const double up = f_fma(-3.0, np, 1.0), vp = f_fma(-3.0, np, 2.0);
    const double um = f_fma(-3.0, nm, 1.0), vm = f_fma(-3.0, nm, 2.0);
    acc[ACC_S3 + FL] = f_fma(crE, f_fma(sp.r1, up, sm.r2 * vm), acc[ACC_S3 + FL]);
    acc[ACC_S4 + FL] = f_fma(crE, f_fma(sp.r2, vp, sm.r1 * um), acc[ACC_S4 + FL]);
    acc[ACC_GP] = f_fma(coef, sp.r1 + sm.r2, acc[ACC_GP]);
    acc[ACC_GPB] = f_fma(coef, sp.r2 + sm.r1, acc[ACC_GPB]);
    acc[ACC_HPP] = f_fma(coef, f_fma(sp.r1, sp.r1, sm.r2 * sm.r2), acc[ACC_HPP]);
    acc[ACC_HPPB] = f_fma(coef, f_fma(sp.r1, sp.r2, sm.r2 * sm.r1), acc[ACC_HPPB]);
    acc[ACC_HPBPB] = f_fma(coef, f_fma(sp.r2, sp.r2, sm.r1 * sm.r1), acc[ACC_HPBPB]);
}

// ------------------------------------------------------------------------------------------------
// Closed-form pieces
// ------------------------------------------------------------------------------------------------
// Vacuum integral I(Lambda, m) and its first two derivatives in m = |M| + 1e-12.
//   16 pi^2 I   = L s (2 L^2 + m^2) - m^4 ln((L + s)/m),   s = sqrt(L^2 + m^2)
//    4 pi^2 I'  = L m s - m^3 ln((L + s)/m)
//    4 pi^2 I'' = L s + 2 L m^2 / s - 3 m^2 ln((L + s)/m)
PNJL_HD void vacuum_terms(double Lam, double M, double& I0, double& I1, double& I2) {
    const double m = fabs(M) + 1e-12;
    const double m2 = m * m;
    const double s = sqrt(Lam * Lam + m2);
    const double lg = log((Lam + s) / m);
    I0 = (Lam * s * (2 * (Lam * Lam) + m2) - (m2 * m2) * lg) / (16 * (kPi * kPi));
    const double sgn = (M > 0.0) ? 1.0 : ((M < 0.0) ? -1.0 : 0.0);
    I1 = sgn * (Lam * m * s - m2 * m * lg) / (4 * (kPi * kPi));
    I2 = (Lam * s + 2 * Lam * m2 / s - 3 * m2 * lg) / (4 * (kPi * kPi));
}

struct UTerms {
    double U, U_P, U_Pb, U_PP, U_PPb, U_PbPb, U_T;
};

// U = T^4 [ -1/2 A(T) Phi Phibar + B(T) ln v ],  v = 1 - 6 Phi Phibar + 4 (Phi^3 + Phibar^3) - 3 (Phi Phibar)^2,
// A = a0 + a1 t + a2 t^2, B = b3 t^3, t = T0/T;  ln v is floored at ln 1e-16 (safe_log) -> zero derivative.
PNJL_HD UTerms polyakov_U(const Model& m, double T, double P, double Pb) {
    UTerms u;
    const double t = m.T0 / T;
    const double A = m.a0 + m.a1 * t + m.a2 * (t * t);
    const double B = m.b3 * (t * t * t);
    const double T2 = T * T, T4 = T2 * T2;
    const double PPb = Pb * P;
    const double v = 1 - 6 * PPb + 4 * (Pb * Pb * Pb + P * P * P) - 3 * (PPb * PPb);
    const bool live = !(v <= 0.0) && !(v < kPolyakovEps);
    const double lv = live ? log(v) : log(kPolyakovEps);
    const double iv = live ? 1.0 / v : 0.0;
    const double vP = -6 * Pb + 12 * P * P - 6 * P * Pb * Pb;
    const double vPb = -6 * P + 12 * Pb * Pb - 6 * P * P * Pb;
    const double vPP = 24 * P - 6 * Pb * Pb;
    const double vPbPb = 24 * Pb - 6 * P * P;
    const double vPPb = -6 - 12 * PPb;
    u.U = T4 * (-0.5 * A * PPb + B * lv);
    u.U_P = T4 * (-0.5 * A * Pb + B * vP * iv);
    u.U_Pb = T4 * (-0.5 * A * P + B * vPb * iv);
    u.U_PP = T4 * B * (vPP * iv - (vP * iv) * (vP * iv));
    u.U_PbPb = T4 * B * (vPbPb * iv - (vPb * iv) * (vPb * iv));
    u.U_PPb = T4 * (-0.5 * A + B * (vPPb * iv - (vP * iv) * (vPb * iv)));
    // dU/dT at fixed Phi: Thermodynamics.jl:147-165
    const double dA = -m.a1 * m.T0 / T2 - 2 * m.a2 * (m.T0 * m.T0) / (T2 * T);
    const double dB = -3 * m.b3 * (m.T0 * m.T0 * m.T0) / T4;
    u.U_T = 4 * (T2 * T) * (-0.5 * A * PPb + B * lv) + T4 * (dA * (-0.5 * PPb) + dB * lv);
    return u;
}

// Assemble F = grad_x P (5) and J = Hess_x P (5x5 row-major) from the reduced accumulators.
PNJL_HD void finish_fj(const Model& m, const PointCtx& c, const double x[5], const double acc[kFJAcc], double F[5],
                       double J[25]) {
    const double T = c.T, invT = c.invT;
    const double twoT = 2.0 * T;
    // dP/dM_i, d2P/dM_i^2, d2P/dM_i dPhi, d2P/dM_i dPhibar  (thermal + vacuum)
    double PM[3], PMM[3], PMP[3], PMPb[3];
    for (int i = 0; i < 3; ++i) {
        double I0, I1, I2;
        vacuum_terms(m.Lambda, c.M[i], I0, I1, I2);
        const double S1 = -3.0 * invT * c.M[i] * acc[ACC_S1 + i];
        const double S2 = 3.0 * invT * invT * c.M2[i] * acc[ACC_S2A + i] - 3.0 * invT * acc[ACC_S2B + i];
        const double S3 = -3.0 * invT * c.M[i] * acc[ACC_S3 + i];
        const double S4 = -3.0 * invT * c.M[i] * acc[ACC_S4 + i];
        PM[i] = twoT * S1 + 2.0 * m.Nc * I1;
        PMM[i] = twoT * S2 + 2.0 * m.Nc * I2;
        PMP[i] = twoT * S3;
        PMPb[i] = twoT * S4;
    }
    // dM_i/dphi_j
    const double g4 = -4.0 * m.G, k2 = 2.0 * m.K;
    double D[3][3];
    D[0][0] = g4;        D[0][1] = k2 * x[2]; D[0][2] = k2 * x[1];
    D[1][0] = k2 * x[2]; D[1][1] = g4;        D[1][2] = k2 * x[0];
    D[2][0] = k2 * x[1]; D[2][1] = k2 * x[0]; D[2][2] = g4;
    const UTerms u = polyakov_U(m, T, x[3], x[4]);
    // -chi: d/dphi_j = -4G phi_j + 4K phi_k phi_l
    const double chi1[3] = {-4 * m.G * x[0] + 4 * m.K * x[1] * x[2], -4 * m.G * x[1] + 4 * m.K * x[0] * x[2],
                            -4 * m.G * x[2] + 4 * m.K * x[0] * x[1]};
    for (int j = 0; j < 3; ++j) {
        F[j] = PM[0] * D[0][j] + PM[1] * D[1][j] + PM[2] * D[2][j] + chi1[j];
        for (int l = 0; l < 3; ++l) {
            double v = PMM[0] * D[0][j] * D[0][l] + PMM[1] * D[1][j] * D[1][l] + PMM[2] * D[2][j] * D[2][l];
            if (j == l) {
                v += -4.0 * m.G;
            } else {
                const int o = 3 - j - l;  // the third flavour
                v += k2 * PM[o] + 4.0 * m.K * x[o];
            }
            J[j * 5 + l] = v;
        }
        const double jp = PMP[0] * D[0][j] + PMP[1] * D[1][j] + PMP[2] * D[2][j];
        const double jpb = PMPb[0] * D[0][j] + PMPb[1] * D[1][j] + PMPb[2] * D[2][j];
        J[j * 5 + 3] = jp;  J[3 * 5 + j] = jp;
        J[j * 5 + 4] = jpb; J[4 * 5 + j] = jpb;
    }
    F[3] = twoT * 3.0 * acc[ACC_GP] - u.U_P;
    F[4] = twoT * 3.0 * acc[ACC_GPB] - u.U_Pb;
    J[3 * 5 + 3] = -9.0 * twoT * acc[ACC_HPP] - u.U_PP;
    J[3 * 5 + 4] = -9.0 * twoT * acc[ACC_HPPB] - u.U_PPb;
    J[4 * 5 + 3] = J[3 * 5 + 4];
    J[4 * 5 + 4] = -9.0 * twoT * acc[ACC_HPBPB] - u.U_PbPb;
}

// ------------------------------------------------------------------------------------------------
// Thermo pass: value of the thermal sum, occupation sums and the T-derivative sum.
//   per flavour i: TN+ = sum c n+, TN- = sum c n-        (-> rho_i, n_i, n_ibar)
//   total:         TL  = sum c (L+ + L-),  TT = sum c [n+ (E - mu) + n- (E + mu)]
// ------------------------------------------------------------------------------------------------
constexpr int kThAcc = 8;
enum { TH_NP = 0, TH_NM = 3, TH_L = 6, TH_T = 7 };

template <int FL>
PNJL_HD void thermo_node(const PointCtx& c, double k2, double coef, double acc[kThAcc]) {
    const double E2 = k2 + c.M2[FL];
    const double rE = f_rsqrt(E2);
    const double E = E2 * rE;
    const double a = (c.mu - E) * c.invT;
    const double b = -(E + c.mu) * c.invT;
    const Species sp = species_eval(a, c.Phi3, c.Phib3);
    const Species sm = species_eval(b, c.Phib3, c.Phi3);
    const double np = f_fma(c.Phi, sp.r1, f_fma(2.0 * c.Phib, sp.r2, sp.r3));
    const double nm = f_fma(c.Phib, sm.r1, f_fma(2.0 * c.Phi, sm.r2, sm.r3));
    double L = f_log(sp.D * sm.D);
    if (sp.pos) L = f_fma(3.0, a, L);
    if (sm.pos) L = f_fma(3.0, b, L);
    acc[TH_NP + FL] = f_fma(coef, np, acc[TH_NP + FL]);
    acc[TH_NM + FL] = f_fma(coef, nm, acc[TH_NM + FL]);
    acc[TH_L] = f_fma(coef, L, acc[TH_L]);
    acc[TH_T] = f_fma(coef, f_fma(np, E - c.mu, nm * (E + c.mu)), acc[TH_T]);
}

struct Thermo {
    double omega, pressure, rho_norm, entropy, energy;
    double rho[3], nq[3], nqb[3], M[3];
};

PNJL_HD void finish_thermo(const Model& m, const PointCtx& c, const double x[5], const double acc[kThAcc], Thermo& th) {
    const double T = c.T;
    const UTerms u = polyakov_U(m, T, x[3], x[4]);
    const double chi = 2 * m.G * ((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]) - 4 * m.K * ((x[0] * x[1]) * x[2]);
    double vac = 0.0;
    for (int i = 0; i < 3; ++i) {
        double I0, I1, I2;
        vacuum_terms(m.Lambda, c.M[i], I0, I1, I2);
        vac += I0;
        th.M[i] = c.M[i];
    }
    const double omega = chi + u.U + (-2.0 * m.Nc) * vac + (-2.0 * T) * acc[TH_L];
    th.omega = omega;
    th.pressure = -omega;
    const double pref = 2.0 * m.Nc;
    double rsum = 0.0, murho = 0.0;
    for (int i = 0; i < 3; ++i) {
        th.nq[i] = pref * acc[TH_NP + i];
        th.nqb[i] = pref * acc[TH_NM + i];
        th.rho[i] = 6.0 * (acc[TH_NP + i] - acc[TH_NM + i]);  // 2T * (3/T) * sum c (n+ - n-)
        rsum += th.rho[i];
        murho += c.mu * th.rho[i];
    }
    th.rho_norm = rsum / (3.0 * m.rho0);
    // s = dP/dT at fixed x:  -dU/dT + 2 sum c L + 2T sum c dL/dT,  dL+/dT = 3 n+ (E - mu)/T^2
    th.entropy = -u.U_T + 2.0 * acc[TH_L] + 6.0 * c.invT * acc[TH_T];
    th.energy = -th.pressure + murho + T * th.entropy;
}

// ------------------------------------------------------------------------------------------------
// 5x5 dense algebra
// ------------------------------------------------------------------------------------------------
// Solve A y = b by LU with partial pivoting.  false on an exactly-zero pivot.
PNJL_HD_NOINL bool lu_solve5(const double A_in[25], const double b_in[5], double y[5]) {
    double A[25], b[5];
#pragma unroll
    for (int i = 0; i < 25; ++i) A[i] = A_in[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) b[i] = b_in[i];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int piv = k;
        double best = fabs(A[k * 5 + k]);
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            const double v = fabs(A[i * 5 + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) ok = false;
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            if (i == piv) {
#pragma unroll
                for (int j = 0; j < 5; ++j) { const double t = A[k * 5 + j]; A[k * 5 + j] = A[i * 5 + j]; A[i * 5 + j] = t; }
                const double t = b[k]; b[k] = b[i]; b[i] = t;
            }
        }
        const double inv = 1.0 / A[k * 5 + k];
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            const double l = A[i * 5 + k] * inv;
#pragma unroll
            for (int j = k + 1; j < 5; ++j) A[i * 5 + j] -= l * A[k * 5 + j];
            b[i] -= l * b[k];
        }
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        double s = b[i];
#pragma unroll
        for (int j = i + 1; j < 5; ++j) s -= A[i * 5 + j] * y[j];
        y[i] = s / A[i * 5 + i];
    }
    return ok;
}

PNJL_HD double norm_inf5(const double v[5]) {
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double a = fabs(v[i]);
        if (a > m || a != a) m = a;
    }
    return m;
}
PNJL_HD bool any_nan5(const double v[5]) {
    bool r = false;
#pragma unroll
    for (int i = 0; i < 5; ++i) r = r || (v[i] != v[i]);
    return r;
}
PNJL_HD bool finite_d(double v) { return fabs(v) <= 1.7976931348623157e308; }
PNJL_HD bool all_finite5(const double v[5]) {
    bool r = true;
#pragma unroll
    for (int i = 0; i < 5; ++i) r = r && finite_d(v[i]);
    return r;
}
PNJL_HD double wnorm5(const double d[5], const double v[5]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) { const double t = d[i] * v[i]; s += t * t; }
    return sqrt(s);
}

}  // namespace pnjl
