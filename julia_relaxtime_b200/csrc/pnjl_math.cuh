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
#define PNJL_HD_NOINL inline
#endif

#ifndef PNJL_FJ_UNROLL
#define PNJL_FJ_UNROLL 1
#endif

namespace pnjl {

constexpr int kFjUnroll = PNJL_FJ_UNROLL;   // unroll factor of the paired FJ loop (experiments; 1 in the product build)
constexpr double kPi = 3.14159265358979323846;
constexpr double kPolyakovEps = 1e-16;  // Integrals.jl:152

struct Model {
    double hbarc, Lambda, m_ud0, m_s0, G, K, T0, a0, a1, a2, b3, rho0;
    double Nc;
};

struct SolverParams {
    double xtol, ftol, residual_norm_max, phi_tol, omega_tie_rel;
    int max_iter, tr_fallback, auto_multiseed_fallback;
    int isospin;   // exploit M_u == M_d when phi_u == phi_d bitwise (mu_u = mu_d on this path)
    double predict_tol;   // a Newton pass after a residual <= predict_tol is run as a fused "final pass" (0: never)
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

// exp polynomial (degree 11, highest power first) and range-reduction constants.  On the device they live in
// constant memory so that every DFMA takes them as a constant-bank operand instead of pinning 2 registers each.
#if defined(__CUDACC__)
__constant__ double kExpC[12] = {0x1.af635e4f6b5eep-26, 0x1.28b43a93fe57ap-22, 0x1.71ddf5514be0cp-19, 0x1.a01991731e6fap-16,
                                 0x1.a01a01b150ad2p-13, 0x1.6c16c1881156bp-10, 0x1.111111110f205p-7,  0x1.555555554f067p-5,
                                 0x1.555555555555ap-3,  0x1.0000000000011p-1,  1.0,                   1.0};
__constant__ double kExpR[4] = {1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01,
                                -1.90821492927058770002e-10};
#endif

// ------------------------------------------------------------------------------------------------
// Branch-free FP64 primitives for the fast path (device: MUFU seed + Newton-Raphson in DFMA, a
// degree-11 polynomial exp; host build: libm).  Domain restrictions are guaranteed by the
// group-uniform fast-path test (fast_path_ok below): arguments are normal, positive where needed,
// and exp arguments lie in [-708, 0].  Keeping them branch-free lets ptxas interleave the independent
// flavour/node chains, which is what keeps the FP64 pipe fed.
// ------------------------------------------------------------------------------------------------
PNJL_HD double fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
    // MUFU seed (rel. error < 2^-22) + one third-order step: y (1 + e + e^2), e = 1 - x y  ->  error ~ e^3
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
#else
    return 1.0 / x;
#endif
}

PNJL_HD double fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    // MUFU seed + one third-order step: y (1 + e/2 + 3 e^2/8), e = 1 - x y^2
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
#else
    return 1.0 / sqrt(x);
#endif
}

// exp(t) for t in [-708, 0]: t = k ln2 + r, |r| <= ln2/2, degree-11 interpolant at Chebyshev nodes
// (max relative error 4.3e-18 before rounding), scaling by an integer add into the exponent field.
PNJL_HD double fast_exp_nonpos(double t) {
#if defined(__CUDA_ARCH__)
    const double kShift = 6755399441055744.0;  // 1.5 * 2^52
    double kd = fma(t, 1.4426950408889634, kShift);
    const int k = __double2loint(kd);
    kd -= kShift;
    double r = fma(kd, -6.93147180369123816490e-01, t);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 0x1.af635e4f6b5eep-26;
    p = fma(p, r, 0x1.28b43a93fe57ap-22);
    p = fma(p, r, 0x1.71ddf5514be0cp-19);
    p = fma(p, r, 0x1.a01991731e6fap-16);
    p = fma(p, r, 0x1.a01a01b150ad2p-13);
    p = fma(p, r, 0x1.6c16c1881156bp-10);
    p = fma(p, r, 0x1.111111110f205p-7);
    p = fma(p, r, 0x1.555555554f067p-5);
    p = fma(p, r, 0x1.555555555555ap-3);
    p = fma(p, r, 0x1.0000000000011p-1);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    return exp(t);
#endif
}

// ln(x) for positive, normal, finite x: exponent/mantissa split with integer ops, m in [sqrt(1/2), sqrt(2)),
// s = f/(2+f), odd series in s (the classic fdlibm e_log.c scheme).  No special cases, no branches.
PNJL_HD double fast_log_pos(double x) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x);
    const int hx = hi & 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double dk = (double)(((hi >> 20) - 1023) + (i >> 20));
    const double f = __hiloint2double(hx | (i ^ 0x3ff00000), __double2loint(x)) - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
#else
    return log(x);
#endif
}

// Guarded versions for arguments that are not known to be tame (the general integrand path, pivots of the elimination): the
// fast sequence for normal, finite arguments well inside the exponent range, the IEEE / libm routine out of line otherwise
// (NaN fails every comparison and goes there too).  Non-physical Newton iterates (MultiSeed bootstraps, Phi < 0) spend their
// sweeps here: libm exp + IEEE division + rsqrt were 30 % of the samples of a config-3 run.
PNJL_HD_NOINL double cold_rcp(double v) { return 1.0 / v; }   // out of line: IEEE division is ~100 instructions
PNJL_HD_NOINL double cold_exp(double t) { return exp(t); }
PNJL_HD_NOINL double cold_log(double x) { return log(x); }
PNJL_HD_NOINL double cold_rsqrt(double x) { return f_rsqrt(x); }
PNJL_HD double guarded_rcp(double v) {
    const double a = fabs(v);
    if (a > 1e-280 && a < 1e280) return fast_rcp(v);
    return cold_rcp(v);
}
PNJL_HD double guarded_rsqrt(double x) {
    if (x > 1e-280 && x < 1e280) return fast_rsqrt(x);
    return cold_rsqrt(x);
}
PNJL_HD double guarded_exp_nonpos(double t) {       // exp(t), t <= 0
    if (t >= -708.0) return fast_exp_nonpos(t);
    return cold_exp(t);
}
PNJL_HD double guarded_log(double x) {
    if (x > 1e-300 && x < 1e300) return fast_log_pos(x);
    return cold_log(x);
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

// Out of line on the device: the general-path loops call it six times per node, and inlined (with libm's exp) their bodies
// grow to tens of KB that no instruction cache level holds — 86 cycles per instruction were measured on the general path of
// the line-march kernel.  The fast path does not come here.
PNJL_HD_NOINL Species species_eval(double a, double P1x3, double P2x3) {
    Species s;
    const double w = guarded_exp_nonpos(-fabs(a));
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
    const double inv = guarded_rcp(D);
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
    c.T = T; c.mu = mu; c.xi = xi;
    c.invT = (T > 1e-300 && T < 1e300) ? fast_rcp(T) : 1.0 / T;
    c.Phi = x[3]; c.Phib = x[4]; c.Phi3 = 3.0 * x[3]; c.Phib3 = 3.0 * x[4];
    masses_of(m, x, c.M);
    for (int i = 0; i < 3; ++i) c.M2[i] = c.M[i] * c.M[i];
}

// One node x one flavour of the FJ pass.  k2 = p^2 + xi (p cos)^2, coef = quadrature coefficient.
template <int FL>
PNJL_HD void fj_node(const PointCtx& c, double k2, double coef, double acc[kFJAcc]) {
    const double E2 = k2 + c.M2[FL];
    const double rE = guarded_rsqrt(E2);
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
// Fast path of the quadrature passes.  Valid (and selected group-uniformly per evaluation) when
//   Phi >= 0 and Phibar >= 0        -> f+- >= 1, so the reference's max(., 1e-16) floors cannot be active,
//   (E_max + |mu|) / T <= 200       -> y^3 = e^{3a} <= e^{600} stays finite, so the reference's a>0 rescaling
//                                      (Integrals.jl:203-218), which only guards against overflow, is the identity.
// Then with e1 = exp(-E/T):  y = e^{-(E-mu)/T} = e1 * e^{mu/T},  z = e^{-(E+mu)/T} = e1 * e^{-mu/T}  (ONE exp per
// node x flavour), f+ = 1 + y (3 Phi + y (3 Phibar + y)), f- likewise with z and Phi <-> Phibar.
// Every other evaluation (non-physical iterates, T -> 0) takes the general path above.
// ------------------------------------------------------------------------------------------------
struct FastCtx {
    double nInvT;         // -1/T
    double nInvT_l2e;     // -log2(e)/T
    double kapP, kapM;    // e^{+mu/T}, e^{-mu/T}
    double Phi, Phib, Phi3, Phib3, Phi2, Phib2, Phi4, Phib4;
};

PNJL_HD bool fast_path_ok(double T, double mu, double Phi, double Phib, double k2max, const double M2[3]) {
    double m2 = M2[0] > M2[1] ? M2[0] : M2[1];
    m2 = m2 > M2[2] ? m2 : M2[2];
    const double r = 200.0 * T - fabs(mu);          // E_max <= r  <=>  (E_max + |mu|) / T <= 200
    return (Phi >= 0.0) && (Phib >= 0.0) && (T > 1e-300) && (r > 0.0) && (k2max + m2 <= r * r);
}

// |mu|/T <= 60: the product (f_u+ f_u-)^2 f_s+ f_s- of one node stays far below DBL_MAX (see th_pair_fast).
PNJL_HD bool one_log_ok(double T, double mu) { return fabs(mu) <= 60.0 * T; }

PNJL_HD void make_fast_ctx(const PointCtx& c, FastCtx& f) {
    f.nInvT = -c.invT;
    f.nInvT_l2e = -c.invT * 1.4426950408889634;
    // |mu|/T <= 200 here, so e^{-|mu|/T} is a normal number and its reciprocal is safe
    const double km = fast_exp_nonpos(-fabs(c.mu) * c.invT);
    const double kp = fast_rcp(km);
    f.kapP = c.mu >= 0.0 ? kp : km;
    f.kapM = c.mu >= 0.0 ? km : kp;
    f.Phi = c.Phi; f.Phib = c.Phib;
    f.Phi3 = c.Phi3; f.Phib3 = c.Phib3;
    f.Phi2 = 2.0 * c.Phi; f.Phib2 = 2.0 * c.Phib;
    f.Phi4 = 4.0 * c.Phi; f.Phib4 = 4.0 * c.Phib;
}

// One species on the fast path: y = e^a; outputs n = g/f, qf = q/f, r1 = y/f, r2 = y^2/f.
//   f = 1 + y (3 P1 + y (3 P2 + y)),  g = y (P1 + y (2 P2 + y)),  q = y (P1 + y (4 P2 + 3 y))
PNJL_HD void species_fast(double y, double P1, double P1x3, double P2x2, double P2x3, double P2x4, double& n, double& qf,
                          double& r1, double& r2, double& fval) {
    const double f = f_fma(y, f_fma(y, y + P2x3, P1x3), 1.0);
    const double g = f_fma(y, y + P2x2, P1);
    const double q = f_fma(y, f_fma(3.0, y, P2x4), P1);
    const double inv = fast_rcp(f);
    r1 = y * inv;
    r2 = r1 * y;
    n = g * r1;
    qf = q * r1;
    fval = f;
}

// One node x one flavour of the FJ pass on the fast path.
//   fl[5] = per-flavour sums {S1, S2A, S2B', S3, S4} with S2B' = sum c (n+ + n-)/E^3
//           (k^2/E^3 = 1/E - M^2/E^3 is folded in finish_fj, flag `fast`)
//   sh[5] = flavour-summed sums {GP, GPB, HPP, HPPB, HPBPB}
PNJL_HD void fj_node_fast(const FastCtx& fc, double M2, double k2, double coef, double fl[5], double sh[5]) {
    const double E2 = k2 + M2;
    const double rE = fast_rsqrt(E2);
    const double E = E2 * rE;
    const double e1 = fast_exp_nonpos(E * fc.nInvT);
    const double y = e1 * fc.kapP;
    const double z = e1 * fc.kapM;
    double np, qp, r1p, r2p, fp, nm, qm, r1m, r2m, fm;
    species_fast(y, fc.Phi, fc.Phi3, fc.Phib2, fc.Phib3, fc.Phib4, np, qp, r1p, r2p, fp);
    species_fast(z, fc.Phib, fc.Phib3, fc.Phi2, fc.Phi3, fc.Phi4, nm, qm, r1m, r2m, fm);
    const double nsum = np + nm;
    const double m3p = -3.0 * np, m3m = -3.0 * nm;
    const double Q = f_fma(m3p, np, qp) + f_fma(m3m, nm, qm);
    const double crE = coef * rE;
    const double crE2 = crE * rE;
    fl[0] = f_fma(crE, nsum, fl[0]);
    fl[1] = f_fma(crE2, Q, fl[1]);
    fl[2] = f_fma(crE2 * rE, nsum, fl[2]);
    fl[3] = f_fma(crE, f_fma(r1p, 1.0 + m3p, r2m * (2.0 + m3m)), fl[3]);
    fl[4] = f_fma(crE, f_fma(r2p, 2.0 + m3p, r1m * (1.0 + m3m)), fl[4]);
    sh[0] = f_fma(coef, r1p + r2m, sh[0]);
    sh[1] = f_fma(coef, r2p + r1m, sh[1]);
    sh[2] = f_fma(coef, f_fma(r1p, r1p, r2m * r2m), sh[2]);
    sh[3] = f_fma(coef, f_fma(r1p, r2p, r2m * r1m), sh[3]);
    sh[4] = f_fma(coef, f_fma(r2p, r2p, r1m * r1m), sh[4]);
}

// Thermo pass, fast path: th[4] = {sum c n+, sum c n-, sum c (L+ + L-), sum c [n+ (E-mu) + n- (E+mu)]} of one flavour.
PNJL_HD void thermo_node_fast(const FastCtx& fc, double mu, double M2, double k2, double coef, double th[4]) {
    const double E2 = k2 + M2;
    const double rE = fast_rsqrt(E2);
    const double E = E2 * rE;
    const double e1 = fast_exp_nonpos(E * fc.nInvT);
    const double y = e1 * fc.kapP;
    const double z = e1 * fc.kapM;
    double np, qp, r1p, r2p, fp, nm, qm, r1m, r2m, fm;
    species_fast(y, fc.Phi, fc.Phi3, fc.Phib2, fc.Phib3, fc.Phib4, np, qp, r1p, r2p, fp);
    species_fast(z, fc.Phib, fc.Phib3, fc.Phi2, fc.Phi3, fc.Phi4, nm, qm, r1m, r2m, fm);
    const double L = fast_log_pos(fp * fm);
    th[0] = f_fma(coef, np, th[0]);
    th[1] = f_fma(coef, nm, th[1]);
    th[2] = f_fma(coef, L, th[2]);
    th[3] = f_fma(coef, f_fma(np, E - mu, nm * (E + mu)), th[3]);
}

// "Final pass" on the fast path: the residual F (no second derivatives) together with the thermo sums, for the
// Newton pass that is expected to satisfy the stopping rule (NLsolve itself evaluates only F after a step,
// ImplicitSolver.jl:112 -> newton_: value!(df, x)).  ft[7] of one flavour =
//   {S1 = sum c (n+ + n-)/E,  GP = sum c (r1+ + r2-),  GPB = sum c (r2+ + r1-),  sum c n+,  sum c n-,
//    sum c ln(f+ f-),  sum c [n+ (E - mu) + n- (E + mu)]}
PNJL_HD void ft_node_fast(const FastCtx& fc, double mu, double M2, double k2, double coef, double ft[7]) {
    const double E2 = k2 + M2;
    const double rE = fast_rsqrt(E2);
    const double E = E2 * rE;
    const double e1 = fast_exp_nonpos(E * fc.nInvT);
    const double y = e1 * fc.kapP;
    const double z = e1 * fc.kapM;
    double np, qp, r1p, r2p, fp, nm, qm, r1m, r2m, fm;
    species_fast(y, fc.Phi, fc.Phi3, fc.Phib2, fc.Phib3, fc.Phib4, np, qp, r1p, r2p, fp);
    species_fast(z, fc.Phib, fc.Phib3, fc.Phi2, fc.Phi3, fc.Phi4, nm, qm, r1m, r2m, fm);
    const double L = fast_log_pos(fp * fm);
    ft[0] = f_fma(coef * rE, np + nm, ft[0]);
    ft[1] = f_fma(coef, r1p + r2m, ft[1]);
    ft[2] = f_fma(coef, r2p + r1m, ft[2]);
    ft[3] = f_fma(coef, np, ft[3]);
    ft[4] = f_fma(coef, nm, ft[4]);
    ft[5] = f_fma(coef, L, ft[5]);
    ft[6] = f_fma(coef, f_fma(np, E - mu, nm * (E + mu)), ft[6]);
}

// ------------------------------------------------------------------------------------------------
// Note on the FP64 pipe: a DFMA whose three sources are three different vector registers occupies the pipe for 3 cycles
// instead of 2 (register-file operand bandwidth; scripts/microbench_dfma_operands*.cu measure 0.333 instead of 0.5 warp
// instructions per cycle per scheduler), while immediates, constant-bank / uniform-register operands, a repeated register
// or a reuse-cache hit are free.  44 of the 177 FP64 instructions of the paired FJ loop are of the slow kind
// (accumulations, mixed products), which puts the loop's own ceiling at 354 / 398 = 89 % of the DFMA peak.
// Paired fast path (isospin case): the u and the s flavour of one node are advanced step by step together, so
// that two (four, inside the species block) independent dependency chains are adjacent in program order.  A
// DFMA result is available after 8 cycles and a warp can issue one every 2, so a lone chain leaves the FP64 pipe
// idle three quarters of the time (scripts/microbench_dfma.cu); ptxas keeps the chains apart when they are
// written one after the other.  Same arithmetic as fast_rsqrt / fast_exp_nonpos / fast_rcp / species_fast.
// ------------------------------------------------------------------------------------------------
template <int W>
PNJL_HD void v_rsqrt(const double x[W], double y[W]) {
#if defined(__CUDA_ARCH__)
    double e[W], p[W], t[W];
#pragma unroll
    for (int j = 0; j < W; ++j) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[j]) : "d"(x[j]));
#pragma unroll
    for (int j = 0; j < W; ++j) t[j] = x[j] * y[j];
#pragma unroll
    for (int j = 0; j < W; ++j) e[j] = fma(-t[j], y[j], 1.0);
#pragma unroll
    for (int j = 0; j < W; ++j) p[j] = fma(0.375, e[j], 0.5);
    // y (1 + e p): the correction factor is a two-register DFMA (e, p, immediate 1) and the update a DMUL, where
    // y + (y e) p needs a DFMA with three distinct register operands (3 FP64-pipe cycles instead of 2, see the note above)
#pragma unroll
    for (int j = 0; j < W; ++j) t[j] = fma(e[j], p[j], 1.0);
#pragma unroll
    for (int j = 0; j < W; ++j) y[j] = y[j] * t[j];
#else
    for (int j = 0; j < W; ++j) y[j] = 1.0 / sqrt(x[j]);
#endif
}

template <int W>
PNJL_HD void v_rcp(const double x[W], double y[W]) {
#if defined(__CUDA_ARCH__)
    double e[W];
#pragma unroll
    for (int j = 0; j < W; ++j) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[j]) : "d"(x[j]));
#pragma unroll
    for (int j = 0; j < W; ++j) e[j] = fma(-x[j], y[j], 1.0);
#pragma unroll
    for (int j = 0; j < W; ++j) e[j] = fma(e[j], e[j], e[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) y[j] = fma(y[j], e[j], y[j]);
#else
    for (int j = 0; j < W; ++j) y[j] = 1.0 / x[j];
#endif
}

// (A table-driven exp — 32-entry 2^(j/32) table in shared memory + degree-6 polynomial, 11 FP64 instructions instead of
// 15 — was measured 3 % SLOWER on cfg5: the LDS and the index arithmetic sit on the critical chain of the front end.)
// exp(t[j]) for t[j] = E[j] * slope (<= 0); slope_l2e = slope * log2(e) is supplied by the caller (a per-pass constant), so that
// the rounding step k = rint(t log2 e) is a DFMA of two registers and a constant instead of three registers.
template <int W>
PNJL_HD void v_exp_nonpos(const double t[W], const double E[W], double slope_l2e, double out[W]) {
#if defined(__CUDA_ARCH__)
    double kd[W], r[W], p[W];
    int k[W];
#pragma unroll
    for (int j = 0; j < W; ++j) kd[j] = fma(E[j], slope_l2e, kExpR[1]);
#pragma unroll
    for (int j = 0; j < W; ++j) k[j] = __double2loint(kd[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) kd[j] -= kExpR[1];
#pragma unroll
    for (int j = 0; j < W; ++j) r[j] = fma(kd[j], kExpR[2], t[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) r[j] = fma(kd[j], kExpR[3], r[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) p[j] = fma(kExpC[0], r[j], kExpC[1]);
#pragma unroll
    for (int q = 2; q < 12; ++q) {
#pragma unroll
        for (int j = 0; j < W; ++j) p[j] = fma(p[j], r[j], kExpC[q]);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) out[j] = __hiloint2double(__double2hiint(p[j]) + (k[j] << 20), __double2loint(p[j]));
#else
    for (int j = 0; j < W; ++j) out[j] = exp(t[j]);
#endif
}

// Four species (u quark, u antiquark, s quark, s antiquark) evaluated together.  Y[4] = {y_u, z_u, y_s, z_s}.
// Even entries use (P1, P2) = (Phi, Phibar), odd entries (Phibar, Phi).
struct Species4 {
    double n[4], qf[4], r1[4], r2[4], f[4];
};
PNJL_HD void species4_fast(const FastCtx& fc, const double Y[4], Species4& o, bool need_q) {
    double g[4], q[4], inv[4];
    // species visited in the order quark u, quark s, antiquark u, antiquark s: consecutive DFMAs then share their
    // loop-invariant addend (3 Phi, then 3 Phibar; Phi, then Phibar) in the same operand slot, where the register reuse cache
    // serves it and the instruction reads two fresh registers instead of three (2 pipe cycles instead of 3)
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const int s = (q4 >> 1) | ((q4 & 1) << 1);
        const double P13 = (s & 1) ? fc.Phib3 : fc.Phi3, P2 = (s & 1) ? fc.Phi : fc.Phib;
        o.f[s] = f_fma(Y[s], f_fma(Y[s], f_fma(3.0, P2, Y[s]), P13), 1.0);
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const int s = (q4 >> 1) | ((q4 & 1) << 1);
        const double P1 = (s & 1) ? fc.Phib : fc.Phi, P2 = (s & 1) ? fc.Phi : fc.Phib;
        g[s] = f_fma(Y[s], f_fma(2.0, P2, Y[s]), P1);
    }
    if (need_q) {
        // q - g = 2 y (P2 + y) in these (y-stripped) forms, so q/f = n + r2 (2 P2 + 2 y): two instructions instead of three
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const double P22 = (s & 1) ? fc.Phi2 : fc.Phib2;
            q[s] = f_fma(2.0, Y[s], P22);
        }
    }
    v_rcp<4>(o.f, inv);
#pragma unroll
    for (int s = 0; s < 4; ++s) o.r1[s] = Y[s] * inv[s];
#pragma unroll
    for (int s = 0; s < 4; ++s) o.r2[s] = o.r1[s] * Y[s];
#pragma unroll
    for (int s = 0; s < 4; ++s) o.n[s] = g[s] * o.r1[s];
    if (need_q) {
        // qf here is (q - g)/f = r2 (2 P2 + 2 y); the caller forms (q - 3 g n)/f = u n + qf with u = 1 - 3 n
#pragma unroll
        for (int s = 0; s < 4; ++s) o.qf[s] = q[s] * o.r2[s];
    }
}

// Common front end of the paired passes: E, 1/E and the four Boltzmann factors of the u and s flavour of one node.
PNJL_HD void pair_front(const FastCtx& fc, double M2u, double M2s, double k2, double rE[2], double E[2], double Y[4]) {
    double E2[2] = {k2 + M2u, k2 + M2s}, t[2], e1[2];
    v_rsqrt<2>(E2, rE);
#pragma unroll
    for (int j = 0; j < 2; ++j) E[j] = E2[j] * rE[j];
#pragma unroll
    for (int j = 0; j < 2; ++j) t[j] = E[j] * fc.nInvT;
    v_exp_nonpos<2>(t, E, fc.nInvT_l2e, e1);
#pragma unroll
    for (int j = 0; j < 2; ++j) { Y[2 * j] = e1[j] * fc.kapP; Y[2 * j + 1] = e1[j] * fc.kapM; }
}

// FJ pass, u and s flavour of one node, after the front end.  fu/fs: per-flavour sums as in fj_node_fast; sh:
// flavour-summed sums with the u contribution counted twice (u and d are the same flavour here).
PNJL_HD void fj_pair_back(const FastCtx& fc, const double rE[2], const double Y[4], double coef, double fu[5], double fs[5],
                          double sh[5]) {
    Species4 sp;
    species4_fast(fc, Y, sp, true);
    double nsum[2], Q[2], crE[2], crE2[2], u[4], s3[2], s4[2], gp[2], gpb[2], hpp[2], hppb[2], hpbpb[2];
    // u = 1 - 3 n, v = 2 - 3 n = u + 1;  (q - 3 g n)/f = q/f - 3 n^2 = u n + (q - g)/f;  r v = r u + r (one DFMA of two registers)
#pragma unroll
    for (int s = 0; s < 4; ++s) u[s] = f_fma(-3.0, sp.n[s], 1.0);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int a = 2 * j, b = 2 * j + 1;   // quark, antiquark
        nsum[j] = sp.n[a] + sp.n[b];
        Q[j] = f_fma(u[a], sp.n[a], sp.qf[a]) + f_fma(u[b], sp.n[b], sp.qf[b]);
        crE[j] = coef * rE[j];
        crE2[j] = crE[j] * rE[j];
        s3[j] = f_fma(sp.r1[a], u[a], f_fma(sp.r2[b], u[b], sp.r2[b]));
        s4[j] = f_fma(sp.r1[b], u[b], f_fma(sp.r2[a], u[a], sp.r2[a]));
        gp[j] = sp.r1[a] + sp.r2[b];
        gpb[j] = sp.r2[a] + sp.r1[b];
        hpp[j] = f_fma(sp.r1[a], sp.r1[a], sp.r2[b] * sp.r2[b]);
        hppb[j] = f_fma(sp.r1[a], sp.r2[a], sp.r2[b] * sp.r1[b]);
        hpbpb[j] = f_fma(sp.r2[a], sp.r2[a], sp.r1[b] * sp.r1[b]);
    }
    fu[0] = f_fma(crE[0], nsum[0], fu[0]);           fs[0] = f_fma(crE[1], nsum[1], fs[0]);
    fu[1] = f_fma(crE2[0], Q[0], fu[1]);             fs[1] = f_fma(crE2[1], Q[1], fs[1]);
    fu[2] = f_fma(crE2[0] * rE[0], nsum[0], fu[2]);  fs[2] = f_fma(crE2[1] * rE[1], nsum[1], fs[2]);
    fu[3] = f_fma(crE[0], s3[0], fu[3]);             fs[3] = f_fma(crE[1], s3[1], fs[3]);
    fu[4] = f_fma(crE[0], s4[0], fu[4]);             fs[4] = f_fma(crE[1], s4[1], fs[4]);
    sh[0] = f_fma(coef, f_fma(2.0, gp[0], gp[1]), sh[0]);
    sh[1] = f_fma(coef, f_fma(2.0, gpb[0], gpb[1]), sh[1]);
    sh[2] = f_fma(coef, f_fma(2.0, hpp[0], hpp[1]), sh[2]);
    sh[3] = f_fma(coef, f_fma(2.0, hppb[0], hppb[1]), sh[3]);
    sh[4] = f_fma(coef, f_fma(2.0, hpbpb[0], hpbpb[1]), sh[4]);
}

PNJL_HD void fj_pair_fast(const FastCtx& fc, double M2u, double M2s, double k2, double coef, double fu[5], double fs[5],
                          double sh[5]) {
    double rE[2], E[2], Y[4];
    pair_front(fc, M2u, M2s, k2, rE, E, Y);
    fj_pair_back(fc, rE, Y, coef, fu, fs, sh);
}

// Thermo pass / fused final pass, u and s flavour of one node.  tu/ts: {sum c n+, sum c n-, sum c ln(f+ f-),
// sum c [n+ (E-mu) + n- (E+mu)]} per flavour; when WITH_F also s1u/s1s (sum c (n+ + n-)/E) and the flavour-summed
// gsh[2] = {GP, GPB} (u counted twice).
// ONE_LOG: the u flavour counts twice (u == d), so the log sum of a node is 2 ln(f_u+ f_u-) + ln(f_s+ f_s-) =
// ln((f_u+ f_u-)^2 f_s+ f_s-): one logarithm per node instead of two, accumulated in ts[2] alone (tu[2] stays 0).
// The caller selects it (uniformly) only when the product cannot overflow: every f is <= 8 max(1, e^{3|mu|/T}), so
// |mu|/T <= 60 bounds the product by 2^18 e^{540}.
template <bool WITH_F, bool ONE_LOG>
PNJL_HD void th_pair_fast(const FastCtx& fc, double mu, double M2u, double M2s, double k2, double coef, double tu[4],
                          double ts[4], double& s1u, double& s1s, double gsh[2]) {
    double rE[2], E[2], Y[4];
    pair_front(fc, M2u, M2s, k2, rE, E, Y);
    Species4 sp;
    species4_fast(fc, Y, sp, false);
    tu[0] = f_fma(coef, sp.n[0], tu[0]);  ts[0] = f_fma(coef, sp.n[2], ts[0]);
    tu[1] = f_fma(coef, sp.n[1], tu[1]);  ts[1] = f_fma(coef, sp.n[3], ts[1]);
    if (ONE_LOG) {
        const double fu = sp.f[0] * sp.f[1];
        const double L = fast_log_pos((fu * fu) * (sp.f[2] * sp.f[3]));
        ts[2] = f_fma(coef, L, ts[2]);
    } else {
        const double Lu = fast_log_pos(sp.f[0] * sp.f[1]);
        const double Ls = fast_log_pos(sp.f[2] * sp.f[3]);
        tu[2] = f_fma(coef, Lu, tu[2]);
        ts[2] = f_fma(coef, Ls, ts[2]);
    }
    // n+ (E - mu) + n- (E + mu) = E (n+ + n-) - mu (n+ - n-): the loop accumulates sum c E (n+ + n-) only, the caller
    // subtracts mu (sum c n+ - sum c n-) from the slots it has anyway (th_pair_finish)
    const double nsu = sp.n[0] + sp.n[1], nss = sp.n[2] + sp.n[3];
    tu[3] = f_fma(coef, E[0] * nsu, tu[3]);
    ts[3] = f_fma(coef, E[1] * nss, ts[3]);
    if (WITH_F) {
        s1u = f_fma(coef * rE[0], nsu, s1u);
        s1s = f_fma(coef * rE[1], nss, s1s);
        gsh[0] = f_fma(coef, f_fma(2.0, sp.r1[0] + sp.r2[1], sp.r1[2] + sp.r2[3]), gsh[0]);
        gsh[1] = f_fma(coef, f_fma(2.0, sp.r2[0] + sp.r1[1], sp.r2[2] + sp.r1[3]), gsh[1]);
    }
}

// After a th_pair_fast loop: turn slot 3 into sum c [n+ (E - mu) + n- (E + mu)] (see there).  Linear, so it may be applied
// to per-lane partial sums.
PNJL_HD void th_pair_finish(double mu, double tu[4], double ts[4]) {
    tu[3] = f_fma(-mu, tu[0] - tu[1], tu[3]);
    ts[3] = f_fma(-mu, ts[0] - ts[1], ts[3]);
}

// Mesh slice seen by one lane: nodes lane, lane+stride, ...   (host build: lane 0, stride 1)
struct MeshView {
    const double* p2;     // p^2
    const double* pc2;    // (p cos)^2
    const double* coef;   // w_p * 2 w_c * p^2 / (2 pi)^2
    int n;
    double p2max, pc2max;
    // Isotropic collapse (n_iso > 0 enables it): for xi == 0 the integrand does not depend on cos(theta), so
    //   sum_ij coef_ij g(p_i) = sum_i (sum_j coef_ij) g(p_i)
    // and a pass needs p_num nodes instead of p_num * t_num.  p2_iso[i] = p_i^2, coef_iso[i] = sum_j coef_ij (summed on
    // the host in j order).  Same value up to the summation order (<= a few ulp of the sums).
    const double* p2_iso;
    const double* coef_iso;
    int n_iso;
};

// The mesh a pass sweeps for anisotropy xi (group-uniform choice).
PNJL_HD MeshView select_mesh(const MeshView& mv, double xi) {
    MeshView r = mv;
    if (xi == 0.0 && mv.n_iso > 0) {
        r.p2 = mv.p2_iso;
        r.pc2 = mv.p2_iso;     // multiplied by xi == 0
        r.coef = mv.coef_iso;
        r.n = mv.n_iso;
    }
    return r;
}

// General path of an FJ pass (floors / rescaling live), out of line: its code and its registers stay out of the fast-path
// loops of the callers (the warp-specialised kernel's workers run those and nothing else in steady state).
// iso: M_u == M_d bitwise, the d flavour is the u flavour.  u is then evaluated with twice the coefficient — the
// flavour-summed slots hold 2 c (u terms) + c (s terms) — and its own slots are halved afterwards (exact: powers of two).
PNJL_HD_NOINL void fj_partial_general(const PointCtx& c, bool iso, const MeshView& mv, int lane, int stride, double* out) {
    double acc[kFJAcc];
#pragma unroll
    for (int i = 0; i < kFJAcc; ++i) acc[i] = 0.0;
    if (iso) {
#pragma unroll 1
        for (int k = lane; k < mv.n; k += stride) {
            const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
            const double cf = mv.coef[k];
            fj_node<0>(c, k2, 2.0 * cf, acc);
            fj_node<2>(c, k2, cf, acc);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            acc[3 * q + 0] *= 0.5;
            acc[3 * q + 1] = acc[3 * q + 0];
        }
    } else {
#pragma unroll 1
        for (int k = lane; k < mv.n; k += stride) {
            const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
            const double cf = mv.coef[k];
            fj_node<0>(c, k2, cf, acc);
            fj_node<1>(c, k2, cf, acc);
            fj_node<2>(c, k2, cf, acc);
        }
    }
#pragma unroll
    for (int i = 0; i < kFJAcc; ++i) out[i] = acc[i];
}

// Per-lane partial sums of one FJ pass in the canonical 20-slot layout (to be summed over the lanes of
// the group, then finish_fj).  Chooses, uniformly for the whole group, between
//   fast + isospin (x[0] == x[1] bitwise -> M_u == M_d bitwise: the d flavour is the u flavour, 2 flavours evaluated),
//   fast, three flavours,
//   general path (floors / rescaling live).
// Returns true when the fast-path slot convention (S2B') is in use.
PNJL_HD bool fj_partial(const Model& m, bool isospin, const PointCtx& c, const double x[5], const MeshView& mv_in, int lane,
                        int stride, double acc[kFJAcc]) {
    const MeshView mv = select_mesh(mv_in, c.xi);
    const double k2max = mv.p2max + (c.xi > 0.0 ? c.xi * mv.pc2max : 0.0);
    if (fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) {
        FastCtx fc;
        make_fast_ctx(c, fc);
        if (isospin && x[0] == x[1]) {
            double fu[5] = {0, 0, 0, 0, 0}, fs[5] = {0, 0, 0, 0, 0}, sh[5] = {0, 0, 0, 0, 0};
#pragma unroll kFjUnroll
            for (int k = lane; k < mv.n; k += stride) {
                const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                fj_pair_fast(fc, c.M2[0], c.M2[2], k2, mv.coef[k], fu, fs, sh);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                acc[3 * q + 0] = fu[q];
                acc[3 * q + 1] = fu[q];
                acc[3 * q + 2] = fs[q];
                acc[15 + q] = sh[q];
            }
        } else {
            double f0[5] = {0, 0, 0, 0, 0}, f1[5] = {0, 0, 0, 0, 0}, f2[5] = {0, 0, 0, 0, 0}, sh[5] = {0, 0, 0, 0, 0};
            for (int k = lane; k < mv.n; k += stride) {
                const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                const double cf = mv.coef[k];
                fj_node_fast(fc, c.M2[0], k2, cf, f0, sh);
                fj_node_fast(fc, c.M2[1], k2, cf, f1, sh);
                fj_node_fast(fc, c.M2[2], k2, cf, f2, sh);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                acc[3 * q + 0] = f0[q];
                acc[3 * q + 1] = f1[q];
                acc[3 * q + 2] = f2[q];
                acc[15 + q] = sh[q];
            }
        }
        return true;
    }
    fj_partial_general(c, isospin && x[0] == x[1], mv, lane, stride, acc);
    return false;
}

// ------------------------------------------------------------------------------------------------
// Closed-form pieces
// ------------------------------------------------------------------------------------------------
// Vacuum integral I(Lambda, m) and its first two derivatives in m = |M| + 1e-12.
//   16 pi^2 I   = L s (2 L^2 + m^2) - m^4 ln((L + s)/m),   s = sqrt(L^2 + m^2)
//    4 pi^2 I'  = L m s - m^3 ln((L + s)/m)
//    4 pi^2 I'' = L s + 2 L m^2 / s - 3 m^2 ln((L + s)/m)
// ASSUME_TAME: the caller has checked vacuum_tame(); the libm fall-backs are then not even compiled in (the line-march
// kernel keeps its per-pass code small that way).  Same arithmetic either way.
PNJL_HD bool vacuum_tame(double Lam, double M) {
    const double m = fabs(M) + 1e-12;
    return (m < 1e100) && (Lam > 1e-100) && (Lam < 1e100);   // everything in vacuum_terms is then a positive normal number
}
template <bool ASSUME_TAME>
PNJL_HD void vacuum_terms_t(double Lam, double M, double& I0, double& I1, double& I2) {
    const double m = fabs(M) + 1e-12;
    const double m2 = m * m;
    const double s2 = Lam * Lam + m2;
    const bool tame = ASSUME_TAME || ((m < 1e100) && (Lam > 1e-100) && (Lam < 1e100));
    double rs, lg;
    if (ASSUME_TAME) rs = fast_rsqrt(s2);
    else rs = tame ? fast_rsqrt(s2) : 1.0 / sqrt(s2);
    const double s = s2 * rs;
    if (ASSUME_TAME) lg = fast_log_pos((Lam + s) * fast_rcp(m));
    else lg = tame ? fast_log_pos((Lam + s) * fast_rcp(m)) : log((Lam + s) / m);
    I0 = (Lam * s * (2 * (Lam * Lam) + m2) - (m2 * m2) * lg) * (1.0 / (16 * (kPi * kPi)));
    const double sgn = (M > 0.0) ? 1.0 : ((M < 0.0) ? -1.0 : 0.0);
    I1 = sgn * (Lam * m * s - m2 * m * lg) * (1.0 / (4 * (kPi * kPi)));
    I2 = (Lam * s + 2 * Lam * m2 * rs - 3 * m2 * lg) * (1.0 / (4 * (kPi * kPi)));
}
PNJL_HD void vacuum_terms(double Lam, double M, double& I0, double& I1, double& I2) { vacuum_terms_t<false>(Lam, M, I0, I1, I2); }

struct UTerms {
    double U, U_P, U_Pb, U_PP, U_PPb, U_PbPb, U_T;
};

// U = T^4 [ -1/2 A(T) Phi Phibar + B(T) ln v ],  v = 1 - 6 Phi Phibar + 4 (Phi^3 + Phibar^3) - 3 (Phi Phibar)^2,
// A = a0 + a1 t + a2 t^2, B = b3 t^3, t = T0/T;  ln v is floored at ln 1e-16 (safe_log) -> zero derivative.
// WITH_VALUE = false: only the first and second derivatives in (Phi, Phibar) — what an Omega-gradient/Jacobian pass needs, no
// logarithm; U and U_T are left untouched.  ASSUME_TAME: the caller has checked polyakov_tame() (v is a positive normal number
// above the floor); the libm fall-backs are then not compiled in.  Same arithmetic in every variant.
PNJL_HD double polyakov_v(double P, double Pb) {
    const double PPb = Pb * P;
    return 1 - 6 * PPb + 4 * (Pb * Pb * Pb + P * P * P) - 3 * (PPb * PPb);
}
PNJL_HD bool polyakov_tame(double P, double Pb) {
    const double v = polyakov_v(P, Pb);
    return !(v <= 0.0) && !(v < kPolyakovEps) && (v < 1e100);
}
template <bool ASSUME_TAME, bool WITH_VALUE>
PNJL_HD void polyakov_eval(const Model& m, double T, double iT, double P, double Pb, UTerms& u) {
    const double t = m.T0 * iT;
    const double A = m.a0 + m.a1 * t + m.a2 * (t * t);
    const double B = m.b3 * (t * t * t);
    const double T2 = T * T, T4 = T2 * T2;
    const double PPb = Pb * P;
    const double v = 1 - 6 * PPb + 4 * (Pb * Pb * Pb + P * P * P) - 3 * (PPb * PPb);
    double lv = 0.0, iv;
    if (ASSUME_TAME) {
        if (WITH_VALUE) lv = fast_log_pos(v);
        iv = fast_rcp(v);
    } else {
        const bool live = !(v <= 0.0) && !(v < kPolyakovEps);
        const bool tame = live && (v < 1e100);
        if (WITH_VALUE) lv = tame ? fast_log_pos(v) : (live ? log(v) : log(kPolyakovEps));
        iv = tame ? fast_rcp(v) : (live ? 1.0 / v : 0.0);
    }
    const double vP = -6 * Pb + 12 * P * P - 6 * P * Pb * Pb;
    const double vPb = -6 * P + 12 * Pb * Pb - 6 * P * P * Pb;
    const double vPP = 24 * P - 6 * Pb * Pb;
    const double vPbPb = 24 * Pb - 6 * P * P;
    const double vPPb = -6 - 12 * PPb;
    u.U_P = T4 * (-0.5 * A * Pb + B * vP * iv);
    u.U_Pb = T4 * (-0.5 * A * P + B * vPb * iv);
    u.U_PP = T4 * B * (vPP * iv - (vP * iv) * (vP * iv));
    u.U_PbPb = T4 * B * (vPbPb * iv - (vPb * iv) * (vPb * iv));
    u.U_PPb = T4 * (-0.5 * A + B * (vPPb * iv - (vP * iv) * (vPb * iv)));
    if (WITH_VALUE) {
        u.U = T4 * (-0.5 * A * PPb + B * lv);
        // dU/dT at fixed Phi: Thermodynamics.jl:147-165
        const double iT2 = iT * iT;
        const double dA = -m.a1 * m.T0 * iT2 - 2 * m.a2 * (m.T0 * m.T0) * (iT2 * iT);
        const double dB = -3 * m.b3 * (m.T0 * m.T0 * m.T0) * (iT2 * iT2);
        u.U_T = 4 * (T2 * T) * (-0.5 * A * PPb + B * lv) + T4 * (dA * (-0.5 * PPb) + dB * lv);
    }
}
PNJL_HD void polyakov_U(const Model& m, double T, double iT, double P, double Pb, UTerms& u) {
    polyakov_eval<false, true>(m, T, iT, P, Pb, u);
}
PNJL_HD void polyakov_derivs(const Model& m, double T, double iT, double P, double Pb, UTerms& u) {
    polyakov_eval<false, false>(m, T, iT, P, Pb, u);
}

// Assemble F = grad_x P (5) and J = Hess_x P (5x5 row-major) from the reduced accumulators.
// The closed-form ingredients are passed in (I1v[i] = dI/dM, I2v[i] = d2I/dM2 of flavour i; u = the derivatives of U), so
// that callers may compute them however they like (finish_fj below: serially; the line-march kernel: one flavour per lane).
PNJL_HD void finish_fj_pre(const Model& m, const PointCtx& c, const double x[5], const double acc[kFJAcc], const double I1v[3],
                           const double I2v[3], const UTerms& u, double F[5], double J[25], bool fast) {
    const double T = c.T, invT = c.invT;
    const double twoT = 2.0 * T;
    // dP/dM_i, d2P/dM_i^2, d2P/dM_i dPhi, d2P/dM_i dPhibar  (thermal + vacuum)
    double PM[3], PMM[3], PMP[3], PMPb[3];
    for (int i = 0; i < 3; ++i) {
        const double I1 = I1v[i], I2 = I2v[i];
        const double S1 = -3.0 * invT * c.M[i] * acc[ACC_S1 + i];
        // general path: S2B = sum c n k^2/E^3;  fast path: S2B = sum c n/E^3 and k^2/E^3 = 1/E - M^2/E^3
        const double s2b = fast ? (acc[ACC_S1 + i] - c.M2[i] * acc[ACC_S2B + i]) : acc[ACC_S2B + i];
        const double S2 = 3.0 * invT * invT * c.M2[i] * acc[ACC_S2A + i] - 3.0 * invT * s2b;
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

PNJL_HD void finish_fj(const Model& m, const PointCtx& c, const double x[5], const double acc[kFJAcc], double F[5],
                       double J[25], bool fast = false) {
    double I1v[3], I2v[3];
    double I0 = 0, I1 = 0, I2 = 0;
    for (int i = 0; i < 3; ++i) {
        if (!(i == 1 && c.M[1] == c.M[0])) vacuum_terms(m.Lambda, c.M[i], I0, I1, I2);   // M_d == M_u: reuse
        I1v[i] = I1;
        I2v[i] = I2;
    }
    UTerms u;
    polyakov_U(m, c.T, c.invT, x[3], x[4], u);
    finish_fj_pre(m, c, x, acc, I1v, I2v, u, F, J, fast);
}

// F = grad_x P alone, from the reduced sums of a fused final pass (same formulas as finish_fj).
PNJL_HD void finish_f_pre(const Model& m, const PointCtx& c, const double x[5], const double facc[5], const double I1v[3],
                          const UTerms& u, double F[5]) {
    const double twoT = 2.0 * c.T;
    double PM[3];
    for (int i = 0; i < 3; ++i) PM[i] = twoT * (-3.0 * c.invT * c.M[i] * facc[i]) + 2.0 * m.Nc * I1v[i];
    const double g4 = -4.0 * m.G, k2 = 2.0 * m.K;
    F[0] = PM[0] * g4 + PM[1] * (k2 * x[2]) + PM[2] * (k2 * x[1]) + (-4 * m.G * x[0] + 4 * m.K * x[1] * x[2]);
    F[1] = PM[0] * (k2 * x[2]) + PM[1] * g4 + PM[2] * (k2 * x[0]) + (-4 * m.G * x[1] + 4 * m.K * x[0] * x[2]);
    F[2] = PM[0] * (k2 * x[1]) + PM[1] * (k2 * x[0]) + PM[2] * g4 + (-4 * m.G * x[2] + 4 * m.K * x[0] * x[1]);
    F[3] = twoT * 3.0 * facc[3] - u.U_P;
    F[4] = twoT * 3.0 * facc[4] - u.U_Pb;
}

PNJL_HD void finish_f(const Model& m, const PointCtx& c, const double x[5], const double facc[5], double F[5]) {
    double I1v[3];
    double I0 = 0, I1 = 0, I2 = 0;
    for (int i = 0; i < 3; ++i) {
        if (!(i == 1 && c.M[1] == c.M[0])) vacuum_terms(m.Lambda, c.M[i], I0, I1, I2);
        I1v[i] = I1;
    }
    UTerms u;
    polyakov_U(m, c.T, c.invT, x[3], x[4], u);
    finish_f_pre(m, c, x, facc, I1v, u, F);
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
    const double rE = guarded_rsqrt(E2);
    const double E = E2 * rE;
    const double a = (c.mu - E) * c.invT;
    const double b = -(E + c.mu) * c.invT;
    const Species sp = species_eval(a, c.Phi3, c.Phib3);
    const Species sm = species_eval(b, c.Phib3, c.Phi3);
    const double np = f_fma(c.Phi, sp.r1, f_fma(2.0 * c.Phib, sp.r2, sp.r3));
    const double nm = f_fma(c.Phib, sm.r1, f_fma(2.0 * c.Phi, sm.r2, sm.r3));
    double L = guarded_log(sp.D * sm.D);
    if (sp.pos) L = f_fma(3.0, a, L);
    if (sm.pos) L = f_fma(3.0, b, L);
    acc[TH_NP + FL] = f_fma(coef, np, acc[TH_NP + FL]);
    acc[TH_NM + FL] = f_fma(coef, nm, acc[TH_NM + FL]);
    acc[TH_L] = f_fma(coef, L, acc[TH_L]);
    acc[TH_T] = f_fma(coef, f_fma(np, E - c.mu, nm * (E + c.mu)), acc[TH_T]);
}

// General path of a thermo pass, out of line (see fj_partial_general; same isospin shortcut).
PNJL_HD_NOINL void thermo_partial_general(const PointCtx& c, bool iso, const MeshView& mv, int lane, int stride, double* out) {
    double acc[kThAcc];
#pragma unroll
    for (int i = 0; i < kThAcc; ++i) acc[i] = 0.0;
    if (iso) {
#pragma unroll 1
        for (int k = lane; k < mv.n; k += stride) {
            const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
            const double cf = mv.coef[k];
            thermo_node<0>(c, k2, 2.0 * cf, acc);
            thermo_node<2>(c, k2, cf, acc);
        }
        acc[TH_NP + 0] *= 0.5; acc[TH_NP + 1] = acc[TH_NP + 0];
        acc[TH_NM + 0] *= 0.5; acc[TH_NM + 1] = acc[TH_NM + 0];
    } else {
#pragma unroll 1
        for (int k = lane; k < mv.n; k += stride) {
            const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
            const double cf = mv.coef[k];
            thermo_node<0>(c, k2, cf, acc);
            thermo_node<1>(c, k2, cf, acc);
            thermo_node<2>(c, k2, cf, acc);
        }
    }
#pragma unroll
    for (int i = 0; i < kThAcc; ++i) out[i] = acc[i];
}

// Per-lane partial sums of one thermo pass in the canonical 8-slot layout.
PNJL_HD void thermo_partial(const Model& m, bool isospin, const PointCtx& c, const double x[5], const MeshView& mv_in, int lane,
                            int stride, double acc[kThAcc]) {
    const MeshView mv = select_mesh(mv_in, c.xi);
    const double k2max = mv.p2max + (c.xi > 0.0 ? c.xi * mv.pc2max : 0.0);
    if (fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) {
        FastCtx fc;
        make_fast_ctx(c, fc);
        double t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
        const bool iso = isospin && x[0] == x[1];
        if (iso) {
            double d0 = 0, d1 = 0, dg[2] = {0, 0};
            if (one_log_ok(c.T, c.mu)) {
#pragma unroll 1
                for (int k = lane; k < mv.n; k += stride) {
                    const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                    th_pair_fast<false, true>(fc, c.mu, c.M2[0], c.M2[2], k2, mv.coef[k], t0, t2, d0, d1, dg);
                }
                th_pair_finish(c.mu, t0, t2);
                // t2[2] already holds 2 L_u + L_s; the combination below expects (L_u, L_d, L_s) = (t0, t1, t2)
#pragma unroll
                for (int q = 0; q < 4; ++q) t1[q] = t0[q];
                t0[2] = 0.0; t1[2] = 0.0;
            } else {
#pragma unroll 1
                for (int k = lane; k < mv.n; k += stride) {
                    const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                    th_pair_fast<false, false>(fc, c.mu, c.M2[0], c.M2[2], k2, mv.coef[k], t0, t2, d0, d1, dg);
                }
                th_pair_finish(c.mu, t0, t2);
#pragma unroll
                for (int q = 0; q < 4; ++q) t1[q] = t0[q];
            }
        } else {
            for (int k = lane; k < mv.n; k += stride) {
                const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                const double cf = mv.coef[k];
                thermo_node_fast(fc, c.mu, c.M2[0], k2, cf, t0);
                thermo_node_fast(fc, c.mu, c.M2[1], k2, cf, t1);
                thermo_node_fast(fc, c.mu, c.M2[2], k2, cf, t2);
            }
        }
        acc[TH_NP + 0] = t0[0]; acc[TH_NP + 1] = t1[0]; acc[TH_NP + 2] = t2[0];
        acc[TH_NM + 0] = t0[1]; acc[TH_NM + 1] = t1[1]; acc[TH_NM + 2] = t2[1];
        acc[TH_L] = (t0[2] + t1[2]) + t2[2];
        acc[TH_T] = (t0[3] + t1[3]) + t2[3];
        return;
    }
    thermo_partial_general(c, isospin && x[0] == x[1], mv, lane, stride, acc);
}

// Per-lane partial sums of a fused "final pass": facc[0..2] = S1 per flavour, facc[3] = GP, facc[4] = GPB (flavour
// sums), tacc[8] in the thermo layout.  Returns false when the state is not on the fast path (the caller then
// runs the two ordinary passes).
constexpr int kFtAcc = 5;
PNJL_HD bool ft_partial(const Model& m, bool isospin, const PointCtx& c, const double x[5], const MeshView& mv_in, int lane,
                        int stride, double facc[kFtAcc], double tacc[kThAcc]) {
    const MeshView mv = select_mesh(mv_in, c.xi);
    const double k2max = mv.p2max + (c.xi > 0.0 ? c.xi * mv.pc2max : 0.0);
    if (!fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) return false;
    FastCtx fc;
    make_fast_ctx(c, fc);
    double a0[7] = {0, 0, 0, 0, 0, 0, 0}, a1[7] = {0, 0, 0, 0, 0, 0, 0}, a2[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool iso = isospin && x[0] == x[1];
    if (iso) {
        double tu[4] = {0, 0, 0, 0}, ts[4] = {0, 0, 0, 0}, s1u = 0, s1s = 0, gsh[2] = {0, 0};
        if (one_log_ok(c.T, c.mu)) {
#pragma unroll 1
            for (int k = lane; k < mv.n; k += stride) {
                const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                th_pair_fast<true, true>(fc, c.mu, c.M2[0], c.M2[2], k2, mv.coef[k], tu, ts, s1u, s1s, gsh);
            }
        } else {
#pragma unroll 1
            for (int k = lane; k < mv.n; k += stride) {
                const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
                th_pair_fast<true, false>(fc, c.mu, c.M2[0], c.M2[2], k2, mv.coef[k], tu, ts, s1u, s1s, gsh);
            }
        }
        th_pair_finish(c.mu, tu, ts);
        facc[0] = s1u; facc[1] = s1u; facc[2] = s1s;
        facc[3] = gsh[0]; facc[4] = gsh[1];
        tacc[TH_NP + 0] = tu[0]; tacc[TH_NP + 1] = tu[0]; tacc[TH_NP + 2] = ts[0];
        tacc[TH_NM + 0] = tu[1]; tacc[TH_NM + 1] = tu[1]; tacc[TH_NM + 2] = ts[1];
        tacc[TH_L] = f_fma(2.0, tu[2], ts[2]);
        tacc[TH_T] = f_fma(2.0, tu[3], ts[3]);
        return true;
    } else {
        for (int k = lane; k < mv.n; k += stride) {
            const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
            const double cf = mv.coef[k];
            ft_node_fast(fc, c.mu, c.M2[0], k2, cf, a0);
            ft_node_fast(fc, c.mu, c.M2[1], k2, cf, a1);
            ft_node_fast(fc, c.mu, c.M2[2], k2, cf, a2);
        }
    }
    facc[0] = a0[0]; facc[1] = a1[0]; facc[2] = a2[0];
    facc[3] = (a0[1] + a1[1]) + a2[1];
    facc[4] = (a0[2] + a1[2]) + a2[2];
    tacc[TH_NP + 0] = a0[3]; tacc[TH_NP + 1] = a1[3]; tacc[TH_NP + 2] = a2[3];
    tacc[TH_NM + 0] = a0[4]; tacc[TH_NM + 1] = a1[4]; tacc[TH_NM + 2] = a2[4];
    tacc[TH_L] = (a0[5] + a1[5]) + a2[5];
    tacc[TH_T] = (a0[6] + a1[6]) + a2[6];
    return true;
}

struct Thermo {
    double omega, pressure, rho_norm, entropy, energy;
    double rho[3], nq[3], nqb[3], M[3];
};

PNJL_HD void finish_thermo_pre(const Model& m, const PointCtx& c, const double x[5], const double acc[kThAcc],
                               const double I0v[3], const UTerms& u, Thermo& th) {
    const double T = c.T;
    const double chi = 2 * m.G * ((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]) - 4 * m.K * ((x[0] * x[1]) * x[2]);
    double vac = 0.0;
    for (int i = 0; i < 3; ++i) {
        vac += I0v[i];
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

PNJL_HD void finish_thermo(const Model& m, const PointCtx& c, const double x[5], const double acc[kThAcc], Thermo& th) {
    UTerms u;
    polyakov_U(m, c.T, c.invT, x[3], x[4], u);
    double I0v[3];
    double I0 = 0, I1 = 0, I2 = 0;
    for (int i = 0; i < 3; ++i) {
        if (!(i == 1 && c.M[1] == c.M[0])) vacuum_terms(m.Lambda, c.M[i], I0, I1, I2);
        I0v[i] = I0;
    }
    finish_thermo_pre(m, c, x, acc, I0v, u, th);
}

// ------------------------------------------------------------------------------------------------
// Derivative pass: the partial derivatives in (T, mu) at FIXED x that the implicit differentiation of the gap equations
// needs (ThermoDerivatives.jl:80-109: dx/dtheta = -J^{-1} dF/dtheta; :186-250, :342-467: ds/dtheta, dn/dtheta).  The
// reference gets them from ForwardDiff through calculate_omega; here they are closed-form sums over the same mesh.
// With a = (mu - E)/T, b = -(E + mu)/T:  da/dT = -a/T, db/dT = -b/T, da/dmu = 1/T, db/dmu = -1/T, and per species
// dn/da = Q = q/f - 3 n^2,  d(y/f)/da = r1 (1 - 3n),  d(y^2/f)/da = r2 (2 - 3n)   (same pieces as the Jacobian pass).
//   per flavour i:  A1_i = sum c (a Q+ + b Q-)/E,   B1_i = sum c (Q+ - Q-)/E          -> d(dP/dM_i)/dT, d(dP/dM_i)/dmu
//   shared:         AG  = sum c (a r1+ u+ + b r2- v-),   BG  = sum c (r1+ u+ - r2- v-)  -> dF_Phi/dT,    dF_Phi/dmu
//                   AGB = sum c (a r2+ v+ + b r1- u-),   BGB = sum c (r2+ v+ - r1- u-)  -> dF_Phibar/dT, dF_Phibar/dmu
//                   A2Q = sum c (a^2 Q+ + b^2 Q-)   -> ds/dT,   AQD = sum c (a Q+ - b Q-)   -> ds/dmu = d(sum rho)/dT,
//                   QQ  = sum c (Q+ + Q-)           -> d(sum rho)/dmu
// (shared sums run over the three flavours).  General path only (floors and the a > 0 rescaling live): this pass runs once per
// requested point, not inside the solve.
// ------------------------------------------------------------------------------------------------
constexpr int kDtAcc = 13;
enum { DT_A1 = 0, DT_B1 = 3, DT_AG = 6, DT_BG = 7, DT_AGB = 8, DT_BGB = 9, DT_A2Q = 10, DT_AQD = 11, DT_QQ = 12 };

template <int FL>
PNJL_HD void dtheta_node(const PointCtx& c, double k2, double coef, double acc[kDtAcc]) {
    const double E2 = k2 + c.M2[FL];
    const double rE = guarded_rsqrt(E2);
    const double E = E2 * rE;
    const double a = (c.mu - E) * c.invT;
    const double b = -(E + c.mu) * c.invT;
    const Species sp = species_eval(a, c.Phi3, c.Phib3);
    const Species sm = species_eval(b, c.Phib3, c.Phi3);
    const double np = f_fma(c.Phi, sp.r1, f_fma(2.0 * c.Phib, sp.r2, sp.r3));
    const double nm = f_fma(c.Phib, sm.r1, f_fma(2.0 * c.Phi, sm.r2, sm.r3));
    const double qp = f_fma(c.Phi, sp.r1, f_fma(4.0 * c.Phib, sp.r2, 3.0 * sp.r3));
    const double qm = f_fma(c.Phib, sm.r1, f_fma(4.0 * c.Phi, sm.r2, 3.0 * sm.r3));
    const double Qp = f_fma(-3.0 * np, np, qp), Qm = f_fma(-3.0 * nm, nm, qm);
    const double up = f_fma(-3.0, np, 1.0), vp = f_fma(-3.0, np, 2.0);
    const double um = f_fma(-3.0, nm, 1.0), vm = f_fma(-3.0, nm, 2.0);
    const double crE = coef * rE;
    acc[DT_A1 + FL] = f_fma(crE, a * Qp + b * Qm, acc[DT_A1 + FL]);
    acc[DT_B1 + FL] = f_fma(crE, Qp - Qm, acc[DT_B1 + FL]);
    const double gP = sp.r1 * up, gM = sm.r2 * vm;      // d/da of the Phi-derivative pieces (quark, antiquark)
    const double hP = sp.r2 * vp, hM = sm.r1 * um;      // same for Phibar
    acc[DT_AG] = f_fma(coef, a * gP + b * gM, acc[DT_AG]);
    acc[DT_BG] = f_fma(coef, gP - gM, acc[DT_BG]);
    acc[DT_AGB] = f_fma(coef, a * hP + b * hM, acc[DT_AGB]);
    acc[DT_BGB] = f_fma(coef, hP - hM, acc[DT_BGB]);
    acc[DT_A2Q] = f_fma(coef, (a * a) * Qp + (b * b) * Qm, acc[DT_A2Q]);
    acc[DT_AQD] = f_fma(coef, a * Qp - b * Qm, acc[DT_AQD]);
    acc[DT_QQ] = f_fma(coef, Qp + Qm, acc[DT_QQ]);
}

PNJL_HD void dtheta_partial(const PointCtx& c, const MeshView& mv_in, int lane, int stride, double acc[kDtAcc]) {
    const MeshView mv = select_mesh(mv_in, c.xi);
#pragma unroll
    for (int i = 0; i < kDtAcc; ++i) acc[i] = 0.0;
    for (int k = lane; k < mv.n; k += stride) {
        const double k2 = f_fma(c.xi, mv.pc2[k], mv.p2[k]);
        const double cf = mv.coef[k];
        dtheta_node<0>(c, k2, cf, acc);
        dtheta_node<1>(c, k2, cf, acc);
        dtheta_node<2>(c, k2, cf, acc);
    }
}

// out[16]: dF/dT [5], dF/dmu [5], ds/dT, ds/dmu, dn_B/dT, dn_B/dmu (n_B = sum_i rho_i / 3), all at fixed x; [14], [15] = 0.
// gp / gpb: the sums sum c (r1+ + r2-), sum c (r2+ + r1-) of the Jacobian pass at the same state (ACC_GP, ACC_GPB).
PNJL_HD void finish_dtheta(const Model& m, const PointCtx& c, const double x[5], const double acc[kDtAcc], double gp, double gpb,
                           double out[16]) {
    const double T = c.T, iT = c.invT;
    // d(dP/dM_i)/dtheta = -6 M_i dS1_i/dtheta,  dS1/dT = -A1/T,  dS1/dmu = B1/T
    double PMT[3], PMm[3];
    for (int i = 0; i < 3; ++i) {
        PMT[i] = 6.0 * c.M[i] * iT * acc[DT_A1 + i];
        PMm[i] = -6.0 * c.M[i] * iT * acc[DT_B1 + i];
    }
    const double g4 = -4.0 * m.G, k2 = 2.0 * m.K;
    const double D[3][3] = {{g4, k2 * x[2], k2 * x[1]}, {k2 * x[2], g4, k2 * x[0]}, {k2 * x[1], k2 * x[0], g4}};
    for (int j = 0; j < 3; ++j) {
        out[j] = PMT[0] * D[0][j] + PMT[1] * D[1][j] + PMT[2] * D[2][j];
        out[5 + j] = PMm[0] * D[0][j] + PMm[1] * D[1][j] + PMm[2] * D[2][j];
    }
    // U(T, Phi, Phibar) = T^4 [-1/2 A Phi Phibar + B ln v]: T-derivatives of U_Phi, U_Phibar and the second T-derivative of U
    const double P = x[3], Pb = x[4];
    const double t = m.T0 * iT;
    const double A = m.a0 + m.a1 * t + m.a2 * (t * t);
    const double B = m.b3 * (t * t * t);
    const double iT2 = iT * iT;
    const double dA = -m.a1 * m.T0 * iT2 - 2 * m.a2 * (m.T0 * m.T0) * (iT2 * iT);
    const double dB = -3 * m.b3 * (m.T0 * m.T0 * m.T0) * (iT2 * iT2);
    const double d2A = 2 * m.a1 * m.T0 * (iT2 * iT) + 6 * m.a2 * (m.T0 * m.T0) * (iT2 * iT2);
    const double d2B = 12 * m.b3 * (m.T0 * m.T0 * m.T0) * (iT2 * iT2 * iT);
    const double PPb = Pb * P;
    const double v = 1 - 6 * PPb + 4 * (Pb * Pb * Pb + P * P * P) - 3 * (PPb * PPb);
    const bool live = !(v <= 0.0) && !(v < kPolyakovEps);
    const double lv = live ? log(v) : log(kPolyakovEps);
    const double iv = live ? 1.0 / v : 0.0;
    const double vP = -6 * Pb + 12 * P * P - 6 * P * Pb * Pb;
    const double vPb = -6 * P + 12 * Pb * Pb - 6 * P * P * Pb;
    const double T2 = T * T, T3 = T2 * T, T4 = T2 * T2;
    const double gP0 = -0.5 * A * Pb + B * vP * iv, gPb0 = -0.5 * A * P + B * vPb * iv;         // U_Phi / T^4, U_Phibar / T^4
    const double U_PT = 4 * T3 * gP0 + T4 * (-0.5 * dA * Pb + dB * vP * iv);
    const double U_PbT = 4 * T3 * gPb0 + T4 * (-0.5 * dA * P + dB * vPb * iv);
    const double w0 = -0.5 * A * PPb + B * lv, w1 = -0.5 * dA * PPb + dB * lv, w2 = -0.5 * d2A * PPb + d2B * lv;
    const double U_TT = 12 * T2 * w0 + 8 * T3 * w1 + T4 * w2;
    // F_Phi = 6 T GP - U_Phi:  dGP/dT = -AG/T, dGP/dmu = BG/T
    out[3] = 6.0 * gp - 6.0 * acc[DT_AG] - U_PT;
    out[4] = 6.0 * gpb - 6.0 * acc[DT_AGB] - U_PbT;
    out[8] = 6.0 * acc[DT_BG];
    out[9] = 6.0 * acc[DT_BGB];
    out[10] = -U_TT + 6.0 * iT * acc[DT_A2Q];          // ds/dT at fixed x
    out[11] = -6.0 * iT * acc[DT_AQD];                 // ds/dmu at fixed x
    out[12] = -2.0 * iT * acc[DT_AQD];                 // dn_B/dT = (1/3) d(sum rho)/dT   (Maxwell: = (1/3) ds/dmu)
    out[13] = 2.0 * iT * acc[DT_QQ];                   // dn_B/dmu = (1/3) (6/T) QQ
    out[14] = 0.0;
    out[15] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// 5x5 dense algebra
// ------------------------------------------------------------------------------------------------
// Solve A y = b by LU with partial pivoting.  false on an exactly-zero pivot.  Deliberately compact
// (rolled loops, local arrays): it runs once per quadrature pass and must not crowd the instruction cache.
PNJL_HD_NOINL bool lu_solve5(const double A_in[25], const double b_in[5], double y[5]) {
    double A[25], b[5];
#pragma unroll 1
    for (int i = 0; i < 25; ++i) A[i] = A_in[i];
#pragma unroll 1
    for (int i = 0; i < 5; ++i) b[i] = b_in[i];
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        int piv = k;
        double best = fabs(A[k * 5 + k]);
#pragma unroll 1
        for (int i = k + 1; i < 5; ++i) {
            const double v = fabs(A[i * 5 + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) ok = false;
        if (piv != k) {
#pragma unroll 1
            for (int j = 0; j < 5; ++j) { const double t = A[k * 5 + j]; A[k * 5 + j] = A[piv * 5 + j]; A[piv * 5 + j] = t; }
            const double t = b[k]; b[k] = b[piv]; b[piv] = t;
        }
        const double inv = 1.0 / A[k * 5 + k];
#pragma unroll 1
        for (int i = k + 1; i < 5; ++i) {
            const double l = A[i * 5 + k] * inv;
#pragma unroll 1
            for (int j = k + 1; j < 5; ++j) A[i * 5 + j] -= l * A[k * 5 + j];
            b[i] -= l * b[k];
        }
    }
#pragma unroll 1
    for (int i = 4; i >= 0; --i) {
        double sacc = b[i];
#pragma unroll 1
        for (int j = i + 1; j < 5; ++j) sacc -= A[i * 5 + j] * y[j];
        y[i] = sacc / A[i * 5 + i];
    }
    return ok;
}

// Same elimination order as lu_solve5, fully unrolled with select-based row swaps so that A, b and y live in
// registers (used in the fused quadrature-pass epilogue; A and b are destroyed).
PNJL_HD bool lu_solve5_regs(double A[25], double b[5], double y[5]) {
    bool ok = true;
    double inv[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        int piv = k;
        double best = fabs(A[k * 5 + k]);
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            const double v = fabs(A[i * 5 + k]);
            const bool g = v > best;
            best = g ? v : best;
            piv = g ? i : piv;
        }
        ok = ok && (best != 0.0);
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            const bool sw = (piv == i);
#pragma unroll
            for (int j = k; j < 5; ++j) {
                const double a = A[k * 5 + j], c = A[i * 5 + j];
                A[k * 5 + j] = sw ? c : a;
                A[i * 5 + j] = sw ? a : c;
            }
            const double a = b[k], c = b[i];
            b[k] = sw ? c : a;
            b[i] = sw ? a : c;
        }
        inv[k] = guarded_rcp(A[k * 5 + k]);
#pragma unroll
        for (int i = k + 1; i < 5; ++i) {
            const double l = A[i * 5 + k] * inv[k];
#pragma unroll
            for (int j = k + 1; j < 5; ++j) A[i * 5 + j] = f_fma(-l, A[k * 5 + j], A[i * 5 + j]);
            b[i] = f_fma(-l, b[k], b[i]);
        }
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        double sacc = b[i];
#pragma unroll
        for (int j = i + 1; j < 5; ++j) sacc = f_fma(-A[i * 5 + j], y[j], sacc);
        y[i] = sacc * inv[i];
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

// ------------------------------------------------------------------------------------------------
// One-loop integral A and the effective couplings K_alpha^+- (the per-point step that follows the gap solve in
// scripts/relaxtime/run_gap_transport_scan.jl:297-305 `build_K_data`).
//   A(m, mu, T, Phi, Phibar) = 4 [ -C(m) + sum_i w_i p_i^2 / E_i (f+(E_i) + f-(E_i)) ]     src/relaxtime/OneLoopIntegrals.jl:531-543
//   C(m) = (Lambda sqrt(Lambda^2 + m^2) - m^2 ln((Lambda + sqrt(Lambda^2 + m^2)) / m)) / 2,  m -> max(m, 0); Lambda^2 / 2 below 1e-14   :507-517
//   f+-  = PNJL occupation numbers (isotropic)                                                src/QuarkDistribution.jl:14-51
//   G_f  = -Nc / (4 pi^2) m_f A_f                                                             src/relaxtime/EffectiveCouplings.jl:56-60
//   K_alpha^+-, det K^+-                                                                      EffectiveCouplings.jl:232-279
// The occupation number is written in the scale-free form (divide through by the largest power of y when y > 1), which
// equals the reference's expression wherever that one is finite (its clamp(exp, 1e-200, 1e200) overflows in y^2 beyond
// |a| > 354; this form does not).
// ------------------------------------------------------------------------------------------------
enum { AUX_A_U = 0, AUX_A_S, AUX_G_U, AUX_G_S, AUX_K0_P, AUX_K0_M, AUX_K123_P, AUX_K123_M, AUX_K4567_P, AUX_K4567_M,
       AUX_K8_P, AUX_K8_M, AUX_K08_P, AUX_K08_M, AUX_DETK_P, AUX_DETK_M, kAuxDoubles };

// n = (P1 y + 2 P2 y^2 + y^3) / (1 + 3 P1 y + 3 P2 y^2 + y^3), y = e^a
PNJL_HD double pnjl_occupation(double a, double P1, double P2) {
    const bool tame = (a >= -708.0) && (a <= 708.0);
    const double na = -fabs(a);
    const double w = tame ? fast_exp_nonpos(na) : exp(na);
    double num, den;
    if (a <= 0.0) {
        num = w * (P1 + w * (2.0 * P2 + w));
        den = 1.0 + w * (3.0 * P1 + w * (3.0 * P2 + w));
    } else {
        num = 1.0 + w * (2.0 * P2 + w * P1);
        den = 1.0 + w * (3.0 * P2 + w * (3.0 * P1 + w));
    }
    return num / den;
}

PNJL_HD double oneloop_const_term_A(double Lam, double m) {
    const double mp = m > 0.0 ? m : 0.0;
    if (mp < 1e-14) return 0.5 * (Lam * Lam);
    const double s = sqrt(Lam * Lam + mp * mp);
    return 0.5 * (Lam * s - (mp * mp) * log((Lam + s) / mp));
}

// rule: p2[i] = p_i^2, wp2[i] = w_i p_i^2
PNJL_HD double oneloop_A(double Lam, double m, double mu, double T, double Phi, double Phib, int n, const double* p2,
                         const double* wp2) {
    double acc = -oneloop_const_term_A(Lam, m);
    const double iT = 1.0 / T;
    const double m2 = m * m;
    for (int i = 0; i < n; ++i) {
        const double E = sqrt(p2[i] + m2);
        const double fq = pnjl_occupation(-(E - mu) * iT, Phi, Phib);
        const double fa = pnjl_occupation(-(E + mu) * iT, Phib, Phi);
        acc += wp2[i] / E * (fq + fa);
    }
    return 4.0 * acc;
}

PNJL_HD void effective_couplings(double G, double K, double Nc, double m_u, double m_s, double A_u, double A_s,
                                 double aux[kAuxDoubles]) {
    const double pref = -Nc / (4.0 * (kPi * kPi));
    const double G_u = pref * (m_u * A_u), G_s = pref * (m_s * A_s);
    aux[AUX_A_U] = A_u; aux[AUX_A_S] = A_s; aux[AUX_G_U] = G_u; aux[AUX_G_S] = G_s;
    const double t0 = (1.0 / 3.0) * K * (2.0 * G_u + G_s);
    const double t123 = 0.5 * K * G_s;
    const double t4567 = 0.5 * K * G_u;
    const double t8 = (1.0 / 6.0) * K * (4.0 * G_u - G_s);
    const double t08 = (1.0 / 6.0) * 1.4142135623730951 * K * (G_u - G_s);
    aux[AUX_K0_P] = G - t0;         aux[AUX_K0_M] = G + t0;
    aux[AUX_K123_P] = G + t123;     aux[AUX_K123_M] = G - t123;
    aux[AUX_K4567_P] = G + t4567;   aux[AUX_K4567_M] = G - t4567;
    aux[AUX_K8_P] = G + t8;         aux[AUX_K8_M] = G - t8;
    aux[AUX_K08_P] = t08;           aux[AUX_K08_M] = -t08;
    aux[AUX_DETK_P] = aux[AUX_K0_P] * aux[AUX_K8_P] - aux[AUX_K08_P] * aux[AUX_K08_P];
    aux[AUX_DETK_M] = aux[AUX_K0_M] * aux[AUX_K8_M] - aux[AUX_K08_M] * aux[AUX_K08_M];
}

}  // namespace pnjl
