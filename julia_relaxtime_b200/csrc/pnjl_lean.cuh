// pnjl_lean.cuh — the closed-form finish of a quadrature pass and the 5x5 elimination, written LANE-PARALLEL over one warp.
//
// Why: in the line-march kernel every warp finishes its own passes.  Done redundantly in every lane (finish_fj +
// lu_solve5_regs) that is ~1500 straight-line instructions (25 KB) per pass with ~490 FP64-pipe instructions; 16 desynchronised
// warps then overflow the SM's 32 KB instruction cache and the FP64 pipe loses ~9 % to replicated scalar work.  Here the 32
// lanes share the work through a small scratch line in shared memory:
//   phase A   lane f < 3 evaluates flavour f (vacuum integral derivatives, dP/dM, d2P/dM2, d2P/dM dPhi, d2P/dM dPhibar), lanes
//             3..11 the 3x3 table dM_k/dphi_j, all lanes the Polyakov-potential derivatives (no logarithm in a Jacobian pass);
//   phase B   lane e < 30 assembles ONE entry of the augmented matrix [J | F] (5 x 6);
//   phase C   Gaussian elimination with partial pivoting, one matrix entry per lane, one step per pivot (a rolled loop), same
//             pivot rule and the same fma sequence per entry as lu_solve5_regs, so the Newton direction is bit-identical to the
//             redundant version for the same [J | F]; back substitution redundantly in every lane (15 fma).
// Every function is a pure function of (lane, scratch) — `PNJL_HD`, so tests/hostsim can run the 32 lanes one after the other on
// the CPU and compare with finish_fj / lu_solve5_regs (tests/test_host_logic.py).  On the device a __syncwarp() separates the
// phases.
#pragma once

#include "pnjl_math.cuh"

namespace pnjl {

// Scratch line of one warp (doubles).  LW_S: the reduced sums of the pass in the layout the pass wrote them
// (FJ: kFJAcc sums + fast flag at 20; fused: 5 F sums + 8 thermo sums).
enum { LW_S = 0, LW_PM = 24, LW_PMM = 27, LW_PMP = 30, LW_PMPB = 33, LW_D = 36, LW_X = 45, LW_U = 50, LW_I0 = 56, LW_INV = 59,
       LW_AUG = 64, LW_F = 96, LW_P = 101, LW_END = 112 };
// LW_F: F[5] of the pass, LW_P: the Newton direction p[5] (results of the unified finish of the line-march kernel)
// LW_U: U_P, U_Pb, U_PP, U_PPb, U_PbPb, (pad)     LW_I0: I(Lambda, M_f) per flavour (fused / thermo passes)

struct LeanConst {           // per-pass uniform scalars every phase needs
    double T, invT, twoT;
    double g4, k2, K4;       // -4G, 2K, 4K
    double Nc2;              // 2 Nc
    double Lambda;
};

PNJL_HD void lean_consts(const Model& m, double T, double invT, LeanConst& k) {
    k.T = T; k.invT = invT; k.twoT = 2.0 * T;
    k.g4 = -4.0 * m.G; k.k2 = 2.0 * m.K; k.K4 = 4.0 * m.K;
    k.Nc2 = 2.0 * m.Nc;
    k.Lambda = m.Lambda;
}

// Phase A, flavour lanes (lane < 3 stores; other lanes may call it with f = 2, their result is dropped by the caller).
// FJ pass: PM, PMM, PMP, PMPb of flavour f from the sums S (finish_fj_pre's per-flavour block, same expressions).
// `fast`: the S2B' convention of the fast path.  Returns false if the vacuum term of this flavour is not tame (caller bails).
PNJL_HD bool lean_flavour_fj(int f, const LeanConst& k, double Mf, double M2f, bool fast, double* W) {
    if (!vacuum_tame(k.Lambda, Mf)) return false;
    double I0, I1, I2;
    vacuum_terms_t<true>(k.Lambda, Mf, I0, I1, I2);
    const double* S = W + LW_S;
    const double a1 = S[ACC_S1 + f], a2a = S[ACC_S2A + f], a2b = S[ACC_S2B + f], a3 = S[ACC_S3 + f], a4 = S[ACC_S4 + f];
    const double S1 = -3.0 * k.invT * Mf * a1;
    const double s2b = fast ? (a1 - M2f * a2b) : a2b;
    const double S2 = 3.0 * k.invT * k.invT * M2f * a2a - 3.0 * k.invT * s2b;
    const double S3 = -3.0 * k.invT * Mf * a3;
    const double S4 = -3.0 * k.invT * Mf * a4;
    W[LW_PM + f] = k.twoT * S1 + k.Nc2 * I1;
    W[LW_PMM + f] = k.twoT * S2 + k.Nc2 * I2;
    W[LW_PMP + f] = k.twoT * S3;
    W[LW_PMPB + f] = k.twoT * S4;
    return true;
}
// Same with the vacuum derivatives I1 = dI/dM, I2 = d2I/dM2 supplied by the caller (the warp-specialised kernel: the
// controller lane computes the closed forms of x while the worker sweeps the mesh, and ships them in the mailbox).
PNJL_HD void lean_flavour_fj_pre(int f, const LeanConst& k, double Mf, double M2f, bool fast, double I1, double I2, double* W) {
    const double* S = W + LW_S;
    const double a1 = S[ACC_S1 + f], a2a = S[ACC_S2A + f], a2b = S[ACC_S2B + f], a3 = S[ACC_S3 + f], a4 = S[ACC_S4 + f];
    const double S1 = -3.0 * k.invT * Mf * a1;
    const double s2b = fast ? (a1 - M2f * a2b) : a2b;
    const double S2 = 3.0 * k.invT * k.invT * M2f * a2a - 3.0 * k.invT * s2b;
    const double S3 = -3.0 * k.invT * Mf * a3;
    const double S4 = -3.0 * k.invT * Mf * a4;
    W[LW_PM + f] = k.twoT * S1 + k.Nc2 * I1;
    W[LW_PMM + f] = k.twoT * S2 + k.Nc2 * I2;
    W[LW_PMP + f] = k.twoT * S3;
    W[LW_PMPB + f] = k.twoT * S4;
}
// Fused final pass: PM of flavour f from the F sums (finish_f_pre) and the vacuum integral itself for the thermo finish.
PNJL_HD bool lean_flavour_ft(int f, const LeanConst& k, double Mf, double* W) {
    if (!vacuum_tame(k.Lambda, Mf)) return false;
    double I0, I1, I2;
    vacuum_terms_t<true>(k.Lambda, Mf, I0, I1, I2);
    W[LW_PM + f] = k.twoT * (-3.0 * k.invT * Mf * W[LW_S + f]) + k.Nc2 * I1;
    W[LW_I0 + f] = I0;
    return true;
}
// Thermo-only pass: just the vacuum integral.
PNJL_HD bool lean_flavour_th(int f, const LeanConst& k, double Mf, double* W) {
    if (!vacuum_tame(k.Lambda, Mf)) return false;
    double I0, I1, I2;
    vacuum_terms_t<true>(k.Lambda, Mf, I0, I1, I2);
    W[LW_I0 + f] = I0;
    return true;
}

// Phase A, table lanes: entry (r, j) of D[r][j] = dM_r / dphi_j, row-major.
PNJL_HD void lean_dtable(int r, int j, const LeanConst& k, const double x[5], double* W) {
    const int o = 3 - r - j;
    const double xo = o == 0 ? x[0] : (o == 1 ? x[1] : x[2]);
    W[LW_D + 3 * r + j] = (r == j) ? k.g4 : k.k2 * xo;
}

// Phase B: entry (i, c) of the augmented matrix [J | F] (lane e = 6 i + c): c < 5 -> J[i][c], c == 5 -> F[i].
// gp_off / gpb_off: where the pass left sum c (r1+ + r2-) and sum c (r2+ + r1-) (FJ: ACC_GP, ACC_GPB; fused: 3, 4).
PNJL_HD double lean_aug_entry(int i, int c, const LeanConst& k, const double* W, int gp_off, int gpb_off) {
    const double* S = W + LW_S;
    if (i >= 3 && c >= 3) {
        if (c == 5) return k.twoT * 3.0 * S[i == 3 ? gp_off : gpb_off] - W[LW_U + (i - 3)];          // F[3], F[4]
        const int q = (i - 3) + (c - 3);                        // (3,3) -> 0, (3,4) and (4,3) -> 1, (4,4) -> 2
        const int hs = q == 0 ? ACC_HPP : (q == 1 ? ACC_HPPB : ACC_HPBPB);
        return -9.0 * k.twoT * S[hs] - W[LW_U + 2 + q];         // U_PP, U_PPb, U_PbPb
    }
    const int j1 = i < 3 ? i : c;                               // the condensate index the mass chain D(., j1) runs over
    const bool both = i < 3 && c < 3;
    const int aoff = both ? LW_PMM : (c == 5 ? LW_PM : ((i < 3 ? c : i) == 3 ? LW_PMP : LW_PMPB));
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double t = W[aoff + q] * W[LW_D + 3 * q + j1];
        if (both) t *= W[LW_D + 3 * q + c];
        v = q == 0 ? t : v + t;
    }
    if (both) {
        if (i == c) v += k.g4;
        else {
            const int o = 3 - i - c;
            v += k.k2 * W[LW_PM + o] + k.K4 * W[LW_X + o];
        }
    } else if (c == 5) {
        const int ja = i == 0 ? 1 : 0, jb = i == 2 ? 1 : 2;     // -chi: d/dphi_i = -4G phi_i + 4K phi_ja phi_jb
        v += k.g4 * W[LW_X + i] + k.K4 * W[LW_X + ja] * W[LW_X + jb];
    }
    return v;
}

// Phase C, elimination step `step` for the lane that owns entry (i, c) (value `a`, also stored at W[LW_AUG + 6 i + c]).  Reads only;
// the caller stores the returned value (and lane 0 the reciprocal pivot) after every lane has read.
PNJL_HD double lean_lu_step(int step, int i, int c, const double* W, double a, double& inv_piv, bool& ok) {
    const double* A = W + LW_AUG;
    int piv = step;
    double best = fabs(A[6 * step + step]);
    for (int r = step + 1; r < 5; ++r) {
        const double v = fabs(A[6 * r + step]);
        if (v > best) { best = v; piv = r; }
    }
    ok = ok && (best != 0.0);
    inv_piv = guarded_rcp(A[6 * piv + step]);
    const int src = (i == step) ? piv : ((i == piv) ? step : i);          // my row after rows `step` and `piv` are swapped
    double mine = (src == i) ? a : A[6 * src + c];
    if (i > step && c > step) {
        const double l = A[6 * src + step] * inv_piv;
        mine = f_fma(-l, A[6 * piv + c], mine);
    }
    return mine;
}

// Back substitution (every lane, redundantly): y = U^{-1} b with the reciprocal pivots at W[LW_INV ..].
PNJL_HD void lean_backsub(const double* W, double y[5]) {
    const double* A = W + LW_AUG;
#pragma unroll
    for (int i = 4; i >= 0; --i) {
        double s = A[6 * i + 5];
#pragma unroll
        for (int j = i + 1; j < 5; ++j) s = f_fma(-A[6 * i + j], y[j], s);
        y[i] = s * W[LW_INV + i];
    }
}

}  // namespace pnjl
