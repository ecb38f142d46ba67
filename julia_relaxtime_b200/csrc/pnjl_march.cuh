// pnjl_march.cuh — the line-march kernel k_march (included by pnjl_kernels.cu inside namespace pnjl, after pnjl_lean.cuh).
//
// One LEADER warp owns a (xi, mu) line and runs its whole continuity march — seed, Newton iterations, final thermodynamics,
// record (run_gap_transport_scan.jl:407-443, ImplicitSolver.jl:211-328) — alone, or with 1..15 FOLLOWER warps that only sweep
// the mesh for it when a GPU holds fewer lines than warps:
//   * a quadrature pass is the paired FP64 loop of k_solve_ws's workers in one small function (march_sweep_part: lanes
//     stride the mesh in shared memory, warp reduction through shared memory); the sums land in the leader's scratch line
//     and the closed-form finish + the 5x5 elimination run there right away, LANE-PARALLEL (march_finish, pnjl_lean.cuh: one
//     flavour / one matrix entry per lane) — no controller warps, no polling;
//   * the common case — plain Newton from the continuity seed converges to a physical state (ImplicitSolver.jl:103-128) — is
//     a small state machine (march_lean_point) whose state lives in shared memory, so that nothing is spilled around the
//     calls of sweep and finish; what a warp walks per pass has to fit the SM's 32 KB instruction cache next to the other
//     warps' code — every version that shrank it got faster.  Everything else (MultiSeed bootstrap,
//     trust-region fallback, unphysical or floored states) goes through the generic Solver<WarpEval> cascade out of line,
//     which redoes the point from its seed;
//   * lines are time-sliced: a global ticket queue hands out (line, next T index) quanta of `quantum` points; a leader that
//     finishes a quantum parks the line's tracker state (64 bytes) in global memory, re-queues the line and takes the oldest
//     waiting one.  All lines therefore advance at the same pace on ALL SMs (no per-SM imbalance, the tail is one quantum);
//   * teams (few lines per GPU: multi-GPU shares of a fixed grid): the leader posts the pass in the team's command block,
//     leader and followers sweep their node ranges between two named barriers, and the leader adds the partial sums in part
//     order — the result does not depend on who was faster.  Followers never execute finish or state machine;
//   * a record leaves as one coalesced 256-byte row written by the 32 lanes of the leader.
#pragma once

#ifndef PNJL_MARCH_LEAN
#define PNJL_MARCH_LEAN 1
#endif
#ifndef PNJL_MARCH_UNROLL
#define PNJL_MARCH_UNROLL 1
#endif
// Unroll factor of the lean sweep loops.  Two nodes per trip (four independent FP64 chains per warp) was measured SLOWER on
// every share of config 5 (1/8 share 64.8 -> 69.0 ms, full grid 363 -> 429 ms): the loop is 2.3x the code for no extra issue rate.
constexpr int kMarchUnroll = PNJL_MARCH_UNROLL;

constexpr int kMarchWarps = 16;      // warps per CTA (512 threads x 128 registers = the register file of an SM)
constexpr int kBufStride = 24;       // doubles per reduced-sum block: 20 sums + the fast-path flag, padded
constexpr int kStageDoubles = 32;    // per-warp staging line: inputs of a pass (9 doubles) or one record row (32)

struct MarchArgs {
    long long n_lines;
    const double* muq_MeV; const double* xi; const int* table_idx;
    int n_T; const double* T_MeV;
    double* records; const long long* out_index;
    LineState* state;                 // [n_lines] parked tracker state
    int* slots;                       // [capacity] ticket queue: slot t holds the line handed to pop ticket t (-1: not yet pushed)
    unsigned long long* counters;     // [0] pop tickets, [1] push tickets, [2] finished lines
    long long capacity;
    int quantum;                      // points per time slice
};

// Launch constants in constant memory (uploaded by launch_march on the launch stream, like the model constants): the per-warp
// context — which lines of shared memory are mine, how big is my team — is derived from threadIdx and these, so no warp keeps
// a context object in (local) memory and the finish code takes its constants as constant-bank operands.
struct MarchConst {
    const DeviceConfig* cfg;
    int parts;                        // warps per team (1 .. 16); 16 / parts teams per CTA, left-over warps exit at once
    signed char team_of[kMarchWarps]; // team of warp w (-1: no team)
    signed char role_of[kMarchWarps]; // share of a sweep warp w takes in its team: 0 = the leader, 1 .. parts-1 the followers
    int state0;                       // line states [16][32] (doubles)
    int stage0, lean0, team0, cmd0, red0;   // offsets (in doubles) into the dynamic shared memory: staging lines [16][32] | scratch
                                      // lines [16][LW_END] | partial sums [16][24] | command blocks [16][16] | reduction scratch [16][20][33]
    int n, n_iso;
    double p2max, pc2max;
    SolverParams sp;
};
__constant__ MarchConst c_mc;
__constant__ Model c_model;
extern __shared__ double g_smem[];    // namespace scope: device functions address their lines by index, i.e. as SHARED memory

__global__ void k_march_init(MarchArgs a) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < a.capacity) a.slots[i] = i < a.n_lines ? (int)i : -1;
    if (i < a.n_lines) {
        LineState st;
#pragma unroll
        for (int q = 0; q < 5; ++q) st.prev[q] = 0.0;
        st.it_next = 0; st.prev_phase = PH_UNKNOWN; st.has_prev = 0; st.its_hint = 0;
        st.pad[0] = st.pad[1] = st.pad[2] = st.pad[3] = 0;
        a.state[i] = st;
    }
    if (i == 0) { a.counters[0] = 0ULL; a.counters[1] = (unsigned long long)a.n_lines; a.counters[2] = 0ULL; }
}

// ---- per-warp context, recomputed where it is needed ----
__device__ __forceinline__ int mc_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int mc_warp() { return threadIdx.x >> 5; }
// Team and role of a warp: tables in constant memory, filled by the launcher (march_layout) so that the leaders — the warps
// that run the scalar phases — are spread evenly over the four schedulers of the SM (warp w issues on scheduler w & 3).
__device__ __forceinline__ int mc_team() { return c_mc.team_of[mc_warp()]; }
__device__ __forceinline__ int mc_part() { return c_mc.role_of[mc_warp()]; }
__device__ __forceinline__ double* mc_stage() { return g_smem + c_mc.stage0 + mc_warp() * kStageDoubles; }
__device__ __forceinline__ double* mc_W() { return g_smem + c_mc.lean0 + mc_warp() * LW_END; }
__device__ __forceinline__ void mc_mesh(MeshView& mv) {
    const int n = c_mc.n;
    mv.p2 = g_smem; mv.pc2 = g_smem + n; mv.coef = g_smem + 2 * n; mv.n = n;
    mv.p2max = c_mc.p2max; mv.pc2max = c_mc.pc2max;
    mv.p2_iso = g_smem + 3 * n; mv.coef_iso = g_smem + 3 * n + c_mc.n_iso; mv.n_iso = c_mc.n_iso;
}
// The team barrier.  Out of line on purpose: leader and followers then arrive through the SAME bar.sync instruction.  The
// hardware does not care, but compute-sanitizer's synccheck reports (and aborts on) a named barrier that its participants
// reach from different instructions, even in a minimal producer/consumer program (scripts/microbench_named_barrier.cu).
__device__ __noinline__ void mc_team_sync() {
    __syncwarp();
    if (c_mc.parts > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + mc_team()), "r"(32 * c_mc.parts) : "memory");
}

// ---- team protocol ---------------------------------------------------------------------------------------------------
// Warp 0 of a team (the LEADER) runs the line: seeds, Newton logic, closed-form finishes, records.  The other warps (FOLLOWERS)
// only sweep their share of the mesh: they wait at the team's named barrier, read the command the leader left in the team's
// command block — kind of pass and the state (T, mu, xi, x) — sweep, leave their partial sums in the team buffer and meet
// the leader at the barrier again.  What a follower executes is the prologue, the quadrature loop and the reduction (~6 KB);
// the finish and the state machine (~17 KB) are fetched by one warp per team instead of all of them, which is what the
// instruction cache of an SM with 16 desynchronised warps needs.  Two barriers per pass; the partial sums are added by the
// leader in part order, so the result does not depend on which warp is faster.
constexpr int kCmdDoubles = 16;       // op, T, mu, xi, x[5], kapP, kapM
constexpr int MOP_EXIT = -1;          // op: -1 exit, WS_FJ / WS_FT / WS_TH generic sweep of that kind, MOP_LEAN + kind the lean sweep
constexpr int MOP_LEAN = 8;
constexpr int kRedStride = 33;        // row stride of the per-warp reduction scratch [kFJAcc][33]
__device__ __forceinline__ double* mc_cmd() { return g_smem + c_mc.cmd0 + mc_team() * kCmdDoubles; }
__device__ __forceinline__ double* mc_partial(int part) { return g_smem + c_mc.team0 + (mc_team() * c_mc.parts + part) * kBufStride; }

// Leader: publish a command (lane 0 writes, the barrier that follows orders it).
__device__ __forceinline__ void march_post(int op, double T, double mu, double xi, double x0, double x1, double x2, double x3, double x4,
                                           double kapP, double kapM) {
    double* cmd = mc_cmd();
    __syncwarp();
    if (mc_lane() == 0) {
        cmd[1] = T; cmd[2] = mu; cmd[3] = xi;
        cmd[4] = x0; cmd[5] = x1; cmd[6] = x2; cmd[7] = x3; cmd[8] = x4;
        cmd[9] = kapP; cmd[10] = kapM;
        reinterpret_cast<int*>(cmd)[0] = op;
    }
    __syncwarp();
}
// Leader, after the closing barrier of a pass: the team's sums in part order -> W[LW_S ..].
__device__ __forceinline__ void march_collect(bool lean) {
    const int lane = mc_lane();
    double* W = mc_W();
    const double* base = mc_partial(0);
    if (lane < kBufStride) {
        double v = base[lane];
        if (lane != 20)
            for (int q = 1; q < c_mc.parts; ++q) v += base[q * kBufStride + lane];
        W[LW_S + lane] = (lean && lane == 20) ? 1.0 : v;
    }
    __syncwarp();
}

// One quadrature pass of kind `type` at (T, mu, xi, x) through the general ws_worker_pass (any kind of state); afterwards
// the sums of the whole mesh are in the leader's W[LW_S ..].  Leader only.
__device__ __noinline__ void march_pass(int type, double T, double mu, double xi, double x0, double x1, double x2, double x3, double x4) {
    march_post(type, T, mu, xi, x0, x1, x2, x3, x4, 0.0, 0.0);
    MeshView mv;
    mc_mesh(mv);
    if (c_mc.parts == 1) {
        ws_worker_pass(c_mc.cfg, mv, mc_cmd() + 1, type, mc_lane(), 0, 1, mc_W() + LW_S);
        __syncwarp();
        return;
    }
    mc_team_sync();
    ws_worker_pass(c_mc.cfg, mv, mc_cmd() + 1, type, mc_lane(), 0, c_mc.parts, mc_partial(0));
    mc_team_sync();
    march_collect(false);
}

// ---- the hot sweep ---------------------------------------------------------------------------------------------------
// This warp's share of a Jacobian pass (WS_FJ) or a fused final pass (WS_FT) for the state every continuity point is in: on
// the integrand's fast path, phi_u == phi_d bitwise, |mu|/T <= 60 (one logarithm per node) — the leader checks.  Same inline
// loops as ws_worker_pass (fj_pair_fast / th_pair_fast<true, true>), but ONE small function for both kinds, the per-point
// constants e^{+-mu/T} supplied by the caller, and the warp reduction through shared memory (20 stores, a rolled loop of
// 32 loads and adds per sum: ~50 instructions instead of the ~300 of the shuffle network).  Sums go to out[0 .. 19].
__device__ __noinline__ void march_sweep_part(int type, double T, double mu, double xi, double x0, double x1, double x2, double x3, double x4,
                                              double kapP, double kapM, double* out) {
    const int lane = mc_lane();
    const double x[5] = {x0, x1, x2, x3, x4};
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    FastCtx fc;
    fc.nInvT = -c.invT;
    fc.nInvT_l2e = -c.invT * 1.4426950408889634;
    fc.kapP = kapP; fc.kapM = kapM;
    fc.Phi = c.Phi; fc.Phib = c.Phib;
    fc.Phi3 = c.Phi3; fc.Phib3 = c.Phib3;
    fc.Phi2 = 2.0 * c.Phi; fc.Phib2 = 2.0 * c.Phib;
    fc.Phi4 = 4.0 * c.Phi; fc.Phib4 = 4.0 * c.Phib;
    const int n = c_mc.n;
    const bool iso = xi == 0.0 && c_mc.n_iso > 0;
    const double* p2 = iso ? g_smem + 3 * n : g_smem;
    const double* pc2 = iso ? p2 : g_smem + n;
    const double* coef = iso ? g_smem + 3 * n + c_mc.n_iso : g_smem + 2 * n;
    const int nn = iso ? c_mc.n_iso : n;
    const int stride = 32 * c_mc.parts;
    double* R = g_smem + c_mc.red0 + mc_warp() * (kFJAcc * kRedStride) + lane;
    if (type == WS_FJ) {
        double fu[5] = {0, 0, 0, 0, 0}, fs[5] = {0, 0, 0, 0, 0}, sh[5] = {0, 0, 0, 0, 0};
#pragma unroll kMarchUnroll
        for (int k = lane + 32 * mc_part(); k < nn; k += stride) {
            const double k2 = f_fma(xi, pc2[k], p2[k]);
            fj_pair_fast(fc, c.M2[0], c.M2[2], k2, coef[k], fu, fs, sh);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            R[(3 * q + 0) * kRedStride] = fu[q];
            R[(3 * q + 1) * kRedStride] = fu[q];
            R[(3 * q + 2) * kRedStride] = fs[q];
            R[(15 + q) * kRedStride] = sh[q];
        }
    } else {
        double tu[4] = {0, 0, 0, 0}, ts[4] = {0, 0, 0, 0}, s1u = 0, s1s = 0, gsh[2] = {0, 0};
#pragma unroll kMarchUnroll
        for (int k = lane + 32 * mc_part(); k < nn; k += stride) {
            const double k2 = f_fma(xi, pc2[k], p2[k]);
            th_pair_fast<true, true>(fc, mu, c.M2[0], c.M2[2], k2, coef[k], tu, ts, s1u, s1s, gsh);
        }
        th_pair_finish(mu, tu, ts);
        R[0 * kRedStride] = s1u; R[1 * kRedStride] = s1u; R[2 * kRedStride] = s1s;
        R[3 * kRedStride] = gsh[0]; R[4 * kRedStride] = gsh[1];
        R[(kFtAcc + TH_NP + 0) * kRedStride] = tu[0]; R[(kFtAcc + TH_NP + 1) * kRedStride] = tu[0]; R[(kFtAcc + TH_NP + 2) * kRedStride] = ts[0];
        R[(kFtAcc + TH_NM + 0) * kRedStride] = tu[1]; R[(kFtAcc + TH_NM + 1) * kRedStride] = tu[1]; R[(kFtAcc + TH_NM + 2) * kRedStride] = ts[1];
        R[(kFtAcc + TH_L) * kRedStride] = f_fma(2.0, tu[2], ts[2]);
        R[(kFtAcc + TH_T) * kRedStride] = f_fma(2.0, tu[3], ts[3]);
    }
    __syncwarp();
    const int n_sums = type == WS_FJ ? kFJAcc : kFtAcc + kThAcc;
    if (lane < kFJAcc) {
        const double* col = R - lane + lane * kRedStride;         // row `lane` of this warp's scratch
        double v0 = 0.0, v1 = 0.0;
        if (lane < n_sums) {
            v0 = col[0]; v1 = col[1];
#pragma unroll 5
            for (int j = 2; j < 32; j += 2) { v0 += col[j]; v1 += col[j + 1]; }       // ten loads in flight per trip
        }
        out[lane] = v0 + v1;
    }
    __syncwarp();
}

// Leader: one lean pass; afterwards the sums of the whole mesh are in W[LW_S ..] (and the fast-path flag W[LW_S + 20] = 1).
template <bool TEAMS>
__device__ __forceinline__ void march_sweep(int type, double T, double mu, double xi, double x0, double x1, double x2, double x3, double x4,
                                            double kapP, double kapM) {
    if (!TEAMS || c_mc.parts == 1) {
        double* W = mc_W();
        march_sweep_part(type, T, mu, xi, x0, x1, x2, x3, x4, kapP, kapM, W + LW_S);
        if (mc_lane() == 0) W[LW_S + 20] = 1.0;
        __syncwarp();
        return;
    }
    march_post(MOP_LEAN + type, T, mu, xi, x0, x1, x2, x3, x4, kapP, kapM);
    mc_team_sync();
    march_sweep_part(type, T, mu, xi, x0, x1, x2, x3, x4, kapP, kapM, mc_partial(0));
    mc_team_sync();
    march_collect(true);
}

// Follower: serve the leader's passes until it says exit.
__device__ __noinline__ void march_follow() {
    const int lane = mc_lane(), part = mc_part();
    const double* cmd = mc_cmd();
    double* mine = mc_partial(part);
    MeshView mv;
    mc_mesh(mv);
    for (;;) {
        mc_team_sync();
        const int op = reinterpret_cast<const int*>(cmd)[0];
        if (op < 0) break;
        if (op >= MOP_LEAN) march_sweep_part(op - MOP_LEAN, cmd[1], cmd[2], cmd[3], cmd[4], cmd[5], cmd[6], cmd[7], cmd[8], cmd[9], cmd[10], mine);
        else ws_worker_pass(c_mc.cfg, mv, cmd + 1, op, lane, part, c_mc.parts, mine);
        mc_team_sync();
    }
}
// Leader, when it has no more work: release the followers.
__device__ __forceinline__ void march_dismiss() {
    if (c_mc.parts == 1) return;
    march_post(MOP_EXIT, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    mc_team_sync();
}

// ---- the unified lean finish -----------------------------------------------------------------------------------------
// One function (one copy of every closed form) for the three kinds of pass whose sums are in W[LW_S ..]:
//   kind 0 (Jacobian pass)  F -> W[LW_F ..], Newton direction p = -J^{-1} F -> W[LW_P ..]; returns 1, or 0 if a pivot was zero
//   kind 1 (fused final)    F -> W[LW_F ..] and the thermodynamic functions straight into the record staging line
//   kind 2 (thermo pass)    the thermodynamic functions only
// Lane-parallel (pnjl_lean.cuh): lane f < 3 flavour f, lanes 3..11 the dM/dphi table, lane e < 30 one entry of [J | F], one
// elimination step per pivot.  kinds 1, 2 return 1 when every thermodynamic function is finite (part of the physicality test).
// Returns -1 (uniformly) when a closed form is not tame: the caller then takes the redundant cold versions.
__device__ __noinline__ int march_finish(int kind, double T, double mu, double xi, double x0, double x1, double x2, double x3, double x4) {
    const int lane = mc_lane();
    double* W = mc_W();
    const double x[5] = {x0, x1, x2, x3, x4};
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    LeanConst k;
    lean_consts(c_model, c.T, c.invT, k);
    // phase A
    bool tame = polyakov_tame(c.Phi, c.Phib);
    if (lane < 3) {
        const double Mf = lane == 0 ? c.M[0] : (lane == 1 ? c.M[1] : c.M[2]);
        const double M2f = lane == 0 ? c.M2[0] : (lane == 1 ? c.M2[1] : c.M2[2]);
        if (vacuum_tame(k.Lambda, Mf)) {
            double I0, I1, I2;
            vacuum_terms_t<true>(k.Lambda, Mf, I0, I1, I2);
            if (kind == 0) lean_flavour_fj_pre(lane, k, Mf, M2f, W[LW_S + 20] != 0.0, I1, I2, W);
            else if (kind == 1) W[LW_PM + lane] = k.twoT * (-3.0 * k.invT * Mf * W[LW_S + lane]) + k.Nc2 * I1;
            W[LW_I0 + lane] = I0;
        } else {
            tame = false;
        }
    } else if (lane < 12) {
        const int r = (lane - 3) / 3, j = (lane - 3) - 3 * r;
        if (kind != 2) lean_dtable(r, j, k, x, W);
    } else if (lane == 12) {
        if (kind != 2) {
#pragma unroll
            for (int q = 0; q < 5; ++q) W[LW_X + q] = x[q];
        }
    }
    tame = __all_sync(0xffffffffu, tame);
    if (!tame) return -1;
    UTerms u;
    if (kind == 0) polyakov_eval<true, false>(c_model, c.T, c.invT, c.Phi, c.Phib, u);
    else polyakov_eval<true, true>(c_model, c.T, c.invT, c.Phi, c.Phib, u);
    if (kind != 2 && lane == 13) {
        W[LW_U + 0] = u.U_P; W[LW_U + 1] = u.U_Pb; W[LW_U + 2] = u.U_PP; W[LW_U + 3] = u.U_PPb; W[LW_U + 4] = u.U_PbPb;
    }
    __syncwarp();
    int rc = 1;
    if (kind != 2) {
        // phase B: kind 0 all 30 entries of [J | F], kind 1 the right-hand side only (same code, other sum offsets)
        const int li = lane / 6, lc = lane - 6 * li;
        double a = 0.0;
        if (lane < 30 && (kind == 0 || lc == 5)) {
            a = lean_aug_entry(li, lc, k, W, kind == 0 ? ACC_GP : 3, kind == 0 ? ACC_GPB : 4);
            W[LW_AUG + lane] = a;
            if (lc == 5) W[LW_F + li] = a;
        }
        __syncwarp();
        if (kind == 0) {
            // phase C
            bool ok = true;
#pragma unroll 1
            for (int step = 0; step < 5; ++step) {
                double inv = 0.0, nxt = a;
                if (lane < 30) nxt = lean_lu_step(step, li, lc, W, a, inv, ok);
                __syncwarp();
                if (lane < 30) { a = nxt; W[LW_AUG + lane] = a; }
                if (lane == 0) W[LW_INV + step] = inv;
                __syncwarp();
            }
            ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
            double y[5];
            lean_backsub(W, y);
            if (lane < 5) {
                const double yl = lane == 0 ? y[0] : (lane == 1 ? y[1] : (lane == 2 ? y[2] : (lane == 3 ? y[3] : y[4])));
                W[LW_P + lane] = -yl;
            }
            __syncwarp();
            return ok ? 1 : 0;
        }
    }
    // thermodynamic functions (kinds 1, 2): finish_thermo_pre, written into the record staging line at their record offsets
    {
        double tacc[kThAcc], I0v[3];
        const int off = kind == 1 ? kFtAcc : 0;
#pragma unroll
        for (int i = 0; i < kThAcc; ++i) tacc[i] = W[LW_S + off + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) I0v[i] = W[LW_I0 + i];
        Thermo th;
        finish_thermo_pre(c_model, c, x, tacc, I0v, u, th);
        bool fin = finite_d(th.omega) && finite_d(th.pressure) && finite_d(th.rho_norm) && finite_d(th.entropy) && finite_d(th.energy);
#pragma unroll
        for (int i = 0; i < 3; ++i) fin = fin && finite_d(th.M[i]) && th.M[i] > 0.0;
        rc = fin ? 1 : 0;
        double* stage = mc_stage();
        if (lane == 0) {
            stage[PNJL_REC_OMEGA] = th.omega; stage[PNJL_REC_PRESSURE] = th.pressure; stage[PNJL_REC_RHO_NORM] = th.rho_norm;
            stage[PNJL_REC_ENTROPY] = th.entropy; stage[PNJL_REC_ENERGY] = th.energy;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                stage[PNJL_REC_MASS + i] = th.M[i]; stage[PNJL_REC_NQ + i] = th.nq[i];
                stage[PNJL_REC_NQBAR + i] = th.nqb[i]; stage[PNJL_REC_RHO + i] = th.rho[i];
            }
        }
        __syncwarp();
    }
    return rc;
}

// Redundant-per-lane finishes (finish_fj / finish_f / finish_thermo with their libm fall-backs) for states that are not tame.
__device__ __noinline__ bool cold_fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
    const double* S = mc_W() + LW_S;
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    double acc[kFJAcc], J[25], b[5];
#pragma unroll
    for (int i = 0; i < kFJAcc; ++i) acc[i] = S[i];
    finish_fj(c_model, c, x, acc, F, J, S[20] != 0.0);
#pragma unroll
    for (int i = 0; i < 5; ++i) b[i] = F[i];
    const bool ok = lu_solve5_regs(J, b, p);
#pragma unroll
    for (int i = 0; i < 5; ++i) p[i] = -p[i];
    return ok;
}
__device__ __noinline__ void cold_fj(double T, double mu, double xi, const double x[5], double F[5], double J[25]) {
    const double* S = mc_W() + LW_S;
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    double acc[kFJAcc];
#pragma unroll
    for (int i = 0; i < kFJAcc; ++i) acc[i] = S[i];
    finish_fj(c_model, c, x, acc, F, J, S[20] != 0.0);
}
__device__ __noinline__ void cold_f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
    const double* S = mc_W() + LW_S;
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    double facc[kFtAcc], tacc[kThAcc];
#pragma unroll
    for (int i = 0; i < kFtAcc; ++i) facc[i] = S[i];
#pragma unroll
    for (int i = 0; i < kThAcc; ++i) tacc[i] = S[kFtAcc + i];
    finish_f(c_model, c, x, facc, F);
    finish_thermo(c_model, c, x, tacc, th);
}
__device__ __noinline__ void cold_thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
    const double* S = mc_W() + LW_S;
    PointCtx c;
    make_ctx(c_model, T, mu, xi, x, c);
    double tacc[kThAcc];
#pragma unroll
    for (int i = 0; i < kThAcc; ++i) tacc[i] = S[i];
    finish_thermo(c_model, c, x, tacc, th);
}

// The thermodynamic functions march_finish left in the record staging line, as a Thermo (generic cascade only).
__device__ __forceinline__ void thermo_from_stage(Thermo& th) {
    const double* stage = mc_stage();
    th.omega = stage[PNJL_REC_OMEGA]; th.pressure = stage[PNJL_REC_PRESSURE]; th.rho_norm = stage[PNJL_REC_RHO_NORM];
    th.entropy = stage[PNJL_REC_ENTROPY]; th.energy = stage[PNJL_REC_ENERGY];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        th.M[i] = stage[PNJL_REC_MASS + i]; th.nq[i] = stage[PNJL_REC_NQ + i];
        th.nqb[i] = stage[PNJL_REC_NQBAR + i]; th.rho[i] = stage[PNJL_REC_RHO + i];
    }
}

// Evaluation policy of the generic cascade (Solver<WarpEval>: MultiSeed bootstrap, trust-region fallback, everything the
// in-kernel fast path hands back).  Stateless: the context comes from threadIdx and constant memory.  Sweeps go through the
// general ws_worker_pass (every kind of state), finishes through march_finish with the cold versions as fall-back.
// A pass of the generic cascade: through the lean sweep (march_sweep, the small loop that is hot in the instruction cache)
// when the state qualifies for it — integrand fast path, phi_u == phi_d bitwise, |mu| <= 60 T: true for nearly every iterate
// of a bootstrap at T >= 50 MeV — and through the general ws_worker_pass otherwise.  The sums land in W[LW_S ..] in the same
// layout either way (fast-path flag at W[LW_S + 20]).  Before this, the 3 % of the passes of a full config-5 grid that belong
// to the cascade (first point of every line, phase flips) took 38 % of the kernel's time: ws_worker_pass is 100 KB of code
// that every warp walked alone, at 86 cycles per instruction (profiles/r02c_march_cfg5_full.md).
__device__ __noinline__ void march_pass_any(int type, double T, double mu, double xi, const double x[5]) {
    if (type != WS_TH && c_mc.sp.isospin && x[0] == x[1] && one_log_ok(T, mu) && T > 1e-300 && T < 1e300) {
        double M[3];
        masses_of(c_model, x, M);
        const double M2[3] = {M[0] * M[0], M[1] * M[1], M[2] * M[2]};
        const double k2max = c_mc.p2max + (xi > 0.0 ? xi * c_mc.pc2max : 0.0);
        if (fast_path_ok(T, mu, x[3], x[4], k2max, M2)) {
            const double km = fast_exp_nonpos(-fabs(mu) * fast_rcp(T));
            const double kp = fast_rcp(km);
            march_sweep<true>(type, T, mu, xi, x[0], x[1], x[2], x[3], x[4], mu >= 0.0 ? kp : km, mu >= 0.0 ? km : kp);
            return;
        }
    }
    march_pass(type, T, mu, xi, x[0], x[1], x[2], x[3], x[4]);
}

struct WarpEval {
    __device__ __noinline__ void fj(double T, double mu, double xi, const double x[5], double F[5], double J[25]) {
        march_pass_any(WS_FJ, T, mu, xi, x);
        cold_fj(T, mu, xi, x, F, J);                     // only the trust-region method wants J itself
    }
    __device__ __noinline__ bool fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
        march_pass_any(WS_FJ, T, mu, xi, x);
        const int rc = march_finish(0, T, mu, xi, x[0], x[1], x[2], x[3], x[4]);
        if (rc < 0) return cold_fj_step(T, mu, xi, x, F, p);
        const double* W = mc_W();
#pragma unroll
        for (int i = 0; i < 5; ++i) { F[i] = W[LW_F + i]; p[i] = W[LW_P + i]; }
        return rc != 0;
    }
    __device__ __noinline__ bool f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
        PointCtx c;
        make_ctx(c_model, T, mu, xi, x, c);
        const double k2max = c_mc.p2max + (xi > 0.0 ? xi * c_mc.pc2max : 0.0);
        if (!fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) return false;   // uniform over the team
        march_pass_any(WS_FT, T, mu, xi, x);
        if (march_finish(1, T, mu, xi, x[0], x[1], x[2], x[3], x[4]) < 0) { cold_f_thermo(T, mu, xi, x, F, th); return true; }
        const double* W = mc_W();
#pragma unroll
        for (int i = 0; i < 5; ++i) F[i] = W[LW_F + i];
        thermo_from_stage(th);
        return true;
    }
    __device__ __noinline__ void thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
        march_pass(WS_TH, T, mu, xi, x[0], x[1], x[2], x[3], x[4]);
        if (march_finish(2, T, mu, xi, x[0], x[1], x[2], x[3], x[4]) < 0) { cold_thermo(T, mu, xi, x, th); return; }
        thermo_from_stage(th);
    }
};

// Record writer of a team: warp 0 of the team stages the row in its shared-memory line and the 32 lanes store it.
__device__ __forceinline__ void march_store_row(const double rec[PNJL_REC_DOUBLES], double* row) {
    const int lane = mc_lane();
    double* stage = mc_stage();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < PNJL_REC_DOUBLES; q += 2) *reinterpret_cast<double2*>(stage + q) = make_double2(rec[q], rec[q + 1]);
    }
    __syncwarp();
    row[lane] = stage[lane];
    __syncwarp();
}
struct MarchSink {
    double* base;
    double xi;
    __device__ __forceinline__ void operator()(int it, const PointRes& r, double T_fm, double mu_fm, int n_fj, int n_th,
                                               int n_ft) {
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T_fm, mu_fm, xi, n_fj, n_th, n_ft, rec);
        march_store_row(rec, base + (long long)PNJL_REC_DOUBLES * it);
    }
};

// Anything that is not the common case: one point through the generic cascade, exactly as scan_line_slice does it.
__device__ __noinline__ void march_generic_point(const PhaseTables* pt, int ti, double muq_MeV, double xi, int n_T, const double* T_MeV,
                                                 LineState& st, double* rows) {
    WarpEval ev;
    Solver<WarpEval> sv(c_model, c_mc.sp, ev);
    MarchSink sink{rows, xi};
    scan_line_slice(sv, pt, ti, muq_MeV, xi, n_T, T_MeV, st, 1, sink);
}

// ---- the line of a leader: its state lives in shared memory --------------------------------------------------------------
// Everything a line carries from pass to pass — the Newton iterate, the tracker of PhaseAwareContinuitySeed, counters, flags —
// sits in a per-warp block of shared memory and every step below loads what it needs and stores what it changed.  Kept in
// registers, that state is live across the calls of the sweep and the finish and ptxas spills it to local memory around
// each of them (~100 LDL/STL per pass); with 512 threads x 4 KB of stack per SM those slots miss in L1 and main() waited on
// them for 15 % of a config-4 run (profiles/r02b_march_cfg4.md: long scoreboard).
enum { DS_X = 0, DS_XOLD = 5, DS_T = 10, DS_MU = 11, DS_XI = 12, DS_KAPP = 13, DS_KAPM = 14, DS_K2MAX = 15, DS_RES = 16, DS_TM = 17,
       DS_MUQ = 18, DS_PREV = 19, DS_ROWS = 24, DS_INTS = 25 };
enum { DJ_LINE = 0, DJ_IT = 1, DJ_ITEND = 2, DJ_KIND = 3, DJ_ITERS = 4, DJ_NFJ = 5, DJ_NFT = 6, DJ_HINT = 7, DJ_FLAGS = 8,
       DJ_PREVPH = 9, DJ_HASPREV = 10, DJ_ITSHINT = 11, DJ_TI = 12 };
enum { DF_FIRST = 1, DF_REFRESH = 2, DF_HAVETH = 4, DF_THFINITE = 8, DF_NONSING = 16, DF_XC = 32, DF_FC = 64, DF_FLIP = 128, DF_FINALTH = 256 };
constexpr int kStateDoubles = 32;     // 25 doubles + 13 ints
__device__ __forceinline__ double* mc_state() { return g_smem + c_mc.state0 + mc_warp() * kStateDoubles; }

// Generic cascade for the current point of this warp's line (bootstrap, fall-backs, anything the lean path hands back).
__device__ __noinline__ void march_generic_step(const MarchArgs& a) {
    double* S = mc_state();
    int* J = reinterpret_cast<int*>(S + DS_INTS);
    LineState st;
#pragma unroll
    for (int q = 0; q < 5; ++q) st.prev[q] = S[DS_PREV + q];
    st.it_next = J[DJ_IT]; st.prev_phase = J[DJ_PREVPH]; st.has_prev = J[DJ_HASPREV]; st.its_hint = J[DJ_ITSHINT];
    double* rows = *reinterpret_cast<double**>(S + DS_ROWS);
    march_generic_point(&c_mc.cfg->pt, J[DJ_TI], S[DS_MUQ], S[DS_XI], a.n_T, a.T_MeV, st, rows);
    __syncwarp();
    if (mc_lane() == 0) {
#pragma unroll
        for (int q = 0; q < 5; ++q) S[DS_PREV + q] = st.prev[q];
        J[DJ_IT] = st.it_next; J[DJ_PREVPH] = st.prev_phase; J[DJ_HASPREV] = st.has_prev; J[DJ_ITSHINT] = st.its_hint;
    }
    __syncwarp();
}

// One point of the line on the lean path: NLsolve newton_ (Solver::newton), common case only — every state on the integrand's
// fast path with phi_u == phi_d and |mu| <= 60 T, every closed form tame, F finite.  One pass site: kind = Jacobian pass or
// fused final pass (predicted), a mispredicted final pass is followed by a Jacobian pass at the same x ("refresh").  Returns
// true when the point was solved and its record written; anything else -> false, the generic cascade redoes the point from
// its seed.
template <bool TEAMS>
__device__ __forceinline__ bool march_lean_point(const MarchArgs& a) {
    const int lane = mc_lane();
    const SolverParams& sp = c_mc.sp;
    double* const S = mc_state();
    int* const J = reinterpret_cast<int*>(S + DS_INTS);
    {
        // PhaseAwareContinuitySeed get_seed (SeedStrategies.jl:795-839) with a previous solution
        const double Tm = a.T_MeV[J[DJ_IT]];
        const double T = Tm / c_model.hbarc;
        const double mu_fm = S[DS_MU];
        if (!one_log_ok(T, mu_fm) || !(T > 1e-300 && T < 1e300)) return false;
        const int cur = current_phase(&c_mc.cfg->pt, J[DJ_TI], T * 197.327, mu_fm * 197.327);
        const int pp = J[DJ_PREVPH];
        const bool flip = (pp == PH_HADRON && cur == PH_QUARK) || (pp == PH_QUARK && cur == PH_HADRON);
        double x[5];
        if (flip) seed_const(cur == PH_HADRON ? 0 : 1, x);
        else {
#pragma unroll
            for (int q = 0; q < 5; ++q) x[q] = S[DS_PREV + q];
        }
        // e^{+-mu/T}: per-point constants of the sweeps (make_fast_ctx's arithmetic)
        const double km = fast_exp_nonpos(-fabs(mu_fm) * fast_rcp(T));
        const double kp = fast_rcp(km);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) S[DS_X + q] = x[q];
            S[DS_T] = T; S[DS_TM] = Tm;
            S[DS_KAPP] = mu_fm >= 0.0 ? kp : km;
            S[DS_KAPM] = mu_fm >= 0.0 ? km : kp;
            S[DS_RES] = 0.0;
            J[DJ_HINT] = flip ? 0 : J[DJ_ITSHINT];
            J[DJ_KIND] = WS_FJ; J[DJ_ITERS] = 0; J[DJ_NFJ] = 0; J[DJ_NFT] = 0;
            J[DJ_FLAGS] = DF_FIRST | DF_NONSING | (flip ? DF_FLIP : 0);
        }
        __syncwarp();
    }
    for (;;) {
        {
            // pre-sweep test, then the sweep
            double x[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) x[q] = S[DS_X + q];
            const double T = S[DS_T], mu_fm = S[DS_MU];
            double M[3];
            masses_of(c_model, x, M);
            const double M2[3] = {M[0] * M[0], M[1] * M[1], M[2] * M[2]};
            if (!(x[0] == x[1]) || !fast_path_ok(T, mu_fm, x[3], x[4], S[DS_K2MAX], M2)) return false;
            march_sweep<TEAMS>(J[DJ_KIND], T, mu_fm, S[DS_XI], x[0], x[1], x[2], x[3], x[4], S[DS_KAPP], S[DS_KAPM]);
        }
        int rc;
        {
            rc = march_finish(J[DJ_KIND] == WS_FJ ? 0 : 1, S[DS_T], S[DS_MU], S[DS_XI], S[DS_X], S[DS_X + 1], S[DS_X + 2], S[DS_X + 3], S[DS_X + 4]);
            if (rc < 0) return false;
        }
        // ---- Newton logic on what the finish left in the scratch line ----
        double x[5], xold[5], F[5], p[5];
        const double* W = mc_W();
#pragma unroll
        for (int q = 0; q < 5; ++q) { x[q] = S[DS_X + q]; xold[q] = S[DS_XOLD + q]; F[q] = W[LW_F + q]; p[q] = W[LW_P + q]; }
        int kind = J[DJ_KIND], iters = J[DJ_ITERS], n_fj = J[DJ_NFJ], n_ft = J[DJ_NFT], flags = J[DJ_FLAGS];
        double res = S[DS_RES];
        if (kind == WS_FJ) {
            flags = rc != 0 ? (flags | DF_NONSING) : (flags & ~DF_NONSING);
            ++n_fj;
        } else {
            flags = rc != 0 ? (flags | DF_THFINITE) : (flags & ~DF_THFINITE);
            ++n_ft;
        }
        if (!all_finite5(F)) return false;
        bool leave = false, again = false;
        if (flags & DF_FINALTH) {
            // the fused pass that follows a solve which ended on a Jacobian pass (below): the thermodynamic functions of the
            // converged x; the convergence tests were made on that Jacobian pass and its residual is the one reported
            flags |= DF_HAVETH;
            leave = true;
        } else if (flags & DF_REFRESH) {
            flags &= ~(DF_REFRESH | DF_HAVETH);       // J(x) is known now; the convergence tests of this x were made on the fused pass
        } else {
            if (flags & DF_FIRST) {
                flags &= ~DF_FIRST;
            } else {
                double dx = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) dx = fmax(dx, fabs(x[i] - xold[i]));
                flags = dx <= sp.xtol ? (flags | DF_XC) : (flags & ~DF_XC);
                flags = kind == WS_FT ? (flags | DF_HAVETH) : (flags & ~DF_HAVETH);
            }
            res = norm_inf5(F);
            flags = res <= sp.ftol ? (flags | DF_FC) : (flags & ~DF_FC);
            if ((flags & (DF_XC | DF_FC)) || iters >= sp.max_iter) leave = true;
            else if (kind == WS_FT) { kind = WS_FJ; flags |= DF_REFRESH; again = true; }
        }
        if (!leave && !again) {
            ++iters;
            if (!(flags & DF_NONSING)) return false;
            p[1] = p[0];                                      // keep the exact u<->d symmetry of the equations (x[0] == x[1] here)
            double pmax = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) { xold[i] = x[i]; x[i] = x[i] + p[i]; pmax = fmax(pmax, fabs(p[i])); }
            const int hint = J[DJ_HINT];
            const bool by_history = hint > 0 && iters <= hint;
            const bool predict = sp.predict_tol > 0.0 && (pmax <= sp.xtol || (by_history ? iters == hint : res <= sp.predict_tol));
            kind = predict ? WS_FT : WS_FJ;
        }
        if (leave && (flags & DF_FC) && !(flags & DF_HAVETH) && kind == WS_FJ) {
            // f-converged on a Jacobian pass — the prediction expected one more iteration (0.5 % of the points of config 5):
            // NLsolve is done; what is missing are the thermodynamic functions of this x.  One fused pass at the same x
            // provides them (handing the point to the generic cascade instead re-solved it from its seed through ~60 KB of
            // cold code).
            flags |= DF_FINALTH;
            kind = WS_FT;
            leave = false;
        }
        if (!leave) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 5; ++q) { S[DS_X + q] = x[q]; S[DS_XOLD + q] = xold[q]; }
                S[DS_RES] = res;
                J[DJ_KIND] = kind; J[DJ_ITERS] = iters; J[DJ_NFJ] = n_fj; J[DJ_NFT] = n_ft; J[DJ_FLAGS] = flags;
            }
            __syncwarp();
            continue;
        }
        // _nlsolve_with_tr_fallback (ImplicitSolver.jl:103-151): the trust-region fallback runs unless the primary solve is
        // f-converged with a finite residual <= residual_norm_max and a physical state -> anything else: generic cascade.
        // A final pass that was not a fused one (x-converged on a Jacobian pass: rare) also goes there.
        const double rfin = (flags & DF_FINALTH) ? res : norm_inf5(F);
        if (!((flags & DF_FC) && (flags & DF_HAVETH) && (flags & DF_THFINITE) && finite_d(rfin) && rfin <= sp.residual_norm_max &&
              finite_d(x[3]) && finite_d(x[4]) && (-sp.phi_tol <= x[3] && x[3] <= 1 + sp.phi_tol) &&
              (-sp.phi_tol <= x[4] && x[4] <= 1 + sp.phi_tol)))
            return false;
        // the record: march_finish left the thermodynamic functions and the masses in the staging line; the rest here
        double* stage = mc_stage();
        double* rows = *reinterpret_cast<double**>(S + DS_ROWS);
        const int it = J[DJ_IT];
        const int ph_now = current_phase(&c_mc.cfg->pt, J[DJ_TI], S[DS_TM], S[DS_MUQ]);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) stage[PNJL_REC_X + i] = x[i];
            const int status = PNJL_ST_CONVERGED | ((flags & DF_FLIP) ? PNJL_ST_PHASE_SWITCH : 0) |
                               (stage[PNJL_REC_MASS + 2] <= stage[PNJL_REC_MASS] ? PNJL_ST_MASS_INVERSION : 0);
            stage[PNJL_REC_RESNORM] = rfin;
            stage[PNJL_REC_ITER] = (double)iters;
            stage[PNJL_REC_STATUS] = (double)status;
            stage[PNJL_REC_NEVAL] = (double)n_fj;
            stage[PNJL_REC_NTHERMO] = 0.0;
            stage[PNJL_REC_T] = S[DS_T]; stage[PNJL_REC_MU] = S[DS_MU]; stage[PNJL_REC_XI] = S[DS_XI];
            stage[PNJL_REC_NFUSED] = (double)n_ft;
            stage[31] = 0.0;
            // tracker update! (SeedStrategies.jl:851-856) and the history for the next point
#pragma unroll
            for (int q = 0; q < 5; ++q) S[DS_PREV + q] = x[q];
            J[DJ_PREVPH] = ph_now; J[DJ_ITSHINT] = iters; J[DJ_IT] = it + 1;
        }
        __syncwarp();
        rows[(long long)PNJL_REC_DOUBLES * it + lane] = stage[lane];
        __syncwarp();
        return true;
    }
}

// Dynamic shared memory: mesh [3 n + 2 n_iso] | staging lines [16][32] | scratch lines [16][LW_END] | partial sums [16][24] |
// command blocks [16][16] | reduction scratch [16][20][33] | line states [16][32]
// TEAMS = false: the instantiation for one warp per line (launch_march picks it when parts == 1): no team code in the hot loop —
// what 16 leaders per SM fetch per pass is what bounds that case (config 4: every pass is two trips of the quadrature loop).
template <bool TEAMS>
__global__ void __launch_bounds__(32 * kMarchWarps, 1) k_march(const double* __restrict__ g_mesh, MarchArgs a) {
    const int n_mesh = 3 * c_mc.n + 2 * c_mc.n_iso;
    for (int i = threadIdx.x; i < n_mesh; i += blockDim.x) g_smem[i] = g_mesh[i];
    __syncthreads();
    if (TEAMS && mc_team() < 0) return;
    if (TEAMS && mc_part() != 0) { march_follow(); return; }
    const int lane = mc_lane();
    double* const S = mc_state();
    int* const J = reinterpret_cast<int*>(S + DS_INTS);
    for (;;) {
        // ---- pop: take a ticket and wait for the line that goes with it (or for the end of the scan) ----
        int line = -1;
        if (lane == 0) {
            volatile int* slots = a.slots;
            volatile unsigned long long* done = a.counters + 2;
            const unsigned long long t = atomicAdd(a.counters + 0, 1ULL);
            for (;;) {
                if ((long long)t < a.capacity) {
                    line = slots[t];
                    if (line >= 0) break;
                }
                if (*done >= (unsigned long long)a.n_lines) { line = -1; break; }
                __nanosleep(500);
            }
            __threadfence();
        }
        line = __shfl_sync(0xffffffffu, line, 0);
        if (line < 0) break;
        // ---- this line's parameters and parked tracker state ----
        if (lane == 0) {
            const LineState* g = a.state + line;
#pragma unroll
            for (int q = 0; q < 5; ++q) S[DS_PREV + q] = __ldcg(&g->prev[q]);
            const int it0 = __ldcg(&g->it_next);
            J[DJ_IT] = it0; J[DJ_PREVPH] = __ldcg(&g->prev_phase);
            J[DJ_HASPREV] = __ldcg(&g->has_prev); J[DJ_ITSHINT] = __ldcg(&g->its_hint);
            J[DJ_ITEND] = (it0 + a.quantum < a.n_T) ? it0 + a.quantum : a.n_T;
            const long long row = a.out_index ? a.out_index[line] : (long long)line;
            *reinterpret_cast<double**>(S + DS_ROWS) = a.records + (long long)PNJL_REC_DOUBLES * a.n_T * row;
            const double muq = a.muq_MeV[line], xi = a.xi[line];
            S[DS_MUQ] = muq; S[DS_XI] = xi; S[DS_MU] = muq / c_model.hbarc;
            S[DS_K2MAX] = c_mc.p2max + (xi > 0.0 ? xi * c_mc.pc2max : 0.0);
            J[DJ_TI] = a.table_idx ? a.table_idx[line] : -1;
            J[DJ_LINE] = line;
        }
        __syncwarp();
        // ---- one time slice ----
        while (J[DJ_IT] < J[DJ_ITEND]) {
            bool solved = false;
#if PNJL_MARCH_LEAN
            if (J[DJ_HASPREV] && c_mc.sp.isospin) solved = march_lean_point<TEAMS>(a);
#endif
            if (!solved) march_generic_step(a);
        }
        // ---- park the line (or retire it) ----
        if (lane == 0) {
            const int it = J[DJ_IT];
            if (it < a.n_T) {
                LineState* g = a.state + line;
#pragma unroll
                for (int q = 0; q < 5; ++q) __stcg(&g->prev[q], S[DS_PREV + q]);
                __stcg(&g->it_next, it); __stcg(&g->prev_phase, J[DJ_PREVPH]);
                __stcg(&g->has_prev, J[DJ_HASPREV]); __stcg(&g->its_hint, J[DJ_ITSHINT]);
                __threadfence();
                const unsigned long long t = atomicAdd(a.counters + 1, 1ULL);
                if ((long long)t < a.capacity) ((volatile int*)a.slots)[t] = line;
            } else {
                __threadfence();
                atomicAdd(a.counters + 2, 1ULL);
            }
        }
        __syncwarp();
    }
    if (TEAMS) march_dismiss();
}

// ---- independent points, one warp (team) per point ----------------------------------------------------------------
// PNJL.solve / solve_multi at n independent (T, mu, xi) triples (ImplicitSolver.jl:211-328, :532-559): a team pulls a point
// from a global counter and runs the whole cascade — up to six seeds, Newton, trust-region fallback — in its own warp(s)
// through Solver<WarpEval>.  Nothing waits for a controller: with MultiSeed at every point (config 3) the passes of the
// warp-specialised kernel starve behind its two controller warps' scalar cascade; here 16 warps per SM run 16 cascades.
struct MarchPointArgs {
    long long n;
    const double* T_fm; const double* mu_fm; const double* xi;
    int seed_mode; int n_seeds; const double* seeds;
    double* records;
    unsigned long long* counter;
};

__global__ void __launch_bounds__(32 * kMarchWarps, 1) k_march_points(const double* __restrict__ g_mesh, MarchPointArgs a) {
    const int n_mesh = 3 * c_mc.n + 2 * c_mc.n_iso;
    for (int i = threadIdx.x; i < n_mesh; i += blockDim.x) g_smem[i] = g_mesh[i];
    __syncthreads();
    if (mc_team() < 0) return;
    if (mc_part() != 0) { march_follow(); return; }
    const int lane = mc_lane();
    WarpEval ev;
    Solver<WarpEval> sv(c_model, c_mc.sp, ev);
    for (;;) {
        long long i = 0;
        if (lane == 0) i = (long long)atomicAdd(a.counter, 1ULL);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= a.n) break;
        const double T = a.T_fm[i], mu = a.mu_fm[i], x_i = a.xi[i];
        sv.set_point(T, mu, x_i);
        sv.n_fj = 0; sv.n_th = 0; sv.n_ft = 0;
        sv.its_hint = 0;
        PointRes r;
        if (a.seed_mode == PNJL_SEED_EXPLICIT && a.n_seeds == 1) {
            double x0[5];
            copy5(x0, a.seeds + 5 * i);
            sv.solve_with_fallback(x0, r);
        } else if (a.seed_mode == PNJL_SEED_EXPLICIT) {
            sv.solve_multi(a.seeds + 5 * (long long)a.n_seeds * i, a.n_seeds, r);
        } else if (a.seed_mode == PNJL_SEED_AUTO) {
            double x0[5];
            default_seed(2, T, mu, x0);
            sv.solve_with_fallback(x0, r);
        } else {
            sv.solve_multi(nullptr, 6, r);
        }
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T, mu, x_i, sv.n_fj, sv.n_th, sv.n_ft, rec);
        march_store_row(rec, a.records + (long long)PNJL_REC_DOUBLES * i);
    }
    march_dismiss();
}
