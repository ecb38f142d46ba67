// pnjl_kernels.cu — sm_100a kernels and the C ABI (include/pnjl_b200.h) of libpnjl_b200.so.
//
// Kernel layout (one group of G lanes per solve; G = 8, 16 or 32 chosen from the mesh size):
//   * the quadrature mesh (p^2, (p cos)^2, coefficient per node; 24 B/node) is staged once per CTA in
//     shared memory; lanes of a group stride it, every node is reused for the three flavours
//     (three independent dependency chains per lane);
//   * each lane keeps 20 (FJ pass) or 8 (thermo pass) FP64 partial sums in registers; a group-wide
//     xor-butterfly of __shfl_xor_sync leaves the bit-identical totals in every lane;
//   * every lane of the group then runs the 5x5 Newton / dogleg / seed-cascade logic redundantly on
//     identical registers (pnjl_solver.cuh), so control flow is group-uniform and no broadcast is needed;
//   * groups pull work (points or whole continuity lines) from a global atomic counter, so slow points
//     (fallback cascades) do not stall a static partition;
//   * results leave as 256-byte records written cooperatively by the lanes of the group.
// There is deliberately no tensor-core / TMA use: the path is FP64 transcendental-heavy quadrature with
// ~7e3 FLOP per byte of HBM traffic, bounded by the FP64 pipe.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "pnjl_solver.cuh"
#include "pnjl_lean.cuh"

namespace pnjl {

struct DeviceConfig {
    Model m;
    SolverParams sp;
    int n_nodes;
    int n_iso;              // p_num when the isotropic collapse is enabled (mesh buffer then carries 2 n_iso more doubles), else 0
    int lockstep;
    double p2max, pc2max;   // mesh maxima (bound on E for the fast-path test)
    unsigned long long* dbg;   // phase-profiling counters (PNJL_PROFILE_PHASES builds only)
    PhaseTables pt;
};

// ---- group-collective evaluation policy ----------------------------------------------------------
template <int G>
struct GroupEval {
    const Model* m;
    MeshView mv;     // mesh in shared memory
    int isospin;
    int lane;        // lane within the group
    unsigned mask;   // lanes of this group within the warp
    int lockstep;    // CTA-wide phase alignment (G == 32 only), see bar_enter()
    int* s_done;     // shared counter of warps that ran out of work
#ifdef PNJL_PROFILE_PHASES
    unsigned long long* dbg;
    long long t_last, t0, t1;
#endif

    // Phase alignment.  The warps of a CTA alternate between the quadrature loops (a few KB of straight-line FP64
    // code) and ~40 KB of scalar Newton / finishing code.  Left alone they drift apart, the SM's instruction cache
    // then has to hold everything at once, misses on a third of its requests (ncu: sm__icc_requests_lookup_miss)
    // and the FP64 pipe starves.  A named barrier in front of every quadrature loop keeps the warps of the CTA in
    // the same phase, so the loop is fetched once and stays resident while it runs.
    //   barrier 1 ("enter"): bar.sync by everybody;
    //   barrier 2 ("leave"): bar.arrive (non-blocking, lockstep == 1) or bar.sync (lockstep == 3) by working warps.
    // Warps that ran out of work keep both barriers going from drain() until every warp of the CTA is there, so
    // the arrival count is always the full CTA and nothing relies on how exited warps are counted.
    __device__ __forceinline__ bool aligned() const { return G == 32 && lockstep != 0; }
    __device__ __forceinline__ void bar_enter() const {
        asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    }
    __device__ __forceinline__ void bar_leave(bool blocking) const {
        if (blocking) asm volatile("bar.sync 2, %0;" ::"r"((int)blockDim.x) : "memory");
        else asm volatile("bar.arrive 2, %0;" ::"r"((int)blockDim.x) : "memory");
    }

    __device__ __forceinline__ void pass_begin() {
#ifdef PNJL_PROFILE_PHASES
        t0 = clock64();
        if (aligned()) bar_enter();
        t1 = clock64();
#else
        if (aligned()) bar_enter();
#endif
    }
    __device__ __forceinline__ void pass_end() {
#ifdef PNJL_PROFILE_PHASES
        const long long t2 = clock64();
        if (aligned()) bar_leave(lockstep == 3);
        const long long t3 = clock64();
        if (lane == 0 && dbg) {
            atomicAdd(dbg + 0, (unsigned long long)(t2 - t1));                 // loop cycles
            atomicAdd(dbg + 1, (unsigned long long)((t1 - t0) + (t3 - t2)));   // barrier waits
            atomicAdd(dbg + 2, 1ULL);                                          // passes
            if (t_last) atomicAdd(dbg + 3, (unsigned long long)(t0 - t_last)); // scalar phase since the last pass
        }
        t_last = t3;
#else
        if (aligned()) bar_leave(lockstep == 3);
#endif
    }

    // Out of work: count this warp in (between the two barriers, so that the value every drained warp reads after
    // barrier 2 is the same) and keep the barriers going until all warps of the CTA have drained.
    __device__ __forceinline__ void drain() const {
        if (!aligned()) return;
        const int n_warps = blockDim.x >> 5;
        bool counted = false;
        for (;;) {
            bar_enter();
            if (!counted && (threadIdx.x & 31) == 0) atomicAdd(s_done, 1);
            counted = true;
            bar_leave(true);
            if (*((volatile int*)s_done) >= n_warps) break;
        }
    }

    // Group-wide sum, bit-identical in every lane.  For G == 32 the mask is a compile-time constant, which
    // keeps each step at SHFL.BFLY x2 + DADD (a runtime mask costs ~15 instructions per shuffle).
    __device__ __forceinline__ double gsum(double v) const {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(G == 32 ? 0xffffffffu : mask, v, off);
        return v;
    }

    __device__ __noinline__ void fj(double T, double mu, double xi, const double x[5], double F[5], double J[25], double* gp2 = nullptr) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kFJAcc];
        pass_begin();
        const bool fast = fj_partial(*m, isospin != 0, c, x, mv, lane, G, acc);
        pass_end();
#pragma unroll
        for (int i = 0; i < kFJAcc; ++i) acc[i] = gsum(acc[i]);
        finish_fj(*m, c, x, acc, F, J, fast);
        if (gp2) { gp2[0] = acc[ACC_GP]; gp2[1] = acc[ACC_GPB]; }
    }

    // Derivative pass (ThermoDerivatives.jl:80-109, :186-250): dF/dT, dF/dmu, ds/dtheta, dn_B/dtheta at fixed x -> out[16].
    __device__ __noinline__ void dtheta(double T, double mu, double xi, const double x[5], double gp, double gpb, double out[16]) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kDtAcc];
        pass_begin();
        dtheta_partial(c, mv, lane, G, acc);
        pass_end();
#pragma unroll
        for (int i = 0; i < kDtAcc; ++i) acc[i] = gsum(acc[i]);
        finish_dtheta(*m, c, x, acc, gp, gpb, out);
    }

    // Fused pass: F(x) and the Newton direction p = -J^{-1} F, with J and the 5x5 elimination in registers.
    __device__ __noinline__ bool fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kFJAcc];
        pass_begin();
        const bool fast = fj_partial(*m, isospin != 0, c, x, mv, lane, G, acc);
        pass_end();
#pragma unroll
        for (int i = 0; i < kFJAcc; ++i) acc[i] = gsum(acc[i]);
        double J[25], b[5];
        finish_fj(*m, c, x, acc, F, J, fast);
#pragma unroll
        for (int i = 0; i < 5; ++i) b[i] = F[i];
        const bool ok = lu_solve5_regs(J, b, p);
#pragma unroll
        for (int i = 0; i < 5; ++i) p[i] = -p[i];
        return ok;
    }

    // Fused final pass: F(x) and the thermo pass at x from one sweep over the mesh (fast path only; returns false
    // without doing anything otherwise).
    __device__ __noinline__ bool f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        {
            const double k2max = mv.p2max + (c.xi > 0.0 ? c.xi * mv.pc2max : 0.0);
            if (!fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) return false;   // group-uniform
        }
        double facc[kFtAcc], tacc[kThAcc];
        pass_begin();
        ft_partial(*m, isospin != 0, c, x, mv, lane, G, facc, tacc);
        pass_end();
#pragma unroll
        for (int i = 0; i < kFtAcc; ++i) facc[i] = gsum(facc[i]);
#pragma unroll
        for (int i = 0; i < kThAcc; ++i) tacc[i] = gsum(tacc[i]);
        finish_f(*m, c, x, facc, F);
        finish_thermo(*m, c, x, tacc, th);
        return true;
    }

    __device__ __noinline__ void thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kThAcc];
        pass_begin();
        thermo_partial(*m, isospin != 0, c, x, mv, lane, G, acc);
        pass_end();
#pragma unroll
        for (int i = 0; i < kThAcc; ++i) acc[i] = gsum(acc[i]);
        finish_thermo(*m, c, x, acc, th);
    }
};

template <int G>
__device__ __forceinline__ void stage_mesh(const double* __restrict__ g_mesh, int n /* doubles */, double* s_mesh, int* s_done) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_mesh[i] = g_mesh[i];
    if (threadIdx.x == 0) *s_done = 0;
    __syncthreads();
}

template <int G>
__device__ __forceinline__ GroupEval<G> make_eval(const DeviceConfig* cfg, const double* s_mesh, int* s_done) {
    GroupEval<G> ev;
    const int n = cfg->n_nodes;
    ev.m = &cfg->m;
    ev.mv.p2 = s_mesh;
    ev.mv.pc2 = s_mesh + n;
    ev.mv.coef = s_mesh + 2 * n;
    ev.mv.n = n;
    ev.mv.p2max = cfg->p2max;
    ev.mv.pc2max = cfg->pc2max;
    ev.mv.p2_iso = s_mesh + 3 * n;
    ev.mv.coef_iso = s_mesh + 3 * n + cfg->n_iso;
    ev.mv.n_iso = cfg->n_iso;
    ev.isospin = cfg->sp.isospin;
    ev.lockstep = cfg->lockstep;
    ev.s_done = s_done;
#ifdef PNJL_PROFILE_PHASES
    ev.dbg = cfg->dbg;
    ev.t_last = 0;
#endif
    const int wl = threadIdx.x & 31;
    ev.lane = wl % G;
    ev.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl - ev.lane));
    return ev;
}

// Next work item for this group (leader pulls, group receives by shuffle).
template <int G>
__device__ __forceinline__ long long next_task(unsigned long long* counter, const GroupEval<G>& ev) {
    unsigned long long t = 0;
    if (ev.lane == 0) t = atomicAdd(counter, 1ULL);
    const int leader = (threadIdx.x & 31) - ev.lane;
    t = __shfl_sync(ev.mask, t, leader);
    return (long long)t;
}

template <int G>
__device__ __forceinline__ void store_record(const GroupEval<G>& ev, const double rec[PNJL_REC_DOUBLES], double* __restrict__ out) {
#pragma unroll
    for (int j = 0; j < PNJL_REC_DOUBLES / G; ++j) {
        const int idx = j * G + ev.lane;
        double v = rec[0];
#pragma unroll
        for (int q = 1; q < PNJL_REC_DOUBLES; ++q) v = (q == idx) ? rec[q] : v;
        out[idx] = v;
    }
}

// ---- independent points --------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(512, 1) k_solve_points(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_mesh,
                                                      long long n_points, const double* __restrict__ T_fm,
                                                      const double* __restrict__ mu_fm, const double* __restrict__ xi,
                                                      int seed_mode, int n_seeds, const double* __restrict__ seeds,
                                                      double* __restrict__ records, unsigned long long* counter) {
    extern __shared__ double s_mesh[];
    __shared__ int s_done;
    stage_mesh<G>(g_mesh, 3 * cfg->n_nodes + 2 * cfg->n_iso, s_mesh, &s_done);
    GroupEval<G> ev = make_eval<G>(cfg, s_mesh, &s_done);
    Solver<GroupEval<G>> sv(cfg->m, cfg->sp, ev);
    for (;;) {
        const long long i = next_task<G>(counter, ev);
        if (i >= n_points) break;
        const double T = T_fm[i], mu = mu_fm[i], x_i = xi[i];
        sv.set_point(T, mu, x_i);
        sv.n_fj = 0;
        sv.n_th = 0;
        sv.n_ft = 0;
        PointRes r;
        if (seed_mode == PNJL_SEED_EXPLICIT && n_seeds == 1) {
            double x0[5];
            copy5(x0, seeds + 5 * i);
            sv.solve_with_fallback(x0, r);
        } else if (seed_mode == PNJL_SEED_EXPLICIT) {
            sv.solve_multi(seeds + 5 * (long long)n_seeds * i, n_seeds, r);
        } else if (seed_mode == PNJL_SEED_AUTO) {
            double x0[5];
            default_seed(2, T, mu, x0);
            sv.solve_with_fallback(x0, r);
        } else {
            sv.solve_multi(nullptr, 6, r);
        }
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T, mu, x_i, sv.n_fj, sv.n_th, sv.n_ft, rec);
        store_record<G>(ev, rec, records + PNJL_REC_DOUBLES * i);
    }
    ev.drain();
}

// ---- continuity lines ----------------------------------------------------------------------------
template <int G>
struct LineSink {
    const GroupEval<G>* ev;
    double* base;
    double xi;
    __device__ __forceinline__ void operator()(int it, const PointRes& r, double T_fm, double mu_fm, int n_fj, int n_th,
                                               int n_ft) {
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T_fm, mu_fm, xi, n_fj, n_th, n_ft, rec);
        store_record<G>(*ev, rec, base + (long long)PNJL_REC_DOUBLES * it);
    }
};

template <int G>
__global__ void __launch_bounds__(512, 1) k_scan_lines(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_mesh,
                                                    long long n_lines, const double* __restrict__ muq_MeV,
                                                    const double* __restrict__ xi, const int* __restrict__ table_idx,
                                                    int n_T, const double* __restrict__ T_MeV, double* __restrict__ records,
                                                    unsigned long long* counter, int mode, const long long* __restrict__ out_index) {
    extern __shared__ double s_mesh[];
    __shared__ int s_done;
    stage_mesh<G>(g_mesh, 3 * cfg->n_nodes + 2 * cfg->n_iso, s_mesh, &s_done);
    GroupEval<G> ev = make_eval<G>(cfg, s_mesh, &s_done);
    Solver<GroupEval<G>> sv(cfg->m, cfg->sp, ev);
    for (;;) {
        const long long l = next_task<G>(counter, ev);
        if (l >= n_lines) break;
        // mode 0: (xi, mu) line marching T (run_gap_transport_scan.jl); mode 1: (xi, T) line marching mu (TmuScan.jl);
        // mode 2: task l = branch (l & 1) of the (xi, T) line l >> 1 (DualBranchScan.jl), records [line][branch][mu]
        const long long li = mode == 2 ? (l >> 1) : l;
        LineSink<G> sink{&ev, records + (long long)PNJL_REC_DOUBLES * n_T * (out_index ? out_index[l] : l), xi[li]};
        if (mode == 0) scan_line(sv, &cfg->pt, table_idx ? table_idx[l] : -1, muq_MeV[l], xi[l], n_T, T_MeV, sink);
        else if (mode == 1) scan_tmu_line(sv, &cfg->pt, table_idx ? table_idx[l] : -1, muq_MeV[l], xi[l], n_T, T_MeV, sink);
        else scan_branch_line(sv, muq_MeV[li], xi[li], n_T, T_MeV, (int)(l & 1), sink);
    }
    ev.drain();
}

// ---- warp-specialised line march -------------------------------------------------------------------
// Workers and controllers.  In the layouts above every lane of a warp runs the scalar solver code redundantly
// (~3000 instructions per quadrature pass, a third of all issued instructions) and the warps of a CTA have to be
// phase-aligned to keep the quadrature loop in the instruction cache.  Here the two kinds of work get their own
// warps:
//   * NW worker warps only ever run quadrature loops + the warp reduction (a few KB of code, always cached);
//   * controller warps run the solver cascade with ONE LINE PER LANE (real SIMT: the scalar code is no longer
//     replicated 32x), and hand quadrature passes to the workers through mailboxes in shared memory.
// Every worker serves two mailboxes, so while it sweeps the mesh for one line the controller lane of its other
// line assembles F/J, does the 5x5 elimination and posts the next request: the FP64 pipe never waits for scalar
// code.  Hand-off is a sequence number per mailbox (volatile shared + __threadfence_block, short __nanosleep
// polls); there are no CTA barriers after the prologue, so divergent controller lanes cannot dead-lock.
// WS_FJ: Jacobian pass whose finish and 5x5 elimination run in the worker (lane-parallel, pnjl_lean.cuh) with the closed
// forms the controller lane ships in the mailbox: the worker returns F and the Newton direction.  WS_FJ_SUMS: the same
// pass returning the 20 sums (the trust-region method wants J itself).
enum { WS_FJ = 0, WS_FT = 1, WS_TH = 2, WS_EXIT = 3, WS_FJ_SUMS = 4 };
constexpr int kWsR = 21;   // doubles per result block: up to 20 sums + the fast-path flag
constexpr int kWsCf = 12;  // closed forms of a Jacobian pass: dI/dM, d2I/dM2 per flavour (6), U_P, U_Pb, U_PP, U_PPb, U_PbPb (5), pad

struct WsSlot {
    double d[9];        // T, mu, xi, x[5]
    double r[kWsR];     // WS_FJ: F[5], p[5], ok ; otherwise reduced sums (20 FJ | 5 + 8 fused | 8 thermo), r[20] = fast-path flag
    double cf[kWsCf];   // closed forms for WS_FJ (written by the controller lane while the worker sweeps)
    int type;
    int parts_done;     // parts of the current pass finished so far (atomic; only when a pass is split)
    volatile int cf_seq;  // round number for which cf[] is valid
    int pad;
};

// Closed forms with their libm fall-backs out of line (one copy each; the per-pass code of the controllers stays small).
__device__ __noinline__ void vacuum_terms_cold(double Lam, double M, double& I0, double& I1, double& I2) {
    vacuum_terms_t<false>(Lam, M, I0, I1, I2);
}
__device__ __noinline__ void polyakov_cold(const Model& m, double T, double iT, double P, double Pb, UTerms& u) {
    polyakov_eval<false, true>(m, T, iT, P, Pb, u);
}
__device__ __noinline__ void ctrl_vacuum(double Lam, double M, double& I0, double& I1, double& I2) {
    if (vacuum_tame(Lam, M)) vacuum_terms_t<true>(Lam, M, I0, I1, I2);
    else vacuum_terms_cold(Lam, M, I0, I1, I2);
}
__device__ __noinline__ void ctrl_polyakov(const Model& m, double T, double iT, double P, double Pb, bool with_value, UTerms& u) {
    if (polyakov_tame(P, Pb)) {
        if (with_value) polyakov_eval<true, true>(m, T, iT, P, Pb, u);
        else polyakov_eval<true, false>(m, T, iT, P, Pb, u);
    } else {
        polyakov_cold(m, T, iT, P, Pb, u);
    }
}

// One request queue per controller warp ("group"): its lanes fill their mailboxes, lane 0 publishes the round,
// any idle worker pulls the next work item, and the lanes resume when all mailboxes of the round are served.
struct WsGroup {
    volatile int round;      // published round number (0 = nothing yet)
    volatile int exit_flag;  // set when every task of the group is finished
    int next;                // next work item (mailbox x part) to hand out in the current round   (atomic)
    int done;                // mailboxes served in total, monotonic                                (atomic)
    int n_slots;             // mailboxes of this group
    int first_slot;          // index of its first mailbox
    volatile int ticket;     // CTA-wide sequence number of the published round: workers serve the oldest round first
    int pad;
};

// What the controller lanes work on: whole continuity lines or independent points.
struct WsTask {
    int mode;                // 0: lines (scan_line), 1: points (solve / solve_multi), 2: TmuScan lines (scan_tmu_line), 3: dual-branch lines
    long long n_tasks;
    long long perm_mult;     // tasks are handed out in the order (k * perm_mult) mod n_tasks (coprime multiplier): neighbouring
                             // lines cost alike, so a contiguous hand-out loads the SMs unevenly (measured 1.12 vs 1.02 max/mean)
    // lines
    const double* muq_MeV; const double* xi; const int* table_idx; int n_T; const double* T_MeV;
    // points
    const double* T_fm; const double* mu_fm; int seed_mode; int n_seeds; const double* seeds;
    double* records;
    const long long* out_index;   // optional: task t writes its rows at record line out_index[t] instead of t (lines modes)
    int worker_solve;             // see CtrlEval::worker_solve
};

struct CtrlEval {
    const Model* m;
    WsSlot* slot;
    WsGroup* group;
    int* ticket_ctr;  // CTA-wide round counter (shared memory)
    int seq;          // rounds posted so far
    unsigned grp;     // lanes of this controller warp that own a mailbox
    bool leader;      // lowest lane of the group
    bool finished;    // this lane has no task left
    bool worker_solve;   // Jacobian passes are finished and eliminated by the worker (WS_FJ) instead of by this lane
    double p2max, pc2max;
#ifdef PNJL_PROFILE_PHASES
    unsigned long long* dbg;
    long long t_ret, t_in;
#endif

    // Post one request / wait for the workers.  One code location each for every caller (noinline) and one wait group per
    // controller warp: all its lanes meet in post() once per round (lanes without tasks come from retire()), and again in
    // wait(), where they poll with one instruction stream and leave together, so the scalar code between passes runs
    // SIMT-converged.  Between the two the lane computes what does not depend on the sums (the closed forms of x), so that
    // work hides behind the workers' sweep.  post() returns true (without posting) once every task of the group is finished.
    __device__ __noinline__ bool post(int type, double T, double mu, double xi, const double x[5]) {
#ifdef PNJL_PROFILE_PHASES
        if (dbg && t_ret) { atomicAdd(dbg + 4, (unsigned long long)(clock64() - t_ret)); atomicAdd(dbg + 5, 1ULL); }
        t_in = clock64();
#endif
        slot->d[0] = T; slot->d[1] = mu; slot->d[2] = xi;
#pragma unroll
        for (int i = 0; i < 5; ++i) slot->d[3 + i] = x[i];
        slot->type = type;
        __threadfence_block();
        if (__ballot_sync(grp, finished) == grp) return true;
        ++seq;
        if (leader) {
            group->next = 0;
            group->ticket = atomicAdd(ticket_ctr, 1);
            __threadfence_block();
            group->round = seq;                // publish the round to the workers
        }
        return false;
    }
    __device__ __noinline__ void wait() {
        const unsigned target = (unsigned)seq * (unsigned)group->n_slots;
        for (;;) {
            const bool ready = (int)(*((volatile unsigned*)&group->done) - target) >= 0;     // wrap-safe
            if (__all_sync(grp, ready)) break;
            __nanosleep(64);
        }
        __threadfence_block();
#ifdef PNJL_PROFILE_PHASES
        t_ret = clock64();
        if (dbg) atomicAdd(dbg + 6, (unsigned long long)(t_ret - t_in));
#endif
    }
    __device__ __forceinline__ bool request(int type, double T, double mu, double xi, const double x[5]) {
        if (post(type, T, mu, xi, x)) return true;
        wait();
        return false;
    }
    // A lane without tasks left keeps posting empty requests so that the round size stays fixed; returns when the
    // whole group has run out of tasks.
    __device__ __noinline__ void retire() {
        finished = true;
        const double zero[5] = {0, 0, 0, 0, 0};
        while (!request(WS_EXIT, 0.0, 0.0, 0.0, zero)) {}   // WS_EXIT as a mailbox type = "nothing to do"
        if (leader) {
            __threadfence_block();
            group->exit_flag = 1;
        }
    }
    // closed forms of the three flavours (M_d taken from M_u when the masses coincide bitwise)
    __device__ __forceinline__ void vacuum_all(const PointCtx& c, double I0v[3], double I1v[3], double I2v[3]) const {
        ctrl_vacuum(m->Lambda, c.M[0], I0v[0], I1v[0], I2v[0]);
        if (c.M[1] == c.M[0]) { I0v[1] = I0v[0]; I1v[1] = I1v[0]; I2v[1] = I2v[0]; }
        else ctrl_vacuum(m->Lambda, c.M[1], I0v[1], I1v[1], I2v[1]);
        ctrl_vacuum(m->Lambda, c.M[2], I0v[2], I1v[2], I2v[2]);
    }
    __device__ __noinline__ void fj(double T, double mu, double xi, const double x[5], double F[5], double J[25]) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        post(WS_FJ_SUMS, T, mu, xi, x);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        ctrl_polyakov(*m, c.T, c.invT, c.Phi, c.Phib, false, u);
        wait();
        double acc[kFJAcc];
#pragma unroll
        for (int i = 0; i < kFJAcc; ++i) acc[i] = slot->r[i];
        finish_fj_pre(*m, c, x, acc, I1v, I2v, u, F, J, slot->r[20] != 0.0);
    }
    // F(x) and the Newton direction: the worker that sweeps the mesh also assembles [J | F] and eliminates (lane-parallel);
    // this lane contributes the closed forms, computed while the sweep runs.
    __device__ __noinline__ bool fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        if (!worker_solve) {
            // the sums come back and this lane assembles J and eliminates in registers (cheap when many lines share the
            // controller warp's instruction stream); the closed forms are still computed while the workers sweep
            post(WS_FJ_SUMS, T, mu, xi, x);
            double I0v[3], I1v[3], I2v[3];
            vacuum_all(c, I0v, I1v, I2v);
            UTerms u;
            ctrl_polyakov(*m, c.T, c.invT, c.Phi, c.Phib, false, u);
            wait();
            double acc[kFJAcc], J[25], b[5];
#pragma unroll
            for (int i = 0; i < kFJAcc; ++i) acc[i] = slot->r[i];
            finish_fj_pre(*m, c, x, acc, I1v, I2v, u, F, J, slot->r[20] != 0.0);
#pragma unroll
            for (int i = 0; i < 5; ++i) b[i] = F[i];
            const bool ok = lu_solve5_regs(J, b, p);
#pragma unroll
            for (int i = 0; i < 5; ++i) p[i] = -p[i];
            return ok;
        }
        post(WS_FJ, T, mu, xi, x);
        {
            double I0v[3], I1v[3], I2v[3];
            vacuum_all(c, I0v, I1v, I2v);
            UTerms u;
            ctrl_polyakov(*m, c.T, c.invT, c.Phi, c.Phib, false, u);
#pragma unroll
            for (int i = 0; i < 3; ++i) { slot->cf[2 * i] = I1v[i]; slot->cf[2 * i + 1] = I2v[i]; }
            slot->cf[6] = u.U_P; slot->cf[7] = u.U_Pb; slot->cf[8] = u.U_PP; slot->cf[9] = u.U_PPb; slot->cf[10] = u.U_PbPb;
            __threadfence_block();
            slot->cf_seq = seq;
        }
        wait();
#pragma unroll
        for (int i = 0; i < 5; ++i) { F[i] = slot->r[i]; p[i] = slot->r[5 + i]; }
        return slot->r[10] != 0.0;
    }
    __device__ __noinline__ bool f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        const double k2max = p2max + (xi > 0.0 ? xi * pc2max : 0.0);
        if (!fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) return false;
        post(WS_FT, T, mu, xi, x);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        ctrl_polyakov(*m, c.T, c.invT, c.Phi, c.Phib, true, u);
        wait();
        double facc[kFtAcc], tacc[kThAcc];
#pragma unroll
        for (int i = 0; i < kFtAcc; ++i) facc[i] = slot->r[i];
#pragma unroll
        for (int i = 0; i < kThAcc; ++i) tacc[i] = slot->r[kFtAcc + i];
        finish_f_pre(*m, c, x, facc, I1v, u, F);
        finish_thermo_pre(*m, c, x, tacc, I0v, u, th);
        return true;
    }
    __device__ __noinline__ void thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        post(WS_TH, T, mu, xi, x);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        ctrl_polyakov(*m, c.T, c.invT, c.Phi, c.Phib, true, u);
        wait();
        double tacc[kThAcc];
#pragma unroll
        for (int i = 0; i < kThAcc; ++i) tacc[i] = slot->r[i];
        finish_thermo_pre(*m, c, x, tacc, I0v, u, th);
    }
};

__device__ __forceinline__ void ctrl_store_record(const double rec[PNJL_REC_DOUBLES], double* out) {
#pragma unroll
    for (int q = 0; q < PNJL_REC_DOUBLES; q += 2) *reinterpret_cast<double2*>(out + q) = make_double2(rec[q], rec[q + 1]);
}

struct CtrlSink {
    double* base;
    double xi;
    __device__ __forceinline__ void operator()(int it, const PointRes& r, double T_fm, double mu_fm, int n_fj, int n_th,
                                               int n_ft) {
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T_fm, mu_fm, xi, n_fj, n_th, n_ft, rec);
        ctrl_store_record(rec, base + (long long)PNJL_REC_DOUBLES * it);
    }
};

// Warp sum of N per-lane values with N/2 + N/4 + ... shuffles instead of 5 N ("transposed" butterfly): at the step
// with lane offset OFF the lanes whose OFF bit is clear keep the first half of the values and the others the second
// half; each lane sends the half it gives up and adds what it receives to the half it keeps.  Every kept value is
// a(l) + a(l ^ OFF) exactly as in the plain xor butterfly (v += shfl_xor(v, off), off = 16 .. 1), so after five steps
// the sums are bit-identical to that butterfly's; they end up spread over the lanes: `idx` is the index of the one
// sum this lane holds in v[0], `valid` is false for the lanes that hold padding.
template <int N, int OFF>
struct WarpSumT {
    static __device__ __forceinline__ void run(double* v, int lane, int& idx, bool& valid) {
        constexpr int H = (N + 1) / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const double hi = (H + j < N) ? v[H + j] : 0.0;
            const double keep = up ? hi : v[j];
            const double send = up ? v[j] : hi;
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        WarpSumT<H, OFF / 2>::run(v, lane, idx, valid);
        idx += up ? H : 0;
        valid = valid && idx < N;
    }
};
template <int N>
struct WarpSumT<N, 0> {
    static_assert(N == 1, "at most 32 values");
    static __device__ __forceinline__ void run(double*, int, int& idx, bool& valid) { idx = 0; valid = true; }
};
template <int N>
__device__ __forceinline__ void warp_sum_store(double (&v)[N], int lane, double* out) {
    int idx;
    bool valid;
    WarpSumT<N, 16>::run(v, lane, idx, valid);
    if (valid) out[idx] = v[0];
}

// One quadrature pass (or one part of it: nodes lane + 32 part, stride 32 parts) of a worker warp for mailbox `sl`;
// the warp-reduced sums go to out[0..20].
__device__ __noinline__ void ws_worker_pass(const DeviceConfig* cfg, const MeshView& mv, const double* d /* T, mu, xi, x[5] */, int type,
                                            int lane, int part, int parts, double* out) {
    const double T = d[0], mu = d[1], xi = d[2];
    double x[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) x[i] = d[3 + i];
    PointCtx c;
    make_ctx(cfg->m, T, mu, xi, x, c);
    const bool iso = cfg->sp.isospin != 0;
    const int l0 = lane + 32 * part, stride = 32 * parts;
    if (type == WS_FJ) {
        double acc[kFJAcc];
        const bool fast = fj_partial(cfg->m, iso, c, x, mv, l0, stride, acc);
        warp_sum_store(acc, lane, out);
        if (lane == 0) out[20] = fast ? 1.0 : 0.0;
    } else if (type == WS_FT) {
        double acc[kFtAcc + kThAcc];
        ft_partial(cfg->m, iso, c, x, mv, l0, stride, acc, acc + kFtAcc);
        warp_sum_store(acc, lane, out);
    } else {
        double tacc[kThAcc];
        thermo_partial(cfg->m, iso, c, x, mv, l0, stride, tacc);
        warp_sum_store(tacc, lane, out);
    }
}

// The finish of a WS_FJ pass in the worker warp that swept the mesh (or that added up the last part): [J | F] one entry per
// lane from the sums in W[LW_S ..] and the closed forms of the mailbox, elimination and back substitution (pnjl_lean.cuh).
// Leaves F[5], p[5] = -J^{-1} F and the non-singular flag in the mailbox.
__device__ __noinline__ void ws_worker_solve(const DeviceConfig* cfg, WsSlot* sl, double* W, int lane, int round) {
    while (sl->cf_seq != round) __nanosleep(32);      // the controller lane writes them while we sweep: normally long there
    __threadfence_block();
    const double T = sl->d[0], mu = sl->d[1], xi = sl->d[2];
    double x[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) x[i] = sl->d[3 + i];
    PointCtx c;
    make_ctx(cfg->m, T, mu, xi, x, c);
    LeanConst k;
    lean_consts(cfg->m, c.T, c.invT, k);
    if (lane < 3) {
        const double Mf = lane == 0 ? c.M[0] : (lane == 1 ? c.M[1] : c.M[2]);
        const double M2f = lane == 0 ? c.M2[0] : (lane == 1 ? c.M2[1] : c.M2[2]);
        lean_flavour_fj_pre(lane, k, Mf, M2f, W[LW_S + 20] != 0.0, sl->cf[2 * lane], sl->cf[2 * lane + 1], W);
    } else if (lane < 12) {
        const int r = (lane - 3) / 3, j = (lane - 3) - 3 * r;
        lean_dtable(r, j, k, x, W);
    } else if (lane == 12) {
#pragma unroll
        for (int q = 0; q < 5; ++q) { W[LW_X + q] = x[q]; W[LW_U + q] = sl->cf[6 + q]; }
    }
    __syncwarp();
    const int li = lane / 6, lc = lane - 6 * li;
    double a = 0.0;
    if (lane < 30) {
        a = lean_aug_entry(li, lc, k, W, ACC_GP, ACC_GPB);
        W[LW_AUG + lane] = a;
    }
    __syncwarp();
    double F[5], y[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) F[i] = W[LW_AUG + 6 * i + 5];
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
    lean_backsub(W, y);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) { sl->r[i] = F[i]; sl->r[5 + i] = -y[i]; }
        sl->r[10] = ok ? 1.0 : 0.0;
    }
    __syncwarp();
}

constexpr int kWsMaxGroups = 8;
#ifndef PNJL_WS_MAX_THREADS
#define PNJL_WS_MAX_THREADS 512   // 16 warps x 128 registers fill the register file of an SM
#endif
constexpr int kWsMaxWarps = PNJL_WS_MAX_THREADS / 32;

// blockDim.x = 32 * (n_workers + n_ctrl_warps).  n_slots mailboxes (one task each) are spread evenly over the
// controller warps; a pass may be split into `parts` node ranges served by different workers (few tasks per SM).
// Dynamic shared memory: mesh [3 n] | mailboxes [n_slots] | partial sums [n_slots][parts][21] (parts > 1 only).
__global__ void __launch_bounds__(PNJL_WS_MAX_THREADS, 1) k_solve_ws(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_mesh,
                                                     WsTask task, unsigned long long* counter, int n_workers,
                                                     int n_ctrl_warps, int n_slots, int parts) {
    extern __shared__ double s_dyn[];
    __shared__ WsGroup s_groups[kWsMaxGroups];
    __shared__ int s_ticket;
    const int n = cfg->n_nodes;
    double* s_mesh = s_dyn;
    const int n_mesh = 3 * n + 2 * cfg->n_iso;
    WsSlot* s_slots = reinterpret_cast<WsSlot*>(s_dyn + ((n_mesh + 1) & ~1));
    double* s_part = reinterpret_cast<double*>(s_slots + n_slots);
    double* s_work = s_part + (parts > 1 ? (size_t)kWsR * n_slots * parts : 0) + (parts > 1 ? ((kWsR * n_slots * parts) & 1) : 0);
    for (int i = threadIdx.x; i < n_mesh; i += blockDim.x) s_mesh[i] = g_mesh[i];
    // One request group per controller warp.  (Two independent groups per warp — divergent half-warps with shorter rounds —
    // were measured 16 % slower on cfg5: the halves serialise their scalar phases.)
    const int n_groups = n_ctrl_warps;
    const int per = (n_slots + n_groups - 1) / n_groups;
    if (threadIdx.x < n_groups) {
        const int cw = threadIdx.x;
        int cnt = n_slots - cw * per;
        cnt = cnt < 0 ? 0 : (cnt > per ? per : cnt);
        s_groups[cw].round = 0;
        s_groups[cw].exit_flag = cnt == 0 ? 1 : 0;
        s_groups[cw].next = 0;
        s_groups[cw].done = 0;
        s_groups[cw].n_slots = cnt;
        s_groups[cw].first_slot = cw * per;
        s_groups[cw].ticket = 0;
        if (cw == 0) s_ticket = 0;
    }
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
        s_slots[i].type = WS_EXIT;
        s_slots[i].parts_done = 0;
        s_slots[i].cf_seq = 0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < n_workers) {
        // ---------------- worker: pull work items from whichever group has a round open ----------------
        MeshView mv;
        mv.p2 = s_mesh; mv.pc2 = s_mesh + n; mv.coef = s_mesh + 2 * n; mv.n = n;
        mv.p2max = cfg->p2max; mv.pc2max = cfg->pc2max;
        mv.p2_iso = s_mesh + 3 * n; mv.coef_iso = s_mesh + 3 * n + cfg->n_iso; mv.n_iso = cfg->n_iso;
#ifdef PNJL_PROFILE_PHASES
        long long t_idle = 0;
#endif
        for (;;) {
            // Pick the open round that was published first (FIFO over the groups).  Serving the groups in turn locks
            // their rounds in phase: all controller warps then sit in their scalar phase (~28 k cycles per round) at
            // the same time and the workers starve meanwhile (measured: 13 % of worker time, scripts/ws_queue_sim.py
            // reproduces it); a fixed priority is worse still.  Oldest-first lets the rounds drift apart, so one
            // group's scalar phase hides behind another group's round.
            bool any_live = false, did = false;
            int best = -1, best_ticket = 0;
            for (int g = 0; g < n_groups; ++g) {
                WsGroup* gr = &s_groups[g];
                if (gr->exit_flag) continue;
                any_live = true;
                const int round = gr->round;
                if (round == 0) continue;
                // rounds are strictly sequential per group: all mailboxes of round r are served before r+1 opens
                if (*((volatile int*)&gr->done) >= round * gr->n_slots) continue;
                if (*((volatile int*)&gr->next) >= gr->n_slots * parts) continue;   // the round is fully handed out
                const int tk = gr->ticket;
                if (best < 0 || tk - best_ticket < 0) { best = g; best_ticket = tk; }
            }
            if (best >= 0) {
                WsGroup* gr = &s_groups[best];
                int idx = 0;
                if (lane == 0) idx = atomicAdd(&gr->next, 1);
                idx = __shfl_sync(0xffffffffu, idx, 0);
                if (idx >= gr->n_slots * parts) continue;    // somebody else took the last item
                __threadfence_block();
                const int round = gr->round;                 // cannot advance before this item is served
                const int si = gr->first_slot + idx / parts, part = idx % parts;
                WsSlot* sl = &s_slots[si];
                const int type = sl->type;
#ifdef PNJL_PROFILE_PHASES
                const long long tp0 = clock64();
                if (lane == 0 && cfg->dbg && t_idle) atomicAdd(cfg->dbg + 3, (unsigned long long)(tp0 - t_idle));
#endif
                bool slot_complete = true;
                const int ptype = type == WS_FJ_SUMS ? WS_FJ : type;     // kind of sweep
                double* Wl = s_work + (size_t)warp * LW_END;             // this worker's scratch line (WS_FJ finish)
                if (type != WS_EXIT) {
                    if (parts == 1) {
                        ws_worker_pass(cfg, mv, sl->d, ptype, lane, 0, 1, type == WS_FJ ? Wl + LW_S : sl->r);
                        if (type == WS_FJ) {
                            __syncwarp();
                            ws_worker_solve(cfg, sl, Wl, lane, round);
                        }
                    } else {
                        double* mine = s_part + ((size_t)si * parts + part) * kWsR;
                        ws_worker_pass(cfg, mv, sl->d, ptype, lane, part, parts, mine);
                        __syncwarp();
                        __threadfence_block();
                        int c = 0;
                        if (lane == 0) c = atomicAdd(&sl->parts_done, 1);
                        c = __shfl_sync(0xffffffffu, c, 0);
                        slot_complete = (c == parts - 1);
                        if (slot_complete) {
                            // last part in: add the partial sums in part order (the same order whoever finishes last)
                            __threadfence_block();
                            if (lane < kWsR) {
                                const double* base = s_part + (size_t)si * parts * kWsR + lane;
                                double v = base[0];
                                for (int q = 1; q < parts; ++q) v += base[q * kWsR];
                                v = (lane == 20) ? base[0] : v;
                                if (type == WS_FJ) Wl[LW_S + lane] = v;
                                else sl->r[lane] = v;
                            }
                            if (lane == 0) sl->parts_done = 0;
                            if (type == WS_FJ) {
                                __syncwarp();
                                ws_worker_solve(cfg, sl, Wl, lane, round);
                            }
                        }
                    }
                } else if (parts > 1) {
                    int c = 0;
                    if (lane == 0) c = atomicAdd(&sl->parts_done, 1);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    slot_complete = (c == parts - 1);
                    if (slot_complete && lane == 0) sl->parts_done = 0;
                }
                __syncwarp();
                __threadfence_block();
                if (slot_complete && lane == 0) atomicAdd(&gr->done, 1);
                did = true;
#ifdef PNJL_PROFILE_PHASES
                t_idle = clock64();
                if (lane == 0 && cfg->dbg && type != WS_EXIT) { atomicAdd(cfg->dbg + 0, (unsigned long long)(t_idle - tp0)); atomicAdd(cfg->dbg + 2, 1ULL); }
#endif
            }
            if (!any_live) break;
            if (!did) __nanosleep(128);
        }
        return;
    }
    // ---------------- controller: lane <-> mailbox <-> one task at a time ----------------
    const int cw = warp - n_workers;
    WsGroup* gr = &s_groups[cw];
    if (lane >= gr->n_slots) return;
    CtrlEval ev;
    ev.m = &cfg->m;
    ev.slot = &s_slots[gr->first_slot + lane];
    ev.group = gr;
    ev.ticket_ctr = &s_ticket;
    ev.seq = 0;
    ev.grp = gr->n_slots >= 32 ? 0xffffffffu : ((1u << gr->n_slots) - 1u);
    ev.leader = lane == 0;
    ev.finished = false;
    ev.worker_solve = task.worker_solve != 0;
    ev.p2max = cfg->p2max;
    ev.pc2max = cfg->pc2max;
#ifdef PNJL_PROFILE_PHASES
    ev.dbg = cfg->dbg;
    ev.t_ret = 0;
#endif
    Solver<CtrlEval> sv(cfg->m, cfg->sp, ev);
    // First task of every mailbox: CTA b, mailbox j -> ticket b + gridDim.x * j, so that a grid that holds all tasks in one
    // wave spreads them evenly over ALL its CTAs (8192 lines on 148 SMs: 55 or 56 per SM instead of 147 x 56 and one idle
    // SM); later tasks come from the global counter.
    bool first_task = true;
    for (;;) {
        long long tk;
        if (first_task) {
            tk = (long long)blockIdx.x + (long long)gridDim.x * (gr->first_slot + lane);
            first_task = false;
        } else {
            tk = (long long)gridDim.x * n_slots + (long long)atomicAdd(counter, 1ULL);
        }
        if (tk >= task.n_tasks) break;
        const long long t = (long long)(((unsigned long long)tk * (unsigned long long)task.perm_mult) % (unsigned long long)task.n_tasks);
        if (task.mode == 0) {
            CtrlSink sink{task.records + (long long)PNJL_REC_DOUBLES * task.n_T * (task.out_index ? task.out_index[t] : t), task.xi[t]};
            scan_line(sv, &cfg->pt, task.table_idx ? task.table_idx[t] : -1, task.muq_MeV[t], task.xi[t], task.n_T, task.T_MeV,
                      sink);
        } else if (task.mode == 2) {
            // here muq_MeV holds the line's T_MeV and T_MeV the shared mu grid
            CtrlSink sink{task.records + (long long)PNJL_REC_DOUBLES * task.n_T * (task.out_index ? task.out_index[t] : t), task.xi[t]};
            scan_tmu_line(sv, &cfg->pt, task.table_idx ? task.table_idx[t] : -1, task.muq_MeV[t], task.xi[t], task.n_T,
                          task.T_MeV, sink);
        } else if (task.mode == 3) {
            // dual-branch scan: task t = branch (t & 1) of line t >> 1; muq_MeV holds the lines' T_MeV, T_MeV the mu grid
            const long long li = t >> 1;
            CtrlSink sink{task.records + (long long)PNJL_REC_DOUBLES * task.n_T * (task.out_index ? task.out_index[t] : t), task.xi[li]};
            scan_branch_line(sv, task.muq_MeV[li], task.xi[li], task.n_T, task.T_MeV, (int)(t & 1), sink);
        } else {
            const double T = task.T_fm[t], mu = task.mu_fm[t], x_i = task.xi[t];
            sv.set_point(T, mu, x_i);
            sv.n_fj = 0; sv.n_th = 0; sv.n_ft = 0;
            PointRes r;
            if (task.seed_mode == PNJL_SEED_EXPLICIT && task.n_seeds == 1) {
                double x0[5];
                copy5(x0, task.seeds + 5 * t);
                sv.solve_with_fallback(x0, r);
            } else if (task.seed_mode == PNJL_SEED_EXPLICIT) {
                sv.solve_multi(task.seeds + 5 * (long long)task.n_seeds * t, task.n_seeds, r);
            } else if (task.seed_mode == PNJL_SEED_AUTO) {
                double x0[5];
                default_seed(2, T, mu, x0);
                sv.solve_with_fallback(x0, r);
            } else {
                sv.solve_multi(nullptr, 6, r);
            }
            double rec[PNJL_REC_DOUBLES];
            fill_record(r, T, mu, x_i, sv.n_fj, sv.n_th, sv.n_ft, rec);
            ctrl_store_record(rec, task.records + (long long)PNJL_REC_DOUBLES * t);
        }
    }
    ev.retire();
}

#include "pnjl_march.cuh"

// ---- single FJ evaluation (test hook) ------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(512, 1) k_eval_fj(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_mesh,
                                                 long long n, const double* __restrict__ T_fm, const double* __restrict__ mu_fm,
                                                 const double* __restrict__ xi, const double* __restrict__ x,
                                                 double* __restrict__ FJ, unsigned long long* counter, int with_thermo) {
    extern __shared__ double s_mesh[];
    __shared__ int s_done;
    stage_mesh<G>(g_mesh, 3 * cfg->n_nodes + 2 * cfg->n_iso, s_mesh, &s_done);
    GroupEval<G> ev = make_eval<G>(cfg, s_mesh, &s_done);
    const int stride = with_thermo == 2 ? PNJL_DERIV_DOUBLES : (with_thermo ? PNJL_STATE_DOUBLES : 30);
    for (;;) {
        const long long i = next_task<G>(counter, ev);
        if (i >= n) break;
        double xs[5], F[5], J[25], gp2[2];
        copy5(xs, x + 5 * i);
        ev.fj(T_fm[i], mu_fm[i], xi[i], xs, F, J, gp2);
        double* o = FJ + (long long)stride * i;
        if (ev.lane == 0) {
            for (int q = 0; q < 5; ++q) o[q] = F[q];
            for (int q = 0; q < 25; ++q) o[5 + q] = J[q];
        }
        if (with_thermo) {
            // the thermodynamic functions at the same (not necessarily converged) state: Thermodynamics.jl:215-281
            Thermo th;
            ev.thermo(T_fm[i], mu_fm[i], xi[i], xs, th);
            if (ev.lane == 0) {
                o[30] = th.omega; o[31] = th.pressure; o[32] = th.rho_norm; o[33] = th.entropy; o[34] = th.energy;
                for (int q = 0; q < 3; ++q) { o[35 + q] = th.rho[q]; o[38 + q] = th.nq[q]; o[41 + q] = th.nqb[q]; o[44 + q] = th.M[q]; }
                o[47] = 0.0;
            }
        }
        if (with_thermo == 2) {
            double d[16];
            ev.dtheta(T_fm[i], mu_fm[i], xi[i], xs, gp2[0], gp2[1], d);
            if (ev.lane == 0) {
                for (int q = 0; q < 16; ++q) o[48 + q] = d[q];
            }
        }
    }
    ev.drain();
}

// ---- one-loop integral A + effective couplings (build_K_data step of the scan script) ---------------------------
// One thread per point: n_rule nodes x 2 flavours in a register loop, rule (p^2, w p^2) in shared memory.  Inputs are
// strided so that the same kernel reads either result records (stride 32) or plain arrays (stride 1).
struct CouplingsIn {
    const double* T; const double* mu; const double* m_u; const double* m_s; const double* Phi; const double* Phib;
    int stride;
};
__global__ void __launch_bounds__(128) k_couplings(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_rule, int n_rule,
                                                   long long n, CouplingsIn in, double* __restrict__ aux) {
    extern __shared__ double s_rule[];
    for (int i = threadIdx.x; i < 2 * n_rule; i += blockDim.x) s_rule[i] = g_rule[i];
    __syncthreads();
    const Model& m = cfg->m;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long o = i * in.stride;
        const double T = in.T[o], mu = in.mu[o], mq = in.m_u[o], ms = in.m_s[o], P = in.Phi[o], Pb = in.Phib[o];
        const double A_u = oneloop_A(m.Lambda, mq, mu, T, P, Pb, n_rule, s_rule, s_rule + n_rule);
        const double A_s = oneloop_A(m.Lambda, ms, mu, T, P, Pb, n_rule, s_rule, s_rule + n_rule);
        double a[kAuxDoubles];
        effective_couplings(m.G, m.K, m.Nc, mq, ms, A_u, A_s, a);
        double* out = aux + i * kAuxDoubles;
#pragma unroll
        for (int q = 0; q < kAuxDoubles; q += 2) *reinterpret_cast<double2*>(out + q) = make_double2(a[q], a[q + 1]);
    }
}

// ---- FP64 FMA peak microbenchmark ----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---- accuracy self-test of the branch-free primitives (test hook) ---------------------------------
__global__ void k_selftest_math(long long n, const double* __restrict__ x, int which, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    out[i] = which == 0 ? fast_exp_nonpos(v) : (which == 1 ? fast_rcp(v) : fast_rsqrt(v));
}

}  // namespace pnjl

// =====================================================================================================
// Host side: handle, buffers, C ABI
// =====================================================================================================
using namespace pnjl;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(PNJL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
    } while (0)

// Gauss-Legendre nodes on [-1, 1]: Newton on P_n with the three-term recurrence in long double,
// Tricomi initial guess; symmetric fill.  (Stands in for FastGaussQuadrature.gausslegendre.)
void gausslegendre_std(int n, std::vector<double>& x, std::vector<double>& w) {
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    const long double pi = 3.141592653589793238462643383279502884L;
    const int half = (n + 1) / 2;
    for (int i = 0; i < half; ++i) {
        const long double th = pi * (4 * (i + 1) - 1) / (4 * n + 2);
        long double z = (1 - (n - 1) / (8.0L * n * n * n)) * cosl(th);
        long double dp = 1;
        for (int iter = 0; iter < 64; ++iter) {
            long double pm1 = 1, p = z;  // P_0, P_1
            for (int k = 2; k <= n; ++k) {
                const long double pk = ((2 * k - 1) * z * p - (k - 1) * pm1) / k;
                pm1 = p;
                p = pk;
            }
            if (n == 1) { p = z; pm1 = 1; }
            dp = n * (pm1 - z * p) / (1 - z * z);
            const long double dz = p / dp;
            z -= dz;
            if (fabsl(dz) < 1e-20L) break;
        }
        // re-evaluate derivative at the converged node
        long double pm1 = 1, p = z;
        for (int k = 2; k <= n; ++k) {
            const long double pk = ((2 * k - 1) * z * p - (k - 1) * pm1) / k;
            pm1 = p;
            p = pk;
        }
        dp = n * (pm1 - z * p) / (1 - z * z);
        const long double wt = 2 / ((1 - z * z) * dp * dp);
        x[n - 1 - i] = (double)z;
        x[i] = (double)(-z);
        w[n - 1 - i] = (double)wt;
        w[i] = (double)wt;
    }
    if (n % 2 == 1) x[n / 2] = 0.0;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct pnjl_handle {
    int device = 0;
    int sm_count = 0;
    int n_nodes = 0;
    int G = 32;
    int G_user = 0;               // lanes_per_solve given by the caller (0 = automatic)
    int G_next = 0;               // layout for the next launch only (host entry points: all-isotropic batches), 0 = G
    int block_threads = 128;
    bool block_threads_forced = false;   // PNJL_BLOCK_THREADS given
    int schedule = 2;             // 0: every warp owns a line (phase-aligned CTAs); 1: worker/controller warps (k_solve_ws);
                                  // 2: automatic — k_march or k_solve_ws for lines by batch shape (dispatch_lines), k_solve_ws for
                                  // points; 3: line march in registers, time-sliced lines (k_march) for every line batch
    int ws_workers = 14, ws_ctrl_warps = 2, ws_spw = 4, ws_slots = 0;
    int march_parts = 0;          // warps per team in k_march (0 = automatic)
    int march_quantum = 0;        // points per time slice in k_march (0 = automatic)
    bool iso_batch = false;       // option "isotropic_batch": the caller promises xi == 0 on every line (device entry points)
    bool iso_next = false;        // the next launch is all-isotropic (set by the host entry points, which see xi)
    DevBuf march_state, march_slots, march_counters;
    DevBuf pt_start, pt_tcep, pt_T, pt_mu;   // flattened phase-boundary tables (pnjl_set_boundaries)
    DeviceConfig host_cfg;
    DeviceConfig* d_cfg = nullptr;
    double* d_mesh = nullptr;
    unsigned long long* d_counter = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf in_T, in_mu, in_xi, in_seeds, in_idx, in_x, out_rec, out_aux, in_c[6];
    const long long* out_index_next = nullptr;   // per-line output index for the next lines launch only (device pointer)
    double* d_rule = nullptr;      // one-loop rule: p^2 [n] | w p^2 [n]
    int n_rule = 0;
    pnjl_stats stats;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class K>
int launch_geometry(pnjl_handle* h, K kernel, size_t smem, long long n_groups_needed, int G, int* blocks, int* threads) {
    // CTA size: h->block_threads when there is enough work to fill every SM with such CTAs, otherwise the
    // largest warp count per CTA that still gives every SM a CTA (few lines per GPU in multi-GPU runs).
    int block = (G == 32 || h->block_threads_forced) ? h->block_threads : 128;   // small CTAs for the 8/16-lane layouts
    {
        const long long warps_needed = (n_groups_needed * G + 31) / 32;
        long long per_sm = (warps_needed + h->sm_count - 1) / h->sm_count;
        if (per_sm < 1) per_sm = 1;
        if (per_sm * 32 < block) block = (int)per_sm * 32;
    }
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) return fail(PNJL_ERR_CUDA, "kernel does not fit on an SM");
    const long long groups_per_block = block / G;
    long long need = (n_groups_needed + groups_per_block - 1) / groups_per_block;
    long long cap = (long long)per_sm * h->sm_count;
    if (need < 1) need = 1;
    *blocks = (int)(need < cap ? need : cap);
    *threads = block;
    h->stats.regs_per_thread = fa.numRegs;
    h->stats.smem_bytes = (int)smem;
    h->stats.blocks = *blocks;
    h->stats.threads = block;
    h->stats.lanes_per_solve = G;
    return PNJL_OK;
}

template <int G>
int launch_points(pnjl_handle* h, long long n, const double* T, const double* mu, const double* xi, int seed_mode,
                  int n_seeds, const double* seeds, double* rec, cudaStream_t st) {
    const size_t smem = sizeof(double) * (3 * h->n_nodes + 2 * h->host_cfg.n_iso);
    int blocks, threads;
    int rc = launch_geometry(h, k_solve_points<G>, smem, n, G, &blocks, &threads);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), st));
    k_solve_points<G><<<blocks, threads, smem, st>>>(h->d_cfg, h->d_mesh, n, T, mu, xi, seed_mode, n_seeds, seeds, rec,
                                                     h->d_counter);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}

// Launch geometry of the warp-specialised kernel for `n_tasks` lines/points on this GPU.
long long coprime_multiplier(long long n) {
    if (n < 4) return 1;
    long long m = (long long)(0.6180339887 * (double)n) | 1;       // near the golden ratio: consecutive hand-outs far apart
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    while (m < n && gcd(m, n) != 1) m += 2;
    return m < n ? m : 1;
}

int launch_ws(pnjl_handle* h, const WsTask& task_in, cudaStream_t st) {
    WsTask task = task_in;
    // Task order.  The first wave is dealt round-robin (CTA b, mailbox j -> ticket b + gridDim * j, see k_solve_ws), so with
    // the identity order every SM samples the (xi, mu) index space with a fixed stride — a stratified sample of a cost that
    // varies smoothly with the line index: SM load max/mean 1.017 on cfg5 against 1.035 for a pseudo-random order
    // (scripts/line_balance.py data).  The golden-ratio permutation dates from the contiguous hand-out (PNJL_WS_PERM=1 restores it).
    task.perm_mult = (getenv("PNJL_WS_PERM") && atoi(getenv("PNJL_WS_PERM")) != 0) ? coprime_multiplier(task.n_tasks) : 1;
    task.worker_solve = getenv("PNJL_WS_WSOLVE") ? atoi(getenv("PNJL_WS_WSOLVE")) : 0;
    int nw = h->ws_workers, nc = h->ws_ctrl_warps, spw = h->ws_spw, parts = 1;
    const long long per_sm = (task.n_tasks + h->sm_count - 1) / h->sm_count;   // tasks an SM has to carry at least
    long long n_slots = h->ws_slots > 0 ? h->ws_slots : (long long)spw * nw;
    if (per_sm < n_slots) {
        // few tasks per SM (multi-GPU slabs, small grids): one mailbox per task and split every pass over
        // several workers so that all workers stay busy
        n_slots = per_sm < 1 ? 1 : per_sm;
        while (parts < 4 && n_slots * parts * 2 <= 2LL * nw) parts *= 2;
        // measured on cfg5 slabs (scripts/strong_slab_variants.sh): with 7 lines per SM two parts beat four (9.8 vs 9.2 M
        // points/s; 8 and 16 parts are slower still) — the per-part prologue/reduction costs more than the shorter sweep saves
        if (n_slots >= 4 && parts > 2) parts = 2;
        if (n_slots * parts < nw) nw = (int)(n_slots * parts);
    }
    if (getenv("PNJL_WS_PARTS")) parts = atoi(getenv("PNJL_WS_PARTS"));
    if (parts < 1) parts = 1;
    if (parts > 16) parts = 16;
    while ((n_slots + nc - 1) / nc > 32 && nc < kWsMaxGroups) ++nc;
    if (nc > (int)n_slots) nc = (int)n_slots;
    if (nw + nc > kWsMaxWarps) nw = kWsMaxWarps - nc;
    // protocol limits: a controller warp owns at most 32 mailboxes (one per lane), there are at most kWsMaxGroups of them
    // and at least one worker must remain; anything else (PNJL_WS_* experiments) would wait for mailboxes nobody serves
    if (nw < 1 || nc < 1 || nc > kWsMaxGroups || (n_slots + nc - 1) / nc > 32)
        return fail(PNJL_ERR_ARG, "warp-specialised kernel: impossible mailbox layout (PNJL_WS_SLOTS / PNJL_WS_CTRL / PNJL_WS_WORKERS)");
    const int threads = 32 * (nw + nc);
    const size_t smem = sizeof(double) * (size_t)(((3 * h->n_nodes + 2 * h->host_cfg.n_iso) + 1) & ~1) + sizeof(WsSlot) * (size_t)n_slots +
                        (parts > 1 ? sizeof(double) * (((size_t)kWsR * n_slots * parts + 1) & ~(size_t)1) : 0) +
                        sizeof(double) * (size_t)LW_END * (size_t)nw;
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_solve_ws));
    if (smem > 40 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_solve_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long need = (task.n_tasks + n_slots - 1) / n_slots;
    int per_sm_ctas = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_ctas, k_solve_ws, threads, smem));
    if (per_sm_ctas < 1) return fail(PNJL_ERR_CUDA, "warp-specialised kernel does not fit on an SM");
    const long long cap = (long long)per_sm_ctas * h->sm_count;
    // with at least one task per SM every SM gets a CTA (the tasks are dealt out round-robin, see k_solve_ws)
    const int blocks = (int)(task.n_tasks >= cap ? cap : (need < cap ? need : cap));
    h->stats.regs_per_thread = fa.numRegs;
    h->stats.smem_bytes = (int)smem;
    h->stats.blocks = blocks;
    h->stats.threads = threads;
    h->stats.lanes_per_solve = 32 * parts;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), st));
    k_solve_ws<<<blocks, threads, smem, st>>>(h->d_cfg, h->d_mesh, task, h->d_counter, nw, nc, (int)n_slots, parts);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}

int launch_lines_ws(pnjl_handle* h, long long n_lines, const double* muq, const double* xi, const int* tidx, int n_T,
                    const double* T, double* rec, cudaStream_t st, int mode) {
    WsTask t;
    std::memset(&t, 0, sizeof(t));
    t.mode = mode == 0 ? 0 : (mode == 1 ? 2 : 3); t.n_tasks = n_lines; t.muq_MeV = muq; t.xi = xi; t.table_idx = tidx; t.n_T = n_T; t.T_MeV = T; t.records = rec;
    t.out_index = h->out_index_next;
    return launch_ws(h, t, st);
}

int launch_points_ws(pnjl_handle* h, long long n, const double* T, const double* mu, const double* xi, int seed_mode,
                     int n_seeds, const double* seeds, double* rec, cudaStream_t st) {
    WsTask t;
    std::memset(&t, 0, sizeof(t));
    t.mode = 1; t.n_tasks = n; t.T_fm = T; t.mu_fm = mu; t.xi = xi; t.seed_mode = seed_mode; t.n_seeds = n_seeds; t.seeds = seeds;
    t.records = rec;
    return launch_ws(h, t, st);
}

// Shared-memory layout and launch constants of the line-march kernels; returns the dynamic shared memory size in bytes.
static size_t march_layout(const pnjl_handle* h, int parts, MarchConst& mc) {
    const int n_mesh = 3 * h->n_nodes + 2 * h->host_cfg.n_iso;
    std::memset(&mc, 0, sizeof(mc));
    mc.cfg = h->d_cfg;
    mc.parts = parts;
    // teams of consecutive warps; in team t the leader is the warp whose scheduler (w & 3) has the fewest leaders so far
    {
        const int n_teams = kMarchWarps / parts;
        int leaders_on[4] = {0, 0, 0, 0};
        for (int w = 0; w < kMarchWarps; ++w) { mc.team_of[w] = -1; mc.role_of[w] = 0; }
        for (int t = 0; t < n_teams; ++t) {
            int lead = t * parts;
            for (int w = t * parts; w < (t + 1) * parts; ++w)
                if (leaders_on[w & 3] < leaders_on[lead & 3]) lead = w;
            ++leaders_on[lead & 3];
            int role = 1;
            for (int w = t * parts; w < (t + 1) * parts; ++w) {
                mc.team_of[w] = (signed char)t;
                mc.role_of[w] = (signed char)(w == lead ? 0 : role++);
            }
        }
    }
    mc.stage0 = (n_mesh + 1) & ~1;
    mc.lean0 = mc.stage0 + kMarchWarps * kStageDoubles;
    mc.team0 = mc.lean0 + kMarchWarps * LW_END;
    mc.cmd0 = mc.team0 + kMarchWarps * kBufStride;
    mc.red0 = mc.cmd0 + kMarchWarps * kCmdDoubles;
    mc.n = h->n_nodes; mc.n_iso = h->host_cfg.n_iso;
    mc.p2max = h->host_cfg.p2max; mc.pc2max = h->host_cfg.pc2max;
    mc.sp = h->host_cfg.sp;
    mc.state0 = mc.red0 + kMarchWarps * kFJAcc * kRedStride;
    return sizeof(double) * (size_t)(mc.state0 + kMarchWarps * kStateDoubles);
}

// Launch geometry of the line-march kernel.  Team size: one warp per line while the GPU holds at least 3/4 as many lines
// as warps; below that the teams grow (a leader and 1, 3, ... followers that only sweep) as long as there are fewer than
// 1.5 teams' worth of lines per team slot and every lane keeps two nodes.  Measured on 1/2, 1/4, 1/8 shares of config 5
// (4096 / 2048 / 1024 lines on 2368 warps): 1, 2 and 4 warps per team are the fastest there; surplus lines are time-sliced.
int launch_march(pnjl_handle* h, long long n_lines, const double* muq, const double* xi, const int* tidx, int n_T,
                 const double* T, double* rec, cudaStream_t st) {
    const long long total_warps = (long long)h->sm_count * kMarchWarps;
    const bool iso = (h->iso_next || h->iso_batch) && h->host_cfg.n_iso > 0;
    h->iso_next = false;
    const int n_eff = iso ? h->host_cfg.n_iso : h->n_nodes;
    int parts = 1;
    while (parts < kMarchWarps && n_lines * 2 * parts <= 3 * total_warps && n_eff / (64 * parts) >= 2) parts *= 2;
    // more lines than four teams per SM can hold at once: five teams of three (1/8 share of config 5: 60.5 ms against 62.5)
    if (parts == 4 && n_lines > 4LL * h->sm_count) parts = 3;
    if (h->march_parts > 0) parts = h->march_parts;
    if (parts < 1 || parts > kMarchWarps) return fail(PNJL_ERR_ARG, "march_parts must be 1 .. 16");
    const long long n_teams = (long long)h->sm_count * (kMarchWarps / parts);
    int quantum = h->march_quantum > 0 ? h->march_quantum : (n_lines <= n_teams ? n_T : 32);
    if (quantum > n_T) quantum = n_T;
    const long long n_quanta = (n_T + quantum - 1) / quantum;
    MarchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n_lines = n_lines; a.muq_MeV = muq; a.xi = xi; a.table_idx = tidx; a.n_T = n_T; a.T_MeV = T; a.records = rec;
    a.out_index = h->out_index_next;
    a.capacity = n_lines * n_quanta;
    a.quantum = quantum;
    CUDA_TRY(h->march_state.reserve(sizeof(LineState) * (size_t)n_lines));
    CUDA_TRY(h->march_slots.reserve(sizeof(int) * (size_t)a.capacity));
    CUDA_TRY(h->march_counters.reserve(sizeof(unsigned long long) * 4));
    a.state = (LineState*)h->march_state.p;
    a.slots = (int*)h->march_slots.p;
    a.counters = (unsigned long long*)h->march_counters.p;
    MarchConst mc;
    const size_t smem = march_layout(h, parts, mc);
    const auto kernel = parts == 1 ? k_march<false> : k_march<true>;
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, kernel));
    if (smem > 40 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)(n_lines < h->sm_count ? n_lines : h->sm_count);
    h->stats.regs_per_thread = fa.numRegs;
    h->stats.smem_bytes = (int)smem;
    h->stats.blocks = blocks;
    h->stats.threads = 32 * kMarchWarps;
    h->stats.lanes_per_solve = 32 * parts;
    // model and launch constants go to constant memory on the launch stream (stream-ordered with the kernel; two handles with
    // DIFFERENT constants must not launch concurrently on the same device — calls on one handle are serial anyway)
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_model, &h->host_cfg.m, sizeof(Model), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_mc, &mc, sizeof(MarchConst), 0, cudaMemcpyHostToDevice, st));
    const long long n_init = a.capacity > n_lines ? a.capacity : n_lines;
    k_march_init<<<(unsigned)((n_init + 255) / 256), 256, 0, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    kernel<<<blocks, 32 * kMarchWarps, smem, st>>>(h->d_mesh, a);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 2;
    return PNJL_OK;
}

// Independent points through the line-march machinery (one team per point, no controller warps).
int launch_march_points(pnjl_handle* h, long long n, const double* T, const double* mu, const double* xi, int seed_mode,
                        int n_seeds, const double* seeds, double* rec, cudaStream_t st) {
    const long long total_warps = (long long)h->sm_count * kMarchWarps;
    int parts = 1;
    while (parts < kMarchWarps && n * 2 * parts <= 3 * total_warps && h->n_nodes / (64 * parts) >= 2) parts *= 2;
    if (h->march_parts > 0) parts = h->march_parts;
    if (parts < 1 || parts > kMarchWarps) return fail(PNJL_ERR_ARG, "march_parts must be 1 .. 16");
    MarchPointArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = n; a.T_fm = T; a.mu_fm = mu; a.xi = xi; a.seed_mode = seed_mode; a.n_seeds = n_seeds; a.seeds = seeds; a.records = rec;
    a.counter = h->d_counter;
    MarchConst mc;
    const size_t smem = march_layout(h, parts, mc);
    cudaFuncAttributes fa;
    CUDA_TRY(cudaFuncGetAttributes(&fa, k_march_points));
    if (smem > 40 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_march_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long teams_per_cta = kMarchWarps / parts;
    const long long need = (n + teams_per_cta - 1) / teams_per_cta;
    const int blocks = (int)(need < h->sm_count ? need : h->sm_count);
    h->stats.regs_per_thread = fa.numRegs;
    h->stats.smem_bytes = (int)smem;
    h->stats.blocks = blocks;
    h->stats.threads = 32 * kMarchWarps;
    h->stats.lanes_per_solve = 32 * parts;
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_model, &h->host_cfg.m, sizeof(Model), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_mc, &mc, sizeof(MarchConst), 0, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), st));
    k_march_points<<<blocks, 32 * kMarchWarps, smem, st>>>(h->d_mesh, a);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}

template <int G>
int launch_lines(pnjl_handle* h, long long n_lines, const double* muq, const double* xi, const int* tidx, int n_T,
                 const double* T, double* rec, cudaStream_t st, int mode) {
    const size_t smem = sizeof(double) * (3 * h->n_nodes + 2 * h->host_cfg.n_iso);
    int blocks, threads;
    int rc = launch_geometry(h, k_scan_lines<G>, smem, n_lines, G, &blocks, &threads);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), st));
    k_scan_lines<G><<<blocks, threads, smem, st>>>(h->d_cfg, h->d_mesh, n_lines, muq, xi, tidx, n_T, T, rec, h->d_counter, mode, h->out_index_next);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}

template <int G>
int launch_fj(pnjl_handle* h, long long n, const double* T, const double* mu, const double* xi, const double* x,
              double* FJ, cudaStream_t st, int with_thermo) {
    const size_t smem = sizeof(double) * (3 * h->n_nodes + 2 * h->host_cfg.n_iso);
    int blocks, threads;
    int rc = launch_geometry(h, k_eval_fj<G>, smem, n, G, &blocks, &threads);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned long long), st));
    k_eval_fj<G><<<blocks, threads, smem, st>>>(h->d_cfg, h->d_mesh, n, T, mu, xi, x, FJ, h->d_counter, with_thermo);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}

#define DISPATCH_G(h, call)                      \
    do {                                         \
        switch ((h)->G) {                        \
            case 8: return call<8>;              \
            case 16: return call<16>;            \
            default: return call<32>;            \
        }                                        \
    } while (0)

// The layout chosen for the next launch only (host entry points), consumed at the very start of every *_device entry so that
// an early error return cannot leak it into a later call.
int take_layout(pnjl_handle* h, bool* iso_hint = nullptr) {
    const int g = h->G_next ? h->G_next : h->G;
    h->G_next = 0;
    if (iso_hint) *iso_hint = h->iso_next;
    h->iso_next = false;
    return g;
}

// Host entry points know xi: when the whole batch is isotropic (xi == 0 everywhere) and the collapse is on, a pass sweeps
// p_num nodes only, so the layout is chosen for that mesh (8 or 16 lanes per solve) instead of the full one — the
// warp-specialised kernel would be bound by its controller warps there.
void choose_layout_for_batch(pnjl_handle* h, long long n, const double* xi, bool march_lines = false) {
    h->G_next = 0;
    h->iso_next = false;
    if (h->G_user != 0 || h->host_cfg.n_iso == 0) return;
    for (long long i = 0; i < n; ++i)
        if (xi[i] != 0.0) return;
    if (march_lines && h->schedule >= 2 && h->G == 32) {
        h->iso_next = true;      // the line-march kernel keeps its layout and only sizes its teams for p_num nodes
        return;
    }
    const int n_eff = h->host_cfg.n_iso;
    h->G_next = n_eff <= 96 ? 8 : (n_eff <= 256 ? 16 : 32);
}

int dispatch_points(pnjl_handle* h, int layout, long long n, const double* T, const double* mu, const double* xi, int seed_mode,
                    int n_seeds, const double* seeds, double* rec, cudaStream_t st) {
    switch (layout) {
        case 8: return launch_points<8>(h, n, T, mu, xi, seed_mode, n_seeds, seeds, rec, st);
        case 16: return launch_points<16>(h, n, T, mu, xi, seed_mode, n_seeds, seeds, rec, st);
        default:
            if (h->schedule == 3 || (getenv("PNJL_POINTS_MARCH") && atoi(getenv("PNJL_POINTS_MARCH")) != 0))
                return launch_march_points(h, n, T, mu, xi, seed_mode, n_seeds, seeds, rec, st);
            if (h->schedule >= 1) return launch_points_ws(h, n, T, mu, xi, seed_mode, n_seeds, seeds, rec, st);
            return launch_points<32>(h, n, T, mu, xi, seed_mode, n_seeds, seeds, rec, st);
    }
}
int dispatch_lines(pnjl_handle* h, int layout, long long n_lines, const double* muq, const double* xi, const int* tidx, int n_T,
                   const double* T, double* rec, cudaStream_t st, int mode = 0) {
    switch (layout) {
        case 8: return launch_lines<8>(h, n_lines, muq, xi, tidx, n_T, T, rec, st, mode);
        case 16: return launch_lines<16>(h, n_lines, muq, xi, tidx, n_T, T, rec, st, mode);
        default: {
            // Which organisation marches the lines (measured on cfg5 / cfg4 shares, profiles/r02_*): the line-march kernel when
            // a pass is short (all-isotropic batch: p_num nodes) or the GPU holds few lines (<= 32 per SM: multi-GPU shares of a
            // fixed grid; 2048 lines: 100 ms against 132, 4096 lines: 171 against 183, 8192 lines: 332 against 312), where the latency of a pass decides; the warp-specialised kernel when there are enough lines to hide
            // its controller step (its workers' small code stays inside the instruction cache).
            const bool iso = (h->iso_next || h->iso_batch) && h->host_cfg.n_iso > 0;
            const bool march = mode == 0 && (h->schedule == 3 || (h->schedule == 2 && (iso || n_lines <= 32LL * h->sm_count)));
            if (march) return launch_march(h, n_lines, muq, xi, tidx, n_T, T, rec, st);
            h->iso_next = false;
            if (h->schedule >= 1) return launch_lines_ws(h, n_lines, muq, xi, tidx, n_T, T, rec, st, mode);
            return launch_lines<32>(h, n_lines, muq, xi, tidx, n_T, T, rec, st, mode);
        }
    }
}
int dispatch_fj(pnjl_handle* h, long long n, const double* T, const double* mu, const double* xi, const double* x,
                double* FJ, cudaStream_t st, int with_thermo = 0) {
    switch (h->G) {
        case 8: return launch_fj<8>(h, n, T, mu, xi, x, FJ, st, with_thermo);
        case 16: return launch_fj<16>(h, n, T, mu, xi, x, FJ, st, with_thermo);
        default: return launch_fj<32>(h, n, T, mu, xi, x, FJ, st, with_thermo);
    }
}

// If `host_ptr` is page-locked host memory that the device can address (cudaHostAlloc / cudaHostRegister under UVA), return
// its device alias: the solve kernels then write their 256-byte records straight into the caller's buffer over PCIe while
// they run (2.1 GB in 340 ms on cfg5 is 6 GB/s) and the device->host copy after the kernel disappears.  nullptr otherwise
// (pageable memory: staged through a device buffer and copied).  PNJL_ZERO_COPY=0 turns it off.
double* device_alias_of_pinned(void* host_ptr) {
    static const bool enabled = !(getenv("PNJL_ZERO_COPY") && atoi(getenv("PNJL_ZERO_COPY")) == 0);
    if (!enabled) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return (double*)a.devicePointer;
}

}  // namespace

extern "C" {

int pnjl_abi_version(void) { return PNJL_ABI_VERSION; }

const char* pnjl_last_error(void) { return g_last_error.c_str(); }

int64_t pnjl_sizeof_config(void) { return (int64_t)sizeof(pnjl_config); }
int64_t pnjl_sizeof_boundary(void) { return (int64_t)sizeof(pnjl_boundary); }
int64_t pnjl_sizeof_stats(void) { return (int64_t)sizeof(pnjl_stats); }
int64_t pnjl_config_field_offset(const char* field) {
    if (!field) return -1;
    const std::string f(field);
#define PNJL_FIELD(name) if (f == #name) return (int64_t)offsetof(pnjl_config, name);
    PNJL_FIELD(hbarc) PNJL_FIELD(Lambda) PNJL_FIELD(m_ud0) PNJL_FIELD(m_s0) PNJL_FIELD(G) PNJL_FIELD(K) PNJL_FIELD(T0)
    PNJL_FIELD(a0) PNJL_FIELD(a1) PNJL_FIELD(a2) PNJL_FIELD(b3) PNJL_FIELD(rho0) PNJL_FIELD(Nc) PNJL_FIELD(p_num) PNJL_FIELD(t_num)
    PNJL_FIELD(p_nodes) PNJL_FIELD(p_w) PNJL_FIELD(c_nodes) PNJL_FIELD(c_w) PNJL_FIELD(xtol) PNJL_FIELD(ftol)
    PNJL_FIELD(residual_norm_max) PNJL_FIELD(phi_tol) PNJL_FIELD(max_iter) PNJL_FIELD(tr_fallback) PNJL_FIELD(auto_multiseed_fallback)
    PNJL_FIELD(omega_tie_rel) PNJL_FIELD(device) PNJL_FIELD(lanes_per_solve) PNJL_FIELD(predict_tol) PNJL_FIELD(isospin_symmetric)
    PNJL_FIELD(schedule) PNJL_FIELD(isotropic_collapse)
#undef PNJL_FIELD
    return -1;
}

void pnjl_default_config(pnjl_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->hbarc = 197.327;
    c->Lambda = 602.3 / c->hbarc;
    c->m_ud0 = 5.5 / c->hbarc;
    c->m_s0 = 140.7 / c->hbarc;
    c->G = 1.835 / (c->Lambda * c->Lambda);
    c->K = 12.36 / std::pow(c->Lambda, 5);
    c->T0 = 210.0 / c->hbarc;
    c->a0 = 3.51; c->a1 = -2.47; c->a2 = 15.2; c->b3 = -1.75;
    c->rho0 = 0.16;
    c->Nc = 3;
    c->p_num = 64; c->t_num = 8;
    c->xtol = 1e-9; c->ftol = 1e-9; c->residual_norm_max = 1e-6; c->phi_tol = 1e-8;
    c->max_iter = 1000;
    c->tr_fallback = 1; c->auto_multiseed_fallback = 1;
    c->omega_tie_rel = 1e-12;
    c->device = -1;
    c->lanes_per_solve = 0;
    c->predict_tol = 1e-4;
    c->isospin_symmetric = 1;
    c->schedule = 0;
    c->isotropic_collapse = 1;
}

int pnjl_gauleg(double a, double b, int32_t n, double* nodes, double* weights) {
    if (n <= 0) return fail(PNJL_ERR_ARG, "gauleg: n must be > 0");        // GaussLegendre.jl:96-98
    if (!(a < b)) return fail(PNJL_ERR_ARG, "gauleg: need a < b");          // GaussLegendre.jl:100-102
    std::vector<double> x, w;
    gausslegendre_std(n, x, w);
    const double scale = (b - a) / 2.0, shift = (b + a) / 2.0;
    for (int i = 0; i < n; ++i) {
        nodes[i] = scale * x[i] + shift;
        weights[i] = scale * w[i];
    }
    return PNJL_OK;
}

int pnjl_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return fail(PNJL_ERR_ARG, "null buffer");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? PNJL_ERR_NOMEM : PNJL_ERR_CUDA, cudaGetErrorString(e)); }
    return PNJL_OK;
}
int pnjl_free_pinned(void* p) {
    if (!p) return PNJL_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PNJL_ERR_CUDA, cudaGetErrorString(e)); }
    return PNJL_OK;
}

int pnjl_create(const pnjl_config* c, pnjl_handle** out) {
    if (!c || !out) return fail(PNJL_ERR_ARG, "null argument");
    if (c->p_num <= 0 || c->t_num <= 0 || (long long)c->p_num * c->t_num > 2048)
        return fail(PNJL_ERR_ARG, "p_num * t_num must be in 1..2048");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PNJL_ERR_CUDA, std::string("no CUDA device: this library has no CPU fallback (") +
                                       cudaGetErrorString(e) + ")");
    int dev = c->device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(PNJL_ERR_ARG, "device ordinal out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(PNJL_ERR_CUDA, "device is not sm_100 (B200); the library is built for sm_100a only");
    DeviceGuard guard(dev);

    pnjl_handle* h = new pnjl_handle();
    h->device = dev;
    h->sm_count = prop.multiProcessorCount;
    h->n_nodes = c->p_num * c->t_num;
    std::memset(&h->stats, 0, sizeof(h->stats));
    h->G_user = c->lanes_per_solve;
    if (c->lanes_per_solve == 8 || c->lanes_per_solve == 16 || c->lanes_per_solve == 32) h->G = c->lanes_per_solve;
    else if (c->lanes_per_solve == 0) h->G = h->n_nodes <= 96 ? 8 : (h->n_nodes <= 256 ? 16 : 32);
    else { delete h; return fail(PNJL_ERR_ARG, "lanes_per_solve must be 0, 8, 16 or 32"); }

    // quadrature rule: caller's nodes (Julia: FastGaussQuadrature) or the library's own
    std::vector<double> pn(c->p_num), pw(c->p_num), cn(c->t_num), cw(c->t_num);
    if (c->p_nodes && c->p_w) {
        std::memcpy(pn.data(), c->p_nodes, sizeof(double) * c->p_num);
        std::memcpy(pw.data(), c->p_w, sizeof(double) * c->p_num);
    } else {
        pnjl_gauleg(0.0, 10.0, c->p_num, pn.data(), pw.data());  // Integrals.jl:75-80
    }
    if (c->c_nodes && c->c_w) {
        std::memcpy(cn.data(), c->c_nodes, sizeof(double) * c->t_num);
        std::memcpy(cw.data(), c->c_w, sizeof(double) * c->t_num);
    } else {
        pnjl_gauleg(0.0, 1.0, c->t_num, cn.data(), cw.data());   // Integrals.jl:67-73
    }
    // mesh in build_nodes order (Integrals.jl:87-96): column-major (p_num, t_num), p fastest
    const int n_iso = c->isotropic_collapse ? c->p_num : 0;
    std::vector<double> mesh(3 * (size_t)h->n_nodes + 2 * (size_t)n_iso);
    const double two_pi = 2 * kPi;
    for (int j = 0; j < c->t_num; ++j)
        for (int i = 0; i < c->p_num; ++i) {
            const int k = j * c->p_num + i;
            const double p = pn[i], t = cn[j];
            mesh[k] = p * p;
            mesh[h->n_nodes + k] = (p * t) * (p * t);
            mesh[2 * h->n_nodes + k] = (pw[i] * (cw[j] * 2.0)) * (p * p) / (two_pi * two_pi);
        }

    for (int i = 0; i < n_iso; ++i) {
        double csum = 0.0;
        for (int j = 0; j < c->t_num; ++j) csum += mesh[2 * h->n_nodes + j * c->p_num + i];
        mesh[3 * h->n_nodes + i] = pn[i] * pn[i];
        mesh[3 * h->n_nodes + n_iso + i] = csum;
    }

    DeviceConfig& dc = h->host_cfg;
    std::memset(&dc, 0, sizeof(dc));
    dc.n_iso = n_iso;
    dc.m.hbarc = c->hbarc; dc.m.Lambda = c->Lambda; dc.m.m_ud0 = c->m_ud0; dc.m.m_s0 = c->m_s0; dc.m.G = c->G; dc.m.K = c->K;
    dc.m.T0 = c->T0; dc.m.a0 = c->a0; dc.m.a1 = c->a1; dc.m.a2 = c->a2; dc.m.b3 = c->b3; dc.m.rho0 = c->rho0; dc.m.Nc = c->Nc;
    dc.sp.xtol = c->xtol; dc.sp.ftol = c->ftol; dc.sp.residual_norm_max = c->residual_norm_max; dc.sp.phi_tol = c->phi_tol;
    dc.sp.omega_tie_rel = c->omega_tie_rel; dc.sp.max_iter = c->max_iter; dc.sp.tr_fallback = c->tr_fallback;
    dc.sp.auto_multiseed_fallback = c->auto_multiseed_fallback;
    dc.sp.isospin = c->isospin_symmetric;
    dc.sp.predict_tol = getenv("PNJL_PREDICT_TOL") ? atof(getenv("PNJL_PREDICT_TOL")) : c->predict_tol;
    {
        // launch shape: lock-stepped 512-thread CTAs (one per SM) for the 32-lane layout, small CTAs otherwise;
        // PNJL_BLOCK_THREADS / PNJL_LOCKSTEP override for experiments
        const char* eb = getenv("PNJL_BLOCK_THREADS");
        const char* el = getenv("PNJL_LOCKSTEP");
        dc.lockstep = el ? atoi(el) : (h->G == 32 ? 1 : 0);   // 0 off, 1 align loop entry, 3 align entry and exit
        h->block_threads = eb ? atoi(eb) : (h->G == 32 ? 512 : 128);
        h->block_threads_forced = eb != nullptr;
        if (h->block_threads < 32 || h->block_threads > 512 || (h->block_threads & 31)) h->block_threads = 128;
        // internal numbering: 1 = warp-specialised, 0 = one warp per line; PNJL_SCHEDULE overrides for experiments
        const char* es = getenv("PNJL_SCHEDULE");
        h->schedule = es ? atoi(es) : (c->schedule == 1 ? 0 : (c->schedule == 2 ? 1 : (c->schedule == 3 ? 3 : 2)));
        if (getenv("PNJL_MARCH_PARTS")) h->march_parts = atoi(getenv("PNJL_MARCH_PARTS"));
        if (getenv("PNJL_MARCH_Q")) h->march_quantum = atoi(getenv("PNJL_MARCH_Q"));
        if (getenv("PNJL_WS_WORKERS")) h->ws_workers = atoi(getenv("PNJL_WS_WORKERS"));
        if (getenv("PNJL_WS_CTRL")) h->ws_ctrl_warps = atoi(getenv("PNJL_WS_CTRL"));
        if (getenv("PNJL_WS_SPW")) h->ws_spw = atoi(getenv("PNJL_WS_SPW"));
        if (getenv("PNJL_WS_SLOTS")) h->ws_slots = atoi(getenv("PNJL_WS_SLOTS"));
        if (h->ws_spw < 1 || h->ws_spw > 8) h->ws_spw = 4;
        if (h->ws_workers < 1 || h->ws_workers > kWsMaxWarps - 1) h->ws_workers = kWsMaxWarps - 1;
        if (h->ws_ctrl_warps < 1 || h->ws_workers + h->ws_ctrl_warps > kWsMaxWarps) h->ws_ctrl_warps = kWsMaxWarps - h->ws_workers;
    }
    dc.n_nodes = h->n_nodes;
    dc.p2max = 0.0;
    dc.pc2max = 0.0;
    for (int k = 0; k < h->n_nodes; ++k) {
        dc.p2max = std::max(dc.p2max, mesh[k]);
        dc.pc2max = std::max(dc.pc2max, mesh[h->n_nodes + k]);
    }
    dc.pt.n_tables = 0;

#define CREATE_TRY(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            pnjl_destroy(h);                                                                          \
            return fail(PNJL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
        }                                                                                             \
    } while (0)
#ifdef PNJL_PROFILE_PHASES
    CREATE_TRY(cudaMalloc(&dc.dbg, 8 * sizeof(unsigned long long)));
    CREATE_TRY(cudaMemset(dc.dbg, 0, 8 * sizeof(unsigned long long)));
#endif
    CREATE_TRY(cudaMalloc(&h->d_cfg, sizeof(DeviceConfig)));
    CREATE_TRY(cudaMalloc(&h->d_mesh, sizeof(double) * mesh.size()));
    CREATE_TRY(cudaMalloc(&h->d_counter, sizeof(unsigned long long)));
    CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&h->ev0));
    CREATE_TRY(cudaEventCreate(&h->ev1));
    CREATE_TRY(cudaMemcpy(h->d_cfg, &dc, sizeof(dc), cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMemcpy(h->d_mesh, mesh.data(), sizeof(double) * mesh.size(), cudaMemcpyHostToDevice));
#undef CREATE_TRY
    *out = h;
    {
        // default one-loop rule = DEFAULT_MOMENTUM_NODES / WEIGHTS = gauleg(0, 10, 64)   (GaussLegendre.jl:121,170)
        std::vector<double> x(64), w(64);
        int rc = pnjl_gauleg(0.0, 10.0, 64, x.data(), w.data());
        if (!rc) rc = pnjl_set_oneloop_rule(h, 64, x.data(), w.data());
        if (rc) { *out = nullptr; pnjl_destroy(h); return rc; }
    }
    return PNJL_OK;
}

void pnjl_destroy(pnjl_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
#ifdef PNJL_PROFILE_PHASES
    if (h->host_cfg.dbg) {
        unsigned long long d[8];
        cudaDeviceSynchronize();
        cudaMemcpy(d, h->host_cfg.dbg, sizeof(d), cudaMemcpyDeviceToHost);
        if (d[2])
            fprintf(stderr, "[phases] passes %llu: loop %.0f cyc/pass; barrier wait %.0f cyc/pass; scalar/idle phase %.0f cyc/pass; "
                            "controller: %llu requests, scalar %.0f cyc/request, wait %.0f cyc/request\n", d[2],
                    (double)d[0] / d[2], (double)d[1] / d[2], (double)d[3] / d[2], d[5], d[5] ? (double)d[4] / d[5] : 0.0,
                    d[5] ? (double)d[6] / d[5] : 0.0);
        cudaFree(h->host_cfg.dbg);
    }
#endif
    h->in_T.release(); h->in_mu.release(); h->in_xi.release(); h->in_seeds.release(); h->in_idx.release();
    h->in_x.release(); h->out_rec.release(); h->out_aux.release();
    h->march_state.release(); h->march_slots.release(); h->march_counters.release();
    h->pt_start.release(); h->pt_tcep.release(); h->pt_T.release(); h->pt_mu.release();
    for (auto& b : h->in_c) b.release();
    if (h->d_rule) cudaFree(h->d_rule);
    if (h->d_cfg) cudaFree(h->d_cfg);
    if (h->d_mesh) cudaFree(h->d_mesh);
    if (h->d_counter) cudaFree(h->d_counter);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int pnjl_set_boundaries(pnjl_handle* h, int32_t n_tables, const pnjl_boundary* tables) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n_tables < 0 || (n_tables > 0 && !tables)) return fail(PNJL_ERR_ARG, "bad boundary table list");
    // validate and flatten into local buffers first: the handle's tables are replaced only when everything is in order
    std::vector<int> start(n_tables + 1, 0);
    std::vector<double> tcep((size_t)(n_tables > 0 ? n_tables : 1), 0.0), Ts, ms;
    for (int t = 0; t < n_tables; ++t) {
        const int n = tables[t].n;
        if (n < 0 || (n > 0 && (!tables[t].T_MeV || !tables[t].mu_c_MeV))) return fail(PNJL_ERR_ARG, "bad boundary table");
        tcep[t] = tables[t].T_CEP_MeV;
        for (int i = 0; i < n; ++i) {
            if (i > 0 && !(tables[t].T_MeV[i] >= tables[t].T_MeV[i - 1])) return fail(PNJL_ERR_ARG, "boundary table must be sorted by T");
            Ts.push_back(tables[t].T_MeV[i]);
            ms.push_back(tables[t].mu_c_MeV[i]);
        }
        start[t + 1] = (int)Ts.size();
    }
    DeviceGuard guard(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const size_t rows = Ts.size() ? Ts.size() : 1;
    DevBuf nb_start, nb_tcep, nb_T, nb_mu;
    cudaError_t e = nb_start.reserve(sizeof(int) * start.size());
    if (e == cudaSuccess) e = nb_tcep.reserve(sizeof(double) * tcep.size());
    if (e == cudaSuccess) e = nb_T.reserve(sizeof(double) * rows);
    if (e == cudaSuccess) e = nb_mu.reserve(sizeof(double) * rows);
    if (e == cudaSuccess) e = cudaMemcpy(nb_start.p, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(nb_tcep.p, tcep.data(), sizeof(double) * tcep.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && Ts.size()) e = cudaMemcpy(nb_T.p, Ts.data(), sizeof(double) * Ts.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && ms.size()) e = cudaMemcpy(nb_mu.p, ms.data(), sizeof(double) * ms.size(), cudaMemcpyHostToDevice);
    DeviceConfig next = h->host_cfg;
    next.pt.n_tables = n_tables;
    next.pt.start = (const int*)nb_start.p;
    next.pt.T_CEP = (const double*)nb_tcep.p;
    next.pt.T = (const double*)nb_T.p;
    next.pt.mu = (const double*)nb_mu.p;
    if (e == cudaSuccess) e = cudaMemcpy(h->d_cfg, &next, sizeof(DeviceConfig), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError();
        nb_start.release(); nb_tcep.release(); nb_T.release(); nb_mu.release();
        return fail(e == cudaErrorMemoryAllocation ? PNJL_ERR_NOMEM : PNJL_ERR_CUDA, std::string("pnjl_set_boundaries: ") + cudaGetErrorString(e));
    }
    // commit
    h->pt_start.release(); h->pt_tcep.release(); h->pt_T.release(); h->pt_mu.release();
    h->pt_start = nb_start; h->pt_tcep = nb_tcep; h->pt_T = nb_T; h->pt_mu = nb_mu;
    h->host_cfg = next;
    return PNJL_OK;
}

int pnjl_set_option(pnjl_handle* h, const char* key, int64_t value) {
    if (!h || !key) return fail(PNJL_ERR_ARG, "null argument");
    const std::string k(key);
    if (k == "schedule") {
        if (value < 0 || value > 3) return fail(PNJL_ERR_ARG, "schedule: 0 automatic, 1 one warp per line, 2 worker/controller warps, 3 line march");
        h->schedule = value == 1 ? 0 : (value == 2 ? 1 : (value == 3 ? 3 : 2));
    } else if (k == "march_parts") {
        if (value < 0 || value > kMarchWarps) return fail(PNJL_ERR_ARG, "march_parts must be 0 (automatic) .. 16");
        h->march_parts = (int)value;
    } else if (k == "march_quantum") {
        if (value < 0 || value > (1 << 30)) return fail(PNJL_ERR_ARG, "march_quantum out of range");
        h->march_quantum = (int)value;
    } else if (k == "isotropic_batch") {
        h->iso_batch = value != 0;
    } else {
        return fail(PNJL_ERR_ARG, "unknown option: " + k);
    }
    return PNJL_OK;
}

int pnjl_solve_points_device(pnjl_handle* h, int64_t n, const double* d_T, const double* d_mu, const double* d_xi,
                             int32_t seed_mode, int32_t n_seeds, const double* d_seeds, double* d_records, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    const int layout = take_layout(h);
    if (n < 0) return fail(PNJL_ERR_ARG, "n < 0");
    if (seed_mode < 0 || seed_mode > 2) return fail(PNJL_ERR_ARG, "bad seed_mode");
    if (seed_mode == PNJL_SEED_EXPLICIT && (!d_seeds || n_seeds < 1 || n_seeds > 6))
        return fail(PNJL_ERR_ARG, "explicit seeds need 1..6 seeds per point");
    h->stats.kernel_launches = 0;
    if (n == 0) return PNJL_OK;
    if (!d_T || !d_mu || !d_xi || !d_records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    return dispatch_points(h, layout, n, d_T, d_mu, d_xi, seed_mode, n_seeds, d_seeds, d_records, (cudaStream_t)stream);
}

int pnjl_scan_lines_device(pnjl_handle* h, int64_t n_lines, const double* d_muq, const double* d_xi,
                           const int32_t* d_tidx, int32_t n_T, const double* d_T, double* d_records, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    bool iso_hint = false;
    const int layout = take_layout(h, &iso_hint);
    if (n_lines < 0 || n_T < 0) return fail(PNJL_ERR_ARG, "negative size");
    h->stats.kernel_launches = 0;
    if (n_lines == 0 || n_T == 0) return PNJL_OK;
    if (!d_muq || !d_xi || !d_T || !d_records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    h->iso_next = iso_hint;       // consumed by dispatch_lines / launch_march below
    return dispatch_lines(h, layout, n_lines, d_muq, d_xi, d_tidx, n_T, d_T, d_records, (cudaStream_t)stream);
}

int pnjl_solve_points_host(pnjl_handle* h, int64_t n, const double* T, const double* mu, const double* xi,
                           int32_t seed_mode, int32_t n_seeds, const double* seeds, double* records) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n < 0) return fail(PNJL_ERR_ARG, "n < 0");
    if (n == 0) { h->stats.kernel_launches = 0; return PNJL_OK; }
    if (!T || !mu || !xi || !records) return fail(PNJL_ERR_ARG, "null buffer");
    if (seed_mode < 0 || seed_mode > 2) return fail(PNJL_ERR_ARG, "bad seed_mode");
    if (seed_mode == PNJL_SEED_EXPLICIT && (!seeds || n_seeds < 1 || n_seeds > 6))
        return fail(PNJL_ERR_ARG, "explicit seeds need 1..6 seeds per point");     // before anything is allocated or enqueued
    DeviceGuard guard(h->device);
    const size_t nb = sizeof(double) * (size_t)n;
    CUDA_TRY(h->in_T.reserve(nb));
    CUDA_TRY(h->in_mu.reserve(nb));
    CUDA_TRY(h->in_xi.reserve(nb));
    double* alias = device_alias_of_pinned(records);
    if (!alias) CUDA_TRY(h->out_rec.reserve(nb * PNJL_REC_DOUBLES));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, T, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_mu.p, mu, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_xi.p, xi, nb, cudaMemcpyHostToDevice, st));
    const double* d_seeds = nullptr;
    if (seed_mode == PNJL_SEED_EXPLICIT) {
        CUDA_TRY(h->in_seeds.reserve(nb * 5 * n_seeds));
        CUDA_TRY(cudaMemcpyAsync(h->in_seeds.p, seeds, nb * 5 * n_seeds, cudaMemcpyHostToDevice, st));
        d_seeds = (const double*)h->in_seeds.p;
    }
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    choose_layout_for_batch(h, n, xi);
    int rc = pnjl_solve_points_device(h, n, (const double*)h->in_T.p, (const double*)h->in_mu.p, (const double*)h->in_xi.p,
                                      seed_mode, n_seeds, d_seeds, alias ? alias : (double*)h->out_rec.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    if (!alias) CUDA_TRY(cudaMemcpyAsync(records, h->out_rec.p, nb * PNJL_REC_DOUBLES, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}

static int scan_lines_host_impl(pnjl_handle* h, int64_t n_lines, const double* muq, const double* xi, const int32_t* tidx,
                                int32_t n_T, const double* T_MeV, double* records, bool keep_device_copy) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n_lines < 0 || n_T < 0) return fail(PNJL_ERR_ARG, "negative size");
    if (n_lines == 0 || n_T == 0) { h->stats.kernel_launches = 0; return PNJL_OK; }
    if (!muq || !xi || !T_MeV || !records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    const size_t nl = sizeof(double) * (size_t)n_lines;
    const size_t nrec = sizeof(double) * (size_t)n_lines * n_T * PNJL_REC_DOUBLES;
    CUDA_TRY(h->in_mu.reserve(nl));
    CUDA_TRY(h->in_xi.reserve(nl));
    CUDA_TRY(h->in_T.reserve(sizeof(double) * n_T));
    CUDA_TRY(h->in_idx.reserve(sizeof(int32_t) * (size_t)n_lines));
    double* alias = keep_device_copy ? nullptr : device_alias_of_pinned(records);
    if (!alias) CUDA_TRY(h->out_rec.reserve(nrec));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_mu.p, muq, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_xi.p, xi, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, T_MeV, sizeof(double) * n_T, cudaMemcpyHostToDevice, st));
    const int32_t* d_idx = nullptr;
    if (tidx) {
        CUDA_TRY(cudaMemcpyAsync(h->in_idx.p, tidx, sizeof(int32_t) * (size_t)n_lines, cudaMemcpyHostToDevice, st));
        d_idx = (const int32_t*)h->in_idx.p;
    }
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    choose_layout_for_batch(h, n_lines, xi, true);
    int rc = pnjl_scan_lines_device(h, n_lines, (const double*)h->in_mu.p, (const double*)h->in_xi.p, d_idx, n_T,
                                    (const double*)h->in_T.p, alias ? alias : (double*)h->out_rec.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    if (!alias) CUDA_TRY(cudaMemcpyAsync(records, h->out_rec.p, nrec, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}
int pnjl_scan_lines_host(pnjl_handle* h, int64_t n_lines, const double* muq, const double* xi, const int32_t* tidx,
                         int32_t n_T, const double* T_MeV, double* records) {
    return scan_lines_host_impl(h, n_lines, muq, xi, tidx, n_T, T_MeV, records, false);
}

int pnjl_set_oneloop_rule(pnjl_handle* h, int32_t n_nodes, const double* nodes, const double* weights) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n_nodes < 1 || n_nodes > PNJL_MAX_ONELOOP_NODES) return fail(PNJL_ERR_ARG, "one-loop rule: need 1 <= n_nodes <= 512");
    if (!nodes || !weights) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    std::vector<double> r(2 * (size_t)n_nodes);
    for (int i = 0; i < n_nodes; ++i) {
        r[i] = nodes[i] * nodes[i];
        r[n_nodes + i] = weights[i] * (nodes[i] * nodes[i]);     // weight_p * node_p^2   OneLoopIntegrals.jl:540
    }
    if (h->stream) CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->d_rule) cudaFree(h->d_rule);
    h->d_rule = nullptr;
    CUDA_TRY(cudaMalloc(&h->d_rule, sizeof(double) * r.size()));
    CUDA_TRY(cudaMemcpy(h->d_rule, r.data(), sizeof(double) * r.size(), cudaMemcpyHostToDevice));
    h->n_rule = n_nodes;
    return PNJL_OK;
}

namespace {
int launch_couplings(pnjl_handle* h, long long n, const CouplingsIn& in, double* d_aux, cudaStream_t st) {
    const int threads = 128;
    long long blocks = (n + threads - 1) / threads;
    const long long cap = (long long)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_couplings<<<(int)blocks, threads, sizeof(double) * 2 * h->n_rule, st>>>(h->d_cfg, h->d_rule, h->n_rule, n, in, d_aux);
    CUDA_TRY(cudaGetLastError());
    h->stats.kernel_launches += 1;
    return PNJL_OK;
}
}  // namespace

int pnjl_effective_couplings_device(pnjl_handle* h, int64_t n, const double* d_records, double* d_aux, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n < 0) return fail(PNJL_ERR_ARG, "negative size");
    h->stats.kernel_launches = 0;
    if (n == 0) return PNJL_OK;
    if (!d_records || !d_aux) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    CouplingsIn in{d_records + PNJL_REC_T, d_records + PNJL_REC_MU, d_records + PNJL_REC_MASS, d_records + PNJL_REC_MASS + 2,
                   d_records + PNJL_REC_X + 3, d_records + PNJL_REC_X + 4, PNJL_REC_DOUBLES};
    return launch_couplings(h, n, in, d_aux, (cudaStream_t)stream);
}

int pnjl_effective_couplings_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* m_u,
                                  const double* m_s, const double* Phi, const double* Phibar, double* aux) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n < 0) return fail(PNJL_ERR_ARG, "negative size");
    h->stats.kernel_launches = 0;
    if (n == 0) return PNJL_OK;
    if (!T_fm || !mu_fm || !m_u || !m_s || !Phi || !Phibar || !aux) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    const double* src[6] = {T_fm, mu_fm, m_u, m_s, Phi, Phibar};
    const size_t nb = sizeof(double) * (size_t)n;
    cudaStream_t st = h->stream;
    for (int q = 0; q < 6; ++q) {
        CUDA_TRY(h->in_c[q].reserve(nb));
        CUDA_TRY(cudaMemcpyAsync(h->in_c[q].p, src[q], nb, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(h->out_aux.reserve(nb * PNJL_AUX_DOUBLES));
    CouplingsIn in{(const double*)h->in_c[0].p, (const double*)h->in_c[1].p, (const double*)h->in_c[2].p,
                   (const double*)h->in_c[3].p, (const double*)h->in_c[4].p, (const double*)h->in_c[5].p, 1};
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    int rc = launch_couplings(h, n, in, (double*)h->out_aux.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(aux, h->out_aux.p, nb * PNJL_AUX_DOUBLES, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}

int pnjl_scan_lines_couplings_host(pnjl_handle* h, int64_t n_lines, const double* muq, const double* xi, const int32_t* tidx,
                                   int32_t n_T, const double* T_MeV, double* records, double* aux) {
    if (!aux) return fail(PNJL_ERR_ARG, "null buffer");
    // the scan itself (records stay resident in out_rec), then the couplings of every record on the same stream
    int rc = scan_lines_host_impl(h, n_lines, muq, xi, tidx, n_T, T_MeV, records, true);
    if (rc || n_lines == 0 || n_T == 0) return rc;
    DeviceGuard guard(h->device);
    const long long n = (long long)n_lines * n_T;
    const size_t nb = sizeof(double) * (size_t)n * PNJL_AUX_DOUBLES;
    CUDA_TRY(h->out_aux.reserve(nb));
    const long long scan_launches = h->stats.kernel_launches;
    rc = pnjl_effective_couplings_device(h, n, (const double*)h->out_rec.p, (double*)h->out_aux.p, h->stream);
    if (rc) return rc;
    h->stats.kernel_launches += scan_launches;
    CUDA_TRY(cudaMemcpyAsync(aux, h->out_aux.p, nb, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PNJL_OK;
}

int pnjl_scan_lines_device_indexed(pnjl_handle* h, int64_t n_lines, const double* d_muq, const double* d_xi,
                                   const int32_t* d_tidx, int32_t n_T, const double* d_T, double* d_records_base,
                                   const int64_t* d_out_index, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    h->out_index_next = (const long long*)d_out_index;
    const int rc = pnjl_scan_lines_device(h, n_lines, d_muq, d_xi, d_tidx, n_T, d_T, d_records_base, stream);
    h->out_index_next = nullptr;
    return rc;
}

// ---- peer-visible result buffers (multi-GPU: every rank's kernel stores its records straight into rank 0's buffer) ----
int pnjl_ipc_alloc(uint64_t bytes, void** dptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!dptr || !handle) return fail(PNJL_ERR_ARG, "null buffer");
    *dptr = nullptr;
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(e == cudaErrorMemoryAllocation ? PNJL_ERR_NOMEM : PNJL_ERR_CUDA, cudaGetErrorString(e)); }
    cudaIpcMemHandle_t hd;
    e = cudaIpcGetMemHandle(&hd, *dptr);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(*dptr); *dptr = nullptr; return fail(PNJL_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    std::memcpy(handle, &hd, 64);
    return PNJL_OK;
}
int pnjl_ipc_open(const unsigned char handle[64], void** dptr) {
    if (!dptr || !handle) return fail(PNJL_ERR_ARG, "null buffer");
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); *dptr = nullptr; return fail(PNJL_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
    return PNJL_OK;
}
int pnjl_ipc_close(void* dptr) {
    if (!dptr) return PNJL_OK;
    cudaError_t e = cudaIpcCloseMemHandle(dptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PNJL_ERR_CUDA, cudaGetErrorString(e)); }
    return PNJL_OK;
}
int pnjl_ipc_free(void* dptr) {
    if (!dptr) return PNJL_OK;
    cudaError_t e = cudaFree(dptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PNJL_ERR_CUDA, cudaGetErrorString(e)); }
    return PNJL_OK;
}

int pnjl_tmu_scan_device(pnjl_handle* h, int64_t n_lines, const double* d_T, const double* d_xi, const int32_t* d_tidx,
                         int32_t n_mu, const double* d_mu, double* d_records, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    const int layout = take_layout(h);
    if (n_lines < 0 || n_mu < 0) return fail(PNJL_ERR_ARG, "negative size");
    h->stats.kernel_launches = 0;
    if (n_lines == 0 || n_mu == 0) return PNJL_OK;
    if (!d_T || !d_xi || !d_mu || !d_records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    return dispatch_lines(h, layout, n_lines, d_T, d_xi, d_tidx, n_mu, d_mu, d_records, (cudaStream_t)stream, 1);
}

int pnjl_tmu_scan_host(pnjl_handle* h, int64_t n_lines, const double* T_MeV, const double* xi, const int32_t* tidx,
                       int32_t n_mu, const double* mu_MeV, double* records) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n_lines < 0 || n_mu < 0) return fail(PNJL_ERR_ARG, "negative size");
    if (n_lines == 0 || n_mu == 0) { h->stats.kernel_launches = 0; return PNJL_OK; }
    if (!T_MeV || !xi || !mu_MeV || !records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    const size_t nl = sizeof(double) * (size_t)n_lines;
    const size_t nrec = sizeof(double) * (size_t)n_lines * n_mu * PNJL_REC_DOUBLES;
    CUDA_TRY(h->in_mu.reserve(nl));
    CUDA_TRY(h->in_xi.reserve(nl));
    CUDA_TRY(h->in_T.reserve(sizeof(double) * n_mu));
    CUDA_TRY(h->in_idx.reserve(sizeof(int32_t) * (size_t)n_lines));
    double* alias = device_alias_of_pinned(records);
    if (!alias) CUDA_TRY(h->out_rec.reserve(nrec));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_mu.p, T_MeV, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_xi.p, xi, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, mu_MeV, sizeof(double) * n_mu, cudaMemcpyHostToDevice, st));
    const int32_t* d_idx = nullptr;
    if (tidx) {
        CUDA_TRY(cudaMemcpyAsync(h->in_idx.p, tidx, sizeof(int32_t) * (size_t)n_lines, cudaMemcpyHostToDevice, st));
        d_idx = (const int32_t*)h->in_idx.p;
    }
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    choose_layout_for_batch(h, n_lines, xi);
    int rc = pnjl_tmu_scan_device(h, n_lines, (const double*)h->in_mu.p, (const double*)h->in_xi.p, d_idx, n_mu,
                                  (const double*)h->in_T.p, alias ? alias : (double*)h->out_rec.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    if (!alias) CUDA_TRY(cudaMemcpyAsync(records, h->out_rec.p, nrec, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}

int pnjl_dual_branch_device(pnjl_handle* h, int64_t n_lines, const double* d_T, const double* d_xi, int32_t n_mu,
                            const double* d_mu, double* d_records, void* stream) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    const int layout = take_layout(h);
    if (n_lines < 0 || n_mu < 0) return fail(PNJL_ERR_ARG, "negative size");
    h->stats.kernel_launches = 0;
    if (n_lines == 0 || n_mu == 0) return PNJL_OK;
    if (!d_T || !d_xi || !d_mu || !d_records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    return dispatch_lines(h, layout, 2 * n_lines, d_T, d_xi, nullptr, n_mu, d_mu, d_records, (cudaStream_t)stream, 2);
}
int pnjl_dual_branch_host(pnjl_handle* h, int64_t n_lines, const double* T_MeV, const double* xi, int32_t n_mu,
                          const double* mu_MeV, double* records) {
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n_lines < 0 || n_mu < 0) return fail(PNJL_ERR_ARG, "negative size");
    if (n_lines == 0 || n_mu == 0) { h->stats.kernel_launches = 0; return PNJL_OK; }
    if (!T_MeV || !xi || !mu_MeV || !records) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    const size_t nl = sizeof(double) * (size_t)n_lines;
    const size_t nrec = sizeof(double) * (size_t)n_lines * 2 * n_mu * PNJL_REC_DOUBLES;
    CUDA_TRY(h->in_mu.reserve(nl));
    CUDA_TRY(h->in_xi.reserve(nl));
    CUDA_TRY(h->in_T.reserve(sizeof(double) * n_mu));
    double* alias = device_alias_of_pinned(records);
    if (!alias) CUDA_TRY(h->out_rec.reserve(nrec));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_mu.p, T_MeV, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_xi.p, xi, nl, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, mu_MeV, sizeof(double) * n_mu, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    choose_layout_for_batch(h, n_lines, xi);
    int rc = pnjl_dual_branch_device(h, n_lines, (const double*)h->in_mu.p, (const double*)h->in_xi.p, n_mu,
                                     (const double*)h->in_T.p, alias ? alias : (double*)h->out_rec.p, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    if (!alias) CUDA_TRY(cudaMemcpyAsync(records, h->out_rec.p, nrec, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}

static int eval_state_impl(pnjl_handle* h, int64_t n, const double* T, const double* mu, const double* xi, const double* x,
                           double* FJ, int with_thermo) {
    const size_t stride = with_thermo == 2 ? PNJL_DERIV_DOUBLES : (with_thermo ? PNJL_STATE_DOUBLES : 30);
    if (!h) return fail(PNJL_ERR_ARG, "null handle");
    if (n <= 0) return n == 0 ? PNJL_OK : fail(PNJL_ERR_ARG, "n < 0");
    if (!T || !mu || !xi || !x || !FJ) return fail(PNJL_ERR_ARG, "null buffer");
    DeviceGuard guard(h->device);
    const size_t nb = sizeof(double) * (size_t)n;
    CUDA_TRY(h->in_T.reserve(nb));
    CUDA_TRY(h->in_mu.reserve(nb));
    CUDA_TRY(h->in_xi.reserve(nb));
    CUDA_TRY(h->in_x.reserve(nb * 5));
    CUDA_TRY(h->out_rec.reserve(nb * stride));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, T, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_mu.p, mu, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_xi.p, xi, nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->in_x.p, x, nb * 5, cudaMemcpyHostToDevice, st));
    h->stats.kernel_launches = 0;
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    int rc = dispatch_fj(h, n, (const double*)h->in_T.p, (const double*)h->in_mu.p, (const double*)h->in_xi.p,
                         (const double*)h->in_x.p, (double*)h->out_rec.p, st, with_thermo);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    CUDA_TRY(cudaMemcpyAsync(FJ, h->out_rec.p, nb * stride, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->stats.kernel_ms = ms;
    return PNJL_OK;
}
int pnjl_eval_fj_host(pnjl_handle* h, int64_t n, const double* T, const double* mu, const double* xi, const double* x,
                      double* FJ) {
    return eval_state_impl(h, n, T, mu, xi, x, FJ, 0);
}
int pnjl_eval_state_host(pnjl_handle* h, int64_t n, const double* T, const double* mu, const double* xi, const double* x,
                         double* out) {
    return eval_state_impl(h, n, T, mu, xi, x, out, 1);
}
int pnjl_eval_derivs_host(pnjl_handle* h, int64_t n, const double* T, const double* mu, const double* xi, const double* x,
                          double* out) {
    return eval_state_impl(h, n, T, mu, xi, x, out, 2);
}

int pnjl_selftest_math(pnjl_handle* h, int64_t n, const double* x, int32_t which, double* out) {
    if (!h || !x || !out || n <= 0 || which < 0 || which > 2) return fail(PNJL_ERR_ARG, "bad argument");
    DeviceGuard guard(h->device);
    const size_t nb = sizeof(double) * (size_t)n;
    CUDA_TRY(h->in_T.reserve(nb));
    CUDA_TRY(h->out_rec.reserve(nb));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->in_T.p, x, nb, cudaMemcpyHostToDevice, st));
    k_selftest_math<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, (const double*)h->in_T.p, which, (double*)h->out_rec.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, h->out_rec.p, nb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PNJL_OK;
}

int pnjl_get_stats(pnjl_handle* h, pnjl_stats* out) {
    if (!h || !out) return fail(PNJL_ERR_ARG, "null argument");
    *out = h->stats;
    return PNJL_OK;
}

int pnjl_measure_fp64_peak(pnjl_handle* h, double seconds, double* tflops_burst, double* tflops_sustained) {
    if (!h || !tflops_burst) return fail(PNJL_ERR_ARG, "null argument");
    DeviceGuard guard(h->device);
    const int threads = 256, blocks = h->sm_count * 8, iters = 4096;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double) * threads * blocks));
    cudaStream_t st = h->stream;
    const double fmas = (double)threads * blocks * (double)iters * 16.0 * 8.0;
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        CUDA_TRY(cudaEventRecord(h->ev0, st));
        k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters, 1.0 + rep);
        CUDA_TRY(cudaEventRecord(h->ev1, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        if (rep >= 1 && ms < best_ms) best_ms = ms;
    }
    *tflops_burst = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    if (tflops_sustained) {
        // back-to-back launches for `seconds` (power-capped clocks), one event pair around the whole train
        int reps = (int)(seconds * 1e3 / best_ms) + 1;
        if (reps > 100000) reps = 100000;
        CUDA_TRY(cudaEventRecord(h->ev0, st));
        for (int rep = 0; rep < reps; ++rep) k_dfma_peak<<<blocks, threads, 0, st>>>(d_out, iters, 2.0 + rep);
        CUDA_TRY(cudaEventRecord(h->ev1, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        *tflops_sustained = 2.0 * fmas * reps / (ms * 1e-3) / 1e12;
    }
    cudaFree(d_out);
    return PNJL_OK;
}

}  // extern "C"
