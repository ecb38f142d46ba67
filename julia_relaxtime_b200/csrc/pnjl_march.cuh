// pnjl_march.cuh — the line-march kernel k_march (included by pnjl_kernels.cu inside namespace pnjl).
//
// One warp (or a team of 2/4/8/16 warps when a GPU holds fewer lines than warps) owns a (xi, mu) line and runs its whole
// continuity march — seed, Newton iterations, final thermodynamics, record — in its own registers
// (run_gap_transport_scan.jl:407-443, ImplicitSolver.jl:211-328 through Solver<WarpEval>):
//   * a quadrature pass is the same paired FP64 loop as in k_solve_ws (ws_worker_pass: lanes stride the mesh in shared
//     memory, transposed shuffle butterfly), the 20 sums are broadcast to all lanes through the warp's shared-memory
//     scratch line, and the closed-form finish + the 5x5 elimination run in the same warp right away — no hand-over, no
//     controller warps, no polling.  The finish is kept small: the vacuum integrals of the u and the s flavour are
//     evaluated by different lanes, a Jacobian pass skips the logarithm of the Polyakov potential, and every libm
//     fall-back sits in an out-of-line cold function;
//   * lines are time-sliced: a global ticket queue hands out (line, next T index) quanta of `quantum` points; a warp that
//     finishes a quantum parks the line's tracker state (64 bytes) in global memory, re-queues the line and takes the oldest
//     waiting one.  All lines therefore advance at the same pace on ALL SMs (no per-SM imbalance, the tail is one quantum),
//     and any number of lines per GPU keeps every warp busy;
//   * few lines per GPU (multi-GPU slabs of a fixed grid, the CEP window of config 4): the warps of a team split every pass
//     by nodes, add their partial sums in a fixed order through shared memory behind one named barrier per pass, and run the
//     scalar part redundantly, so the latency of a pass drops with the team size;
//   * a record leaves as one coalesced 256-byte row written by the 32 lanes of the owning warp.
#pragma once

#ifndef PNJL_MARCH_LEAN
#define PNJL_MARCH_LEAN 1
#endif

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
    int parts;                        // warps per team
};

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

__device__ __noinline__ void vacuum_terms_cold(double Lam, double M, double& I0, double& I1, double& I2) {
    vacuum_terms_t<false>(Lam, M, I0, I1, I2);
}
__device__ __noinline__ void polyakov_cold(const Model& m, double T, double iT, double P, double Pb, UTerms& u) {
    polyakov_eval<false, true>(m, T, iT, P, Pb, u);
}

struct WarpEval {
    const DeviceConfig* cfg;
    const Model* m;
    MeshView mv;
    double* stage;       // this warp's staging line in shared memory [kStageDoubles]
    double* team_buf;    // this team's reduced partial sums [2][parts][kBufStride] (double-buffered by pass parity)
    int lane, part, parts, bar_id, parity, isospin;

    __device__ __forceinline__ void team_sync() const {
        if (parts > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * parts) : "memory");
        else __syncwarp();
    }

    // One quadrature pass of kind `type` at (T, mu, xi, x): every lane of every warp of the team ends up with the N sums.
    // Partial sums of the team's warps are added in part order, so the result does not depend on which warp is faster.
    template <int N>
    __device__ __forceinline__ bool pass(int type, double T, double mu, double xi, const double x[5], double (&acc)[N]) {
        __syncwarp();
        if (lane == 0) {
            stage[0] = T; stage[1] = mu; stage[2] = xi;
#pragma unroll
            for (int i = 0; i < 5; ++i) stage[3 + i] = x[i];
        }
        __syncwarp();
        double* base = team_buf + (size_t)parity * parts * kBufStride;
        ws_worker_pass(cfg, mv, stage, type, lane, part, parts, base + part * kBufStride);
        team_sync();
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const double2 v = *reinterpret_cast<const double2*>(base + i);
            acc[i] = v.x;
            if (i + 1 < N) acc[i + 1] = v.y;
        }
        for (int q = 1; q < parts; ++q) {
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                const double2 v = *reinterpret_cast<const double2*>(base + q * kBufStride + i);
                acc[i] += v.x;
                if (i + 1 < N) acc[i + 1] += v.y;
            }
        }
        const bool fast = base[20] != 0.0;
        parity ^= 1;
        return fast;
    }

    // Closed-form ingredients, one flavour per lane: even lanes take the u flavour (M[0]; M[1] is taken from it when the
    // masses coincide bitwise), odd lanes the s flavour; the d flavour falls back to the generic evaluation when it differs.
    __device__ __forceinline__ void vacuum_all(const PointCtx& c, double I0v[3], double I1v[3], double I2v[3]) const {
#if PNJL_MARCH_LEAN
        const bool odd = (lane & 1) != 0;
        const double Mf = odd ? c.M[2] : c.M[0];
        double I0, I1, I2;
        if (vacuum_tame(m->Lambda, Mf)) vacuum_terms_t<true>(m->Lambda, Mf, I0, I1, I2);
        else vacuum_terms_cold(m->Lambda, Mf, I0, I1, I2);
        I0v[0] = __shfl_sync(0xffffffffu, I0, 0); I0v[2] = __shfl_sync(0xffffffffu, I0, 1);
        I1v[0] = __shfl_sync(0xffffffffu, I1, 0); I1v[2] = __shfl_sync(0xffffffffu, I1, 1);
        I2v[0] = __shfl_sync(0xffffffffu, I2, 0); I2v[2] = __shfl_sync(0xffffffffu, I2, 1);
        if (c.M[1] == c.M[0]) { I0v[1] = I0v[0]; I1v[1] = I1v[0]; I2v[1] = I2v[0]; }
        else vacuum_terms_cold(m->Lambda, c.M[1], I0v[1], I1v[1], I2v[1]);
#else
        double I0 = 0, I1 = 0, I2 = 0;
        for (int i = 0; i < 3; ++i) {
            if (!(i == 1 && c.M[1] == c.M[0])) vacuum_terms(m->Lambda, c.M[i], I0, I1, I2);
            I0v[i] = I0; I1v[i] = I1; I2v[i] = I2;
        }
#endif
    }
    template <bool WITH_VALUE>
    __device__ __forceinline__ void polyakov(const PointCtx& c, UTerms& u) const {
#if PNJL_MARCH_LEAN
        if (polyakov_tame(c.Phi, c.Phib)) polyakov_eval<true, WITH_VALUE>(*m, c.T, c.invT, c.Phi, c.Phib, u);
        else polyakov_cold(*m, c.T, c.invT, c.Phi, c.Phib, u);
#else
        polyakov_U(*m, c.T, c.invT, c.Phi, c.Phib, u);
#endif
    }

    __device__ __noinline__ void fj(double T, double mu, double xi, const double x[5], double F[5], double J[25]) {
        double acc[kFJAcc];
        const bool fast = pass<kFJAcc>(WS_FJ, T, mu, xi, x, acc);
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        polyakov<false>(c, u);
        finish_fj_pre(*m, c, x, acc, I1v, I2v, u, F, J, fast);
    }
    __device__ __noinline__ bool fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
        double acc[kFJAcc];
        const bool fast = pass<kFJAcc>(WS_FJ, T, mu, xi, x, acc);
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        polyakov<false>(c, u);
        double J[25], b[5];
        finish_fj_pre(*m, c, x, acc, I1v, I2v, u, F, J, fast);
#pragma unroll
        for (int i = 0; i < 5; ++i) b[i] = F[i];
        const bool ok = lu_solve5_regs(J, b, p);
#pragma unroll
        for (int i = 0; i < 5; ++i) p[i] = -p[i];
        return ok;
    }
    __device__ __noinline__ bool f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        {
            const double k2max = mv.p2max + (xi > 0.0 ? xi * mv.pc2max : 0.0);
            if (!fast_path_ok(c.T, c.mu, c.Phi, c.Phib, k2max, c.M2)) return false;   // uniform over the team
        }
        double acc[kFtAcc + kThAcc + 1];
        pass<kFtAcc + kThAcc + 1>(WS_FT, T, mu, xi, x, acc);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        polyakov<true>(c, u);
        finish_f_pre(*m, c, x, acc, I1v, u, F);
        finish_thermo_pre(*m, c, x, acc + kFtAcc, I0v, u, th);
        return true;
    }
    __device__ __noinline__ void thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
        double acc[kThAcc];
        pass<kThAcc>(WS_TH, T, mu, xi, x, acc);
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double I0v[3], I1v[3], I2v[3];
        vacuum_all(c, I0v, I1v, I2v);
        UTerms u;
        polyakov<true>(c, u);
        finish_thermo_pre(*m, c, x, acc, I0v, u, th);
    }
};

// Record writer of a team: warp 0 of the team stages the row in its shared-memory line and the 32 lanes store it.
struct MarchSink {
    WarpEval* ev;
    double* base;
    double xi;
    __device__ __forceinline__ void operator()(int it, const PointRes& r, double T_fm, double mu_fm, int n_fj, int n_th,
                                               int n_ft) {
        if (ev->part != 0) return;
        double rec[PNJL_REC_DOUBLES];
        fill_record(r, T_fm, mu_fm, xi, n_fj, n_th, n_ft, rec);
        __syncwarp();
        if (ev->lane == 0) {
#pragma unroll
            for (int q = 0; q < PNJL_REC_DOUBLES; q += 2) *reinterpret_cast<double2*>(ev->stage + q) = make_double2(rec[q], rec[q + 1]);
        }
        __syncwarp();
        base[(long long)PNJL_REC_DOUBLES * it + ev->lane] = ev->stage[ev->lane];
        __syncwarp();
    }
};

// Dynamic shared memory: mesh [3 n + 2 n_iso] | staging lines [16][32] | team buffers [2][16][24] | popped line per team [16]
__global__ void __launch_bounds__(32 * kMarchWarps, 1) k_march(const DeviceConfig* __restrict__ cfg, const double* __restrict__ g_mesh,
                                                               MarchArgs a) {
    extern __shared__ double s_dyn[];
    const int n = cfg->n_nodes;
    const int n_mesh = 3 * n + 2 * cfg->n_iso;
    double* s_mesh = s_dyn;
    double* s_stage = s_dyn + ((n_mesh + 1) & ~1);
    double* s_team = s_stage + kMarchWarps * kStageDoubles;
    int* s_line = reinterpret_cast<int*>(s_team + 2 * kMarchWarps * kBufStride);
    for (int i = threadIdx.x; i < n_mesh; i += blockDim.x) s_mesh[i] = g_mesh[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int parts = a.parts;
    const int team = warp / parts, part = warp - team * parts;
    WarpEval ev;
    ev.cfg = cfg;
    ev.m = &cfg->m;
    ev.mv.p2 = s_mesh; ev.mv.pc2 = s_mesh + n; ev.mv.coef = s_mesh + 2 * n; ev.mv.n = n;
    ev.mv.p2max = cfg->p2max; ev.mv.pc2max = cfg->pc2max;
    ev.mv.p2_iso = s_mesh + 3 * n; ev.mv.coef_iso = s_mesh + 3 * n + cfg->n_iso; ev.mv.n_iso = cfg->n_iso;
    ev.stage = s_stage + warp * kStageDoubles;
    ev.team_buf = s_team + (size_t)team * parts * 2 * kBufStride;
    ev.lane = lane; ev.part = part; ev.parts = parts; ev.bar_id = 1 + team; ev.parity = 0;
    ev.isospin = cfg->sp.isospin;
    Solver<WarpEval> sv(cfg->m, cfg->sp, ev);
    volatile int* slots = a.slots;
    volatile unsigned long long* done = a.counters + 2;
    for (;;) {
        // ---- pop: the team leader takes a ticket and waits for the line that goes with it (or for the end of the scan) ----
        int line = -1;
        if (part == 0) {
            if (lane == 0) {
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
            if (parts > 1 && lane == 0) s_line[team] = line;
        }
        if (parts > 1) {
            ev.team_sync();
            line = s_line[team];
            ev.team_sync();          // everybody has read the slot before the leader can overwrite it
        }
        if (line < 0) break;
        // ---- run one time slice of the line ----
        LineState st;
        {
            const LineState* g = a.state + line;
#pragma unroll
            for (int q = 0; q < 5; ++q) st.prev[q] = __ldcg(&g->prev[q]);
            st.it_next = __ldcg(&g->it_next); st.prev_phase = __ldcg(&g->prev_phase);
            st.has_prev = __ldcg(&g->has_prev); st.its_hint = __ldcg(&g->its_hint);
        }
        const long long row = a.out_index ? a.out_index[line] : (long long)line;
        MarchSink sink{&ev, a.records + (long long)PNJL_REC_DOUBLES * a.n_T * row, a.xi[line]};
        scan_line_slice(sv, &cfg->pt, a.table_idx ? a.table_idx[line] : -1, a.muq_MeV[line], a.xi[line], a.n_T, a.T_MeV, st,
                        a.quantum, sink);
        // ---- park the line (or retire it) ----
        if (part == 0 && lane == 0) {
            if (st.it_next < a.n_T) {
                LineState* g = a.state + line;
#pragma unroll
                for (int q = 0; q < 5; ++q) __stcg(&g->prev[q], st.prev[q]);
                __stcg(&g->it_next, st.it_next); __stcg(&g->prev_phase, st.prev_phase);
                __stcg(&g->has_prev, st.has_prev); __stcg(&g->its_hint, st.its_hint);
                __threadfence();
                const unsigned long long t = atomicAdd(a.counters + 1, 1ULL);
                if ((long long)t < a.capacity) slots[t] = line;
            } else {
                __threadfence();
                atomicAdd(a.counters + 2, 1ULL);
            }
        }
    }
}
