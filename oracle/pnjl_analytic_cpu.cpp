// pnjl_analytic_cpu.cpp — TEST / BASELINE INFRASTRUCTURE: a host (g++, OpenMP) build of the product's own math and solver headers.
//
// Compiles julia_relaxtime_b200/csrc/pnjl_math.cuh + pnjl_solver.cuh + pnjl_lean.cuh without CUDA and with a sequential
// evaluation policy.  Two uses, both outside the product path:
//   * tests (tests/hostsim/hostsim.py loads it): the analytic derivatives, the solve cascade, the resumable line march and the
//     lane-parallel finish that the CUDA kernels run are checked against the AD oracle on a machine without a GPU;
//   * bench.py's cpu_baseline leg: the "analytic-Jacobian CPU baseline" SURVEY.md §8d names — the same algorithm as the GPU
//     kernel (closed-form Jacobian, isospin shortcut, fused final pass) on all host cores, next to the AD oracle that mirrors
//     the reference's ForwardDiff arithmetic.
// It is NOT part of libpnjl_b200.so and nothing in the package julia_relaxtime_b200/ loads it: the product has no CPU path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../julia_relaxtime_b200/csrc/pnjl_solver.cuh"
#include "../julia_relaxtime_b200/csrc/pnjl_lean.cuh"

using namespace pnjl;

namespace {

struct HostMesh {
    int n;
    std::vector<double> p2, pc2, coef;
    std::vector<double> p2_iso, coef_iso;   // isotropic collapse (empty: disabled), as in pnjl_create
};

struct HostEval {
    const Model* m;
    const HostMesh* mesh;
    int isospin;
    MeshView view() const {
        MeshView mv;
        mv.p2 = mesh->p2.data(); mv.pc2 = mesh->pc2.data(); mv.coef = mesh->coef.data(); mv.n = mesh->n;
        mv.p2max = 0; mv.pc2max = 0;
        for (int k = 0; k < mesh->n; ++k) {
            mv.p2max = mesh->p2[k] > mv.p2max ? mesh->p2[k] : mv.p2max;
            mv.pc2max = mesh->pc2[k] > mv.pc2max ? mesh->pc2[k] : mv.pc2max;
        }
        mv.p2_iso = mesh->p2_iso.data(); mv.coef_iso = mesh->coef_iso.data(); mv.n_iso = (int)mesh->p2_iso.size();
        return mv;
    }
    void fj(double T, double mu, double xi, const double x[5], double F[5], double J[25]) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kFJAcc];
        const bool fast = fj_partial(*m, isospin != 0, c, x, view(), 0, 1, acc);
        finish_fj(*m, c, x, acc, F, J, fast);
    }
    bool fj_step(double T, double mu, double xi, const double x[5], double F[5], double p[5]) {
        double J[25], b[5];
        fj(T, mu, xi, x, F, J);
        for (int i = 0; i < 5; ++i) b[i] = F[i];
        const bool ok = lu_solve5_regs(J, b, p);
        for (int i = 0; i < 5; ++i) p[i] = -p[i];
        return ok;
    }
    bool f_thermo(double T, double mu, double xi, const double x[5], double F[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double facc[kFtAcc], tacc[kThAcc];
        if (!ft_partial(*m, isospin != 0, c, x, view(), 0, 1, facc, tacc)) return false;
        finish_f(*m, c, x, facc, F);
        finish_thermo(*m, c, x, tacc, th);
        return true;
    }
    void thermo(double T, double mu, double xi, const double x[5], Thermo& th) {
        PointCtx c;
        make_ctx(*m, T, mu, xi, x, c);
        double acc[kThAcc];
        thermo_partial(*m, isospin != 0, c, x, view(), 0, 1, acc);
        finish_thermo(*m, c, x, acc, th);
    }
};

Model model_of(const pnjl_config* c) {
    Model m;
    m.hbarc = c->hbarc; m.Lambda = c->Lambda; m.m_ud0 = c->m_ud0; m.m_s0 = c->m_s0; m.G = c->G; m.K = c->K;
    m.T0 = c->T0; m.a0 = c->a0; m.a1 = c->a1; m.a2 = c->a2; m.b3 = c->b3; m.rho0 = c->rho0; m.Nc = c->Nc;
    return m;
}
SolverParams params_of(const pnjl_config* c) {
    SolverParams s;
    s.xtol = c->xtol; s.ftol = c->ftol; s.residual_norm_max = c->residual_norm_max; s.phi_tol = c->phi_tol;
    s.omega_tie_rel = c->omega_tie_rel; s.max_iter = c->max_iter; s.tr_fallback = c->tr_fallback;
    s.auto_multiseed_fallback = c->auto_multiseed_fallback; s.isospin = c->isospin_symmetric; s.predict_tol = c->predict_tol;
    return s;
}
HostMesh mesh_of(const pnjl_config* c) {
    HostMesh h;
    h.n = c->p_num * c->t_num;
    h.p2.resize(h.n); h.pc2.resize(h.n); h.coef.resize(h.n);
    const double two_pi = 2 * kPi;
    for (int j = 0; j < c->t_num; ++j)
        for (int i = 0; i < c->p_num; ++i) {
            const int k = j * c->p_num + i;
            const double p = c->p_nodes[i], t = c->c_nodes[j];
            h.p2[k] = p * p;
            h.pc2[k] = (p * t) * (p * t);
            h.coef[k] = (c->p_w[i] * (c->c_w[j] * 2.0)) * (p * p) / (two_pi * two_pi);
        }
    if (c->isotropic_collapse) {
        for (int i = 0; i < c->p_num; ++i) {
            double csum = 0.0;
            for (int j = 0; j < c->t_num; ++j) csum += h.coef[j * c->p_num + i];
            h.p2_iso.push_back(c->p_nodes[i] * c->p_nodes[i]);
            h.coef_iso.push_back(csum);
        }
    }
    return h;
}

}  // namespace

extern "C" {

void hostsim_fj(const pnjl_config* c, const double* x, double T, double mu, double xi, double* F, double* J) {
    Model m = model_of(c);
    HostMesh mesh = mesh_of(c);
    HostEval ev{&m, &mesh, c->isospin_symmetric};
    ev.fj(T, mu, xi, x, F, J);
}

// The lane-parallel finish + elimination of the line-march kernel (csrc/pnjl_lean.cuh), emulated lane by lane: phases run for
// all 32 lanes one after the other, as the __syncwarp()s order them on the GPU.  out: F[5], p[5] (Newton direction), ok.
int hostsim_lean_fj_step(const pnjl_config* c, const double* x, double T, double mu, double xi, double* F, double* pdir) {
    Model m = model_of(c);
    HostMesh mesh = mesh_of(c);
    HostEval ev{&m, &mesh, c->isospin_symmetric};
    PointCtx ctx;
    make_ctx(m, T, mu, xi, x, ctx);
    double W[LW_END] = {0};
    double acc[kFJAcc];
    const bool fast = fj_partial(m, c->isospin_symmetric != 0, ctx, x, ev.view(), 0, 1, acc);
    for (int i = 0; i < kFJAcc; ++i) W[LW_S + i] = acc[i];
    LeanConst k;
    lean_consts(m, ctx.T, ctx.invT, k);
    // phase A
    for (int lane = 0; lane < 32; ++lane) {
        if (lane < 3) { if (!lean_flavour_fj(lane, k, ctx.M[lane], ctx.M2[lane], fast, W)) return -1; }
        else if (lane < 12) lean_dtable((lane - 3) / 3, (lane - 3) % 3, k, x, W);
        if (lane == 0) {
            if (!polyakov_tame(x[3], x[4])) return -2;
            UTerms u;
            polyakov_eval<true, false>(m, ctx.T, ctx.invT, x[3], x[4], u);
            W[LW_U + 0] = u.U_P; W[LW_U + 1] = u.U_Pb; W[LW_U + 2] = u.U_PP; W[LW_U + 3] = u.U_PPb; W[LW_U + 4] = u.U_PbPb;
            for (int q = 0; q < 5; ++q) W[LW_X + q] = x[q];
        }
    }
    // phase B
    double a[32] = {0};
    for (int lane = 0; lane < 30; ++lane) a[lane] = lean_aug_entry(lane / 6, lane % 6, k, W, ACC_GP, ACC_GPB);
    for (int lane = 0; lane < 30; ++lane) W[LW_AUG + lane] = a[lane];
    for (int i = 0; i < 5; ++i) F[i] = W[LW_AUG + 6 * i + 5];
    // phase C
    bool ok = true;
    for (int step = 0; step < 5; ++step) {
        double nxt[32], inv = 0;
        for (int lane = 0; lane < 30; ++lane) { bool okl = ok; nxt[lane] = lean_lu_step(step, lane / 6, lane % 6, W, a[lane], inv, okl); if (lane == 0) ok = okl; }
        for (int lane = 0; lane < 30; ++lane) { a[lane] = nxt[lane]; W[LW_AUG + lane] = nxt[lane]; }
        W[LW_INV + step] = inv;
    }
    double y[5];
    lean_backsub(W, y);
    for (int i = 0; i < 5; ++i) pdir[i] = -y[i];
    return ok ? 1 : 0;
}

// The redundant-per-lane version of the same step (finish_fj + lu_solve5_regs), for comparison.
int hostsim_fj_step(const pnjl_config* c, const double* x, double T, double mu, double xi, double* F, double* pdir) {
    Model m = model_of(c);
    HostMesh mesh = mesh_of(c);
    HostEval ev{&m, &mesh, c->isospin_symmetric};
    return ev.fj_step(T, mu, xi, x, F, pdir) ? 1 : 0;
}

// The derivative pass (dtheta_node / finish_dtheta of csrc/pnjl_math.cuh) at one state: out16 as pnjl_eval_derivs_host's tail.
void hostsim_derivs(const pnjl_config* c, const double* x, double T, double mu, double xi, double* out16) {
    Model m = model_of(c);
    HostMesh mesh = mesh_of(c);
    HostEval ev{&m, &mesh, c->isospin_symmetric};
    PointCtx ctx;
    make_ctx(m, T, mu, xi, x, ctx);
    double acc[kFJAcc], dacc[kDtAcc], F[5], J[25];
    const bool fast = fj_partial(m, c->isospin_symmetric != 0, ctx, x, ev.view(), 0, 1, acc);
    finish_fj(m, ctx, x, acc, F, J, fast);
    dtheta_partial(ctx, ev.view(), 0, 1, dacc);
    finish_dtheta(m, ctx, x, dacc, acc[ACC_GP], acc[ACC_GPB], out16);
}

void hostsim_thermo(const pnjl_config* c, const double* x, double T, double mu, double xi, double* out17) {
    Model m = model_of(c);
    HostMesh mesh = mesh_of(c);
    HostEval ev{&m, &mesh, c->isospin_symmetric};
    Thermo th;
    ev.thermo(T, mu, xi, x, th);
    out17[0] = th.omega; out17[1] = th.pressure; out17[2] = th.rho_norm; out17[3] = th.entropy; out17[4] = th.energy;
    for (int i = 0; i < 3; ++i) { out17[5 + i] = th.rho[i]; out17[8 + i] = th.M[i]; out17[11 + i] = th.nq[i]; out17[14 + i] = th.nqb[i]; }
}

void hostsim_solve_points(const pnjl_config* c, int64_t n, const double* T, const double* mu, const double* xi,
                          int32_t seed_mode, int32_t n_seeds, const double* seeds, double* records) {
    Model m = model_of(c);
    SolverParams sp = params_of(c);
    HostMesh mesh = mesh_of(c);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        HostEval ev{&m, &mesh, c->isospin_symmetric};
        Solver<HostEval> sv(m, sp, ev);
        sv.set_point(T[i], mu[i], xi[i]);
        PointRes r;
        if (seed_mode == PNJL_SEED_EXPLICIT && n_seeds == 1) sv.solve_with_fallback(seeds + 5 * i, r);
        else if (seed_mode == PNJL_SEED_EXPLICIT) sv.solve_multi(seeds + 5 * n_seeds * i, n_seeds, r);
        else if (seed_mode == PNJL_SEED_AUTO) { double x0[5]; default_seed(2, T[i], mu[i], x0); sv.solve_with_fallback(x0, r); }
        else sv.solve_multi(nullptr, 6, r);
        fill_record(r, T[i], mu[i], xi[i], sv.n_fj, sv.n_th, sv.n_ft, records + PNJL_REC_DOUBLES * i);
    }
}

struct Sink {
    double* base;
    double xi;
    void operator()(int it, const PointRes& r, double T_fm, double mu_fm, int n_fj, int n_th, int n_ft) {
        fill_record(r, T_fm, mu_fm, xi, n_fj, n_th, n_ft, base + PNJL_REC_DOUBLES * it);
    }
};

void hostsim_scan_lines_mode(const pnjl_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                             const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                             const pnjl_boundary* tables, double* records, int32_t mode);

void hostsim_scan_lines(const pnjl_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                        const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                        const pnjl_boundary* tables, double* records) {
    hostsim_scan_lines_mode(c, n_lines, muq_MeV, xi, table_idx, n_T, T_MeV, n_tables, tables, records, 0);
}

// mode 0: (xi, muq) lines marching T; mode 1: TmuScan lines, i.e. muq_MeV = per-line T_MeV and T_MeV = the mu grid
void hostsim_scan_lines_mode(const pnjl_config* c, int64_t n_lines, const double* muq_MeV, const double* xi,
                             const int32_t* table_idx, int32_t n_T, const double* T_MeV, int32_t n_tables,
                             const pnjl_boundary* tables, double* records, int32_t mode) {
    Model m = model_of(c);
    SolverParams sp = params_of(c);
    HostMesh mesh = mesh_of(c);
    std::vector<int> pt_start(n_tables + 1, 0);
    std::vector<double> pt_tcep(n_tables > 0 ? n_tables : 1), pt_T, pt_mu;
    for (int t = 0; t < n_tables; ++t) {
        pt_tcep[t] = tables[t].T_CEP_MeV;
        for (int i = 0; i < tables[t].n; ++i) { pt_T.push_back(tables[t].T_MeV[i]); pt_mu.push_back(tables[t].mu_c_MeV[i]); }
        pt_start[t + 1] = (int)pt_T.size();
    }
    PhaseTables pts{n_tables, pt_start.data(), pt_tcep.data(), pt_T.data(), pt_mu.data()};
    const PhaseTables* pt = &pts;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t l = 0; l < n_lines; ++l) {
        HostEval ev{&m, &mesh, c->isospin_symmetric};
        Solver<HostEval> sv(m, sp, ev);
        Sink sink{records + PNJL_REC_DOUBLES * n_T * (mode == 2 ? 0 : l), xi[l]};
        if (mode == 0) scan_line(sv, pt, table_idx ? table_idx[l] : -1, muq_MeV[l], xi[l], n_T, T_MeV, sink);
        else if (mode >= 100) {
            // the resumable march of the line-march kernel, in slices of (mode - 100) points, each slice with a FRESH solver
            // object (as when another warp picks the line up)
            LineState st;
            std::memset(&st, 0, sizeof(st));
            st.prev_phase = PH_UNKNOWN;
            while (st.it_next < n_T) {
                HostEval ev2{&m, &mesh, c->isospin_symmetric};
                Solver<HostEval> sv2(m, sp, ev2);
                scan_line_slice(sv2, pt, table_idx ? table_idx[l] : -1, muq_MeV[l], xi[l], n_T, T_MeV, st, mode - 100, sink);
            }
        }
        else if (mode == 1) scan_tmu_line(sv, pt, table_idx ? table_idx[l] : -1, muq_MeV[l], xi[l], n_T, T_MeV, sink);
        else {
            // dual-branch: records [line][2][n_mu]
            for (int b = 0; b < 2; ++b) {
                Sink sb{records + PNJL_REC_DOUBLES * n_T * (2 * l + b), xi[l]};
                scan_branch_line(sv, muq_MeV[l], xi[l], n_T, T_MeV, b, sb);
            }
        }
    }
}

// build_K_data through the product's header (oneloop_A + effective_couplings): aux [n][16]
void hostsim_couplings(const pnjl_config* c, int64_t n, const double* T, const double* mu, const double* m_u, const double* m_s,
                       const double* Phi, const double* Phib, int32_t n_rule, const double* nodes, const double* weights,
                       double* aux) {
    Model m = model_of(c);
    std::vector<double> p2(n_rule), wp2(n_rule);
    for (int i = 0; i < n_rule; ++i) { p2[i] = nodes[i] * nodes[i]; wp2[i] = weights[i] * (nodes[i] * nodes[i]); }
    for (int64_t i = 0; i < n; ++i) {
        const double A_u = oneloop_A(m.Lambda, m_u[i], mu[i], T[i], Phi[i], Phib[i], n_rule, p2.data(), wp2.data());
        const double A_s = oneloop_A(m.Lambda, m_s[i], mu[i], T[i], Phi[i], Phib[i], n_rule, p2.data(), wp2.data());
        effective_couplings(m.G, m.K, m.Nc, m_u[i], m_s[i], A_u, A_s, aux + 16 * i);
    }
}

}  // extern "C"
