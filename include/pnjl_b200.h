/* pnjl_b200.h — C ABI of libpnjl_b200.so: the B200 (sm_100a) implementation of the batched PNJL
 * equilibrium solve behind Julia_RelaxTime's
 *
 *     PNJL.solve(FixedMu(), T_fm, mu_fm; xi, seed_strategy, p_num, t_num, iterations)      src/pnjl/solver/ImplicitSolver.jl:211-328
 *     PNJL.solve_multi(FixedMu(), T_fm, mu_fm; ...)                                          src/pnjl/solver/ImplicitSolver.jl:532-559
 *     the (xi, muB, T) loop of scripts/relaxtime/run_gap_transport_scan.jl                   :407-443
 *
 * The reference has no FFI on this path (Julia calling Julia); these entry points are what a Julia
 * `ccall` shim binds (see julia/PNJLB200.jl and INTEGRATION.md).  Plain pointers and sizes only.
 *
 * Conventions
 *   - All floating point is FP64.  Solver quantities are in fm^-1 units like the reference
 *     (T_fm = T_MeV / 197.327); the line-scan entry takes MeV like the script does.
 *   - Results come back as one 32-double RECORD per point (256 B; a Julia Matrix{Float64}(32, n)),
 *     field offsets PNJL_REC_* below.  Integers (iterations, status, evaluation counts) are stored
 *     as exactly-representable doubles.
 *   - Return value 0 = ok, < 0 = call-level error, text via pnjl_last_error().  Per-point trouble
 *     (no convergence, all seeds failed, non-finite seed residual) is reported in the record's
 *     status bits, never as a call error: the shim maps PNJL_ST_ALL_SEEDS_FAILED back to the
 *     reference's `error("All seeds failed ...")` (ImplicitSolver.jl:553,556) where callers rely on it.
 *   - *_host entry points take caller-owned HOST buffers and do H2D / kernels / D2H internally
 *     (what `ccall` uses).  *_device entry points take DEVICE pointers and a cudaStream_t passed as
 *     void* and only enqueue work (what the torch-based multi-GPU driver and bench.py use).
 *   - If the `records` buffer of a *_host call is page-locked host memory the device can address (cudaHostAlloc,
 *     cudaHostRegister, torch pin_memory; pnjl_alloc_pinned below), the kernels write the records straight into it while
 *     they run and no device->host copy follows (cfg5: e2e 22.3 -> 24.8 M points/s).  Pageable buffers are staged.
 *   - One handle per GPU; calls on one handle are not re-entrant.  There is no CPU fallback: every
 *     entry fails with PNJL_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef PNJL_B200_H
#define PNJL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNJL_ABI_VERSION 12

/* ---- result record layout (doubles) --------------------------------------------------------- */
#define PNJL_REC_DOUBLES 32
#define PNJL_REC_X 0           /* [5] phi_u, phi_d, phi_s, Phi, Phibar      SolverResult.x_state   ImplicitSolver.jl:182 */
#define PNJL_REC_MASS 5        /* [3] M_u, M_d, M_s (fm^-1)                 SolverResult.masses    :189 */
#define PNJL_REC_OMEGA 8       /* Omega (fm^-4)                             :184 */
#define PNJL_REC_PRESSURE 9    /* P = -Omega                                :185 */
#define PNJL_REC_RHO_NORM 10   /* sum_i rho_i / (3 rho0)                    :186 */
#define PNJL_REC_ENTROPY 11    /* s = dP/dT                                 :187 */
#define PNJL_REC_ENERGY 12     /* eps = -P + sum mu_i rho_i + T s           :188 */
#define PNJL_REC_NQ 13         /* [3] n_u, n_d, n_s        calculate_number_densities  Thermodynamics.jl:255-281 */
#define PNJL_REC_NQBAR 16      /* [3] n_ubar, n_dbar, n_sbar */
#define PNJL_REC_RESNORM 19    /* ||F||_inf of the returned nlsolve result  SolverResult.residual_norm :191 */
#define PNJL_REC_ITER 20       /* iterations of the returned nlsolve result :190 */
#define PNJL_REC_STATUS 21     /* PNJL_ST_* bits */
#define PNJL_REC_NEVAL 22      /* full Omega-gradient/Jacobian quadrature passes spent on this point (all seeds, all fallbacks) */
#define PNJL_REC_RHO 23        /* [3] rho_i = dP/dmu_i                      calculate_rho  Thermodynamics.jl:215-220 */
#define PNJL_REC_NTHERMO 26    /* thermo quadrature passes spent on this point */
#define PNJL_REC_T 27          /* echo: T_fm */
#define PNJL_REC_MU 28         /* echo: mu_fm */
#define PNJL_REC_XI 29         /* echo: xi */
#define PNJL_REC_NFUSED 30     /* fused final passes (residual F + thermo sums in one quadrature pass) spent on this point */
/* 31 reserved (zero) */

/* ---- status bits ---------------------------------------------------------------------------- */
#define PNJL_ST_CONVERGED 1          /* SolverResult.converged  ImplicitSolver.jl:287 */
#define PNJL_ST_USED_TR 2            /* returned candidate is the trust-region one  :72-101 */
#define PNJL_ST_TR_ATTEMPTED 4       /* fallback solve ran  :130-150 */
#define PNJL_ST_USED_MULTISEED 8     /* result came out of solve_multi  :532-559 */
#define PNJL_ST_SEED_SHIFT 4         /* bits 4..6: index of the chosen MultiSeed candidate */
#define PNJL_ST_SEED_MASK 0x70
#define PNJL_ST_PHASE_SWITCH 128     /* PhaseAwareContinuitySeed re-seeded at a hadron<->quark flip  SeedStrategies.jl:818-826 */
#define PNJL_ST_NONFINITE 256        /* seed residual not finite (NLsolve IsFiniteException) */
#define PNJL_ST_ALL_SEEDS_FAILED 512 /* solve_multi found no converged candidate */
#define PNJL_ST_PROMOTED 1024        /* TmuScan: residual <= 1e-4 force-marked converged   src/pnjl/scans/TmuScan.jl:422-458 */
#define PNJL_ST_REFINED 2048         /* TmuScan: re-solved from a near-converged state      TmuScan.jl:391-408 */
#define PNJL_ST_CAND_SHIFT 12        /* bits 12..13: index of the TmuScan seed candidate that succeeded   TmuScan.jl:269-300 */
#define PNJL_ST_CAND_MASK 0x3000
#define PNJL_ST_NO_RESULT 16384      /* TmuScan: every candidate failed (the reference writes an all-NaN row) */
#define PNJL_ST_MASS_INVERSION 32768 /* converged with M_s <= M_u: the high-Omega root (phi_u ~ -5.3, phi_s ~ +1.9, Omega ~ -18.1) that passes
                                        the reference's physicality filter (ImplicitSolver.jl:50-60) and that continuity seeding then
                                        follows; the reference returns it silently — here it is flagged (the values are unchanged) */

/* ---- errors --------------------------------------------------------------------------------- */
#define PNJL_OK 0
#define PNJL_ERR_ARG (-1)
#define PNJL_ERR_CUDA (-2)
#define PNJL_ERR_NOMEM (-3)

/* ---- seed modes of pnjl_solve_points_* (mirror of the reference's seed strategies) ---------- */
#define PNJL_SEED_EXPLICIT 0 /* seeds[n][n_seeds][5]; n_seeds==1: DefaultSeed(seed,seed,:hadron) + auto-MultiSeed fallback per config;
                                n_seeds>1: solve_multi over the given candidates  (ImplicitSolver.jl:539-546) */
#define PNJL_SEED_AUTO 1     /* DefaultSeed(phase_hint=:auto)   SeedStrategies.jl:193-225 */
#define PNJL_SEED_MULTI 2    /* MultiSeed(): the six built-in candidates  SeedStrategies.jl:251-284 */

typedef struct pnjl_handle pnjl_handle;

/* Model constants in fm units (src/Constants_PNJL.jl:82-103), quadrature rule and solver options. */
typedef struct pnjl_config {
    double hbarc, Lambda, m_ud0, m_s0, G, K, T0, a0, a1, a2, b3, rho0;
    int32_t Nc;
    int32_t p_num, t_num;   /* Integrals.jl:87-96; p_num * t_num <= 2048 */
    const double* p_nodes;  /* [p_num] gauleg(0, 10, p_num)   Integrals.jl:75-80   (NULL: library generates them) */
    const double* p_w;      /* [p_num] */
    const double* c_nodes;  /* [t_num] gauleg(0, 1, t_num)    Integrals.jl:67-73   (NULL: library generates them) */
    const double* c_w;      /* [t_num] raw weights; the library doubles them like theta_nodes() does */
    double xtol, ftol;              /* 1e-9, 1e-9     ImplicitSolver.jl:112 */
    double residual_norm_max;       /* 1e-6           :221 */
    double phi_tol;                 /* 1e-8           :50 */
    int32_t max_iter;               /* NLsolve `iterations`: 1000 default, 40 in run_gap_transport_scan.jl:112 */
    int32_t tr_fallback;            /* trust_region_fallback=true   :217 */
    int32_t auto_multiseed_fallback;/* auto_multiseed_fallback=true :218 */
    double omega_tie_rel;           /* MultiSeed argmin-Omega tie window (relative); ties -> lowest seed index. 1e-12 */
    int32_t device;                 /* CUDA device ordinal; -1 = current device */
    int32_t lanes_per_solve;        /* 0 = auto (8/16/32 by mesh size); else 8, 16 or 32 */
    double predict_tol;             /* a Newton pass that follows a residual <= predict_tol is run as a fused "final pass"
                                       (F + thermo sums, no Jacobian); if the solve does not stop there, J is evaluated
                                       by an ordinary pass at the same x.  Iterates are unaffected.  Default 1e-4; 0 = off */
    int32_t isospin_symmetric;      /* 1 (default): when phi_u == phi_d bitwise (always true from the built-in seeds, since
                                       mu_u = mu_d and m_u0 = m_d0 on this path) evaluate the d flavour as the u flavour and
                                       keep Newton/dogleg steps u<->d symmetric (a <= 1 ulp change of the step).  0: three
                                       independent flavours everywhere, like the reference's loop (Integrals.jl:250-257). */
    int32_t schedule;               /* kernel organisation for the 32-lane layout.
                                       0 (default): automatic.  Continuity lines go to the line-march kernel (3) when a pass is
                                       short (all-isotropic batch) or the GPU holds few lines, else to the warp-specialised
                                       kernel (2); independent points use 2.
                                       1: every warp owns a line/point, CTAs phase-aligned by named barriers.
                                       2: warp-specialised — worker warps run only quadrature loops, controller lanes own one
                                       line/point each and run the solve cascade in SIMT, passes are handed over through
                                       shared-memory mailboxes.
                                       3: line march — a warp (or a team of warps when the GPU holds few lines) keeps a line's
                                       Newton solve in its registers, lines are time-sliced through a global queue.
                                       Results agree to round-off; the 8- and 16-lane layouts always use organisation 1. */
    int32_t isotropic_collapse;     /* 1 (default): for xi == 0 exactly the integrand does not depend on cos(theta)
                                       (E = sqrt(p^2 + M^2), Integrals.jl:178-180), so a pass sums over the p_num momentum
                                       nodes with the cos(theta) weights pre-summed instead of over p_num * t_num nodes.
                                       Same sums up to the order of additions.  0: always the full mesh like the reference. */
} pnjl_config;

/* First-order phase boundary mu_c(T) for one xi (data/reference/pnjl/boundary.csv + cep.csv;
 * PhaseBoundaryData, SeedStrategies.jl:365-371).  n may be 0 and T_CEP NaN (no data for that xi). */
typedef struct pnjl_boundary {
    const double* T_MeV;    /* [n] ascending */
    const double* mu_c_MeV; /* [n] */
    int32_t n;              /* any length (bisection on the device) */
    double T_CEP_MeV;       /* NaN if unknown */
} pnjl_boundary;

void pnjl_default_config(pnjl_config* cfg);  /* config/pnjl/default.toml values, 64x8 nodes, max_iter 1000 */
int pnjl_abi_version(void);
const char* pnjl_last_error(void);

/* Layout self-check for hand-written mirrors of the structs above (the Julia `struct Config` of julia/PNJLB200.jl, the ctypes
 * Structure of julia_relaxtime_b200/_abi.py): sizes in bytes, and the byte offset of a field by name (-1: no such field).
 * Bindings assert these against their own sizeof/fieldoffset when they load the library. */
int64_t pnjl_sizeof_config(void);
int64_t pnjl_sizeof_boundary(void);
int64_t pnjl_sizeof_stats(void);
int64_t pnjl_config_field_offset(const char* field);

/* Page-locked, device-addressable host memory for result buffers (what a Julia caller wraps with unsafe_wrap). */
int pnjl_alloc_pinned(uint64_t bytes, void** out);
int pnjl_free_pinned(void* p);

int pnjl_create(const pnjl_config* cfg, pnjl_handle** out);
void pnjl_destroy(pnjl_handle* h);

/* gauleg(a, b, n) (src/integration/GaussLegendre.jl:94-119): host-side helper so callers without
 * FastGaussQuadrature can build the same rule the library would. */
int pnjl_gauleg(double a, double b, int32_t n, double* nodes, double* weights);

/* Run-time options of a handle (launch geometry, nothing that changes results beyond round-off):
 *   "schedule"        0 (automatic), 1, 2, 3: as pnjl_config.schedule
 *   "march_parts"     warps per line (a leader and followers that only sweep) in the line-march kernel: 0 automatic, 1 .. 16
 *   "march_quantum"   points per time slice of a line in the line-march kernel (0 automatic)
 *   "isotropic_batch" 1: the caller promises xi == 0 on every line of the following *_device calls (the host entry points
 *                     look at xi themselves), so teams are sized for p_num nodes instead of p_num * t_num */
int pnjl_set_option(pnjl_handle* h, const char* key, int64_t value);

/* Record and aux buffers (`records`, `d_records`, `aux`, ...) must be 16-byte aligned: the kernels write them with 16-byte
 * stores.  malloc, numpy, Julia arrays and cudaMalloc all satisfy this. */

/* Independent points: PNJL.solve / solve_multi at n (T, mu, xi) triples.  records: [n][32]. */
int pnjl_solve_points_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* xi,
                           int32_t seed_mode, int32_t n_seeds, const double* seeds, double* records);
int pnjl_solve_points_device(pnjl_handle* h, int64_t n, const double* d_T_fm, const double* d_mu_fm,
                             const double* d_xi, int32_t seed_mode, int32_t n_seeds, const double* d_seeds,
                             double* d_records, void* stream);

/* Continuity lines in run_gap_transport_scan.jl order (:407-443): line l = (xi[l], muq_MeV[l]) marches
 * T_MeV[0..n_T) ascending; MultiSeed while its tracker has no previous converged solution, then
 * PhaseAwareContinuitySeed with table table_idx[l] (-1: no data).  records: [n_lines][n_T][32].
 * muq_MeV, xi, table_idx are per line; T_MeV is shared by all lines. */
int pnjl_set_boundaries(pnjl_handle* h, int32_t n_tables, const pnjl_boundary* tables);
int pnjl_scan_lines_host(pnjl_handle* h, int64_t n_lines, const double* muq_MeV, const double* xi,
                         const int32_t* table_idx, int32_t n_T, const double* T_MeV, double* records);
int pnjl_scan_lines_device(pnjl_handle* h, int64_t n_lines, const double* d_muq_MeV, const double* d_xi,
                           const int32_t* d_table_idx, int32_t n_T, const double* d_T_MeV, double* d_records,
                           void* stream);

/* Same scan, but line l writes its n_T records at record line d_out_index[l] of d_records_base (instead of l).  With a base that
 * is another GPU's buffer (pnjl_ipc_alloc on rank 0, pnjl_ipc_open on the others) every rank's kernel stores its mu-slab
 * straight into rank 0's result array over NVLink while it runs: the final gather of the multi-GPU scan is fused into the
 * kernel's stores and costs no time of its own (julia_relaxtime_b200/distributed.py: PeerRecords, scan_sharded_peer). */
int pnjl_scan_lines_device_indexed(pnjl_handle* h, int64_t n_lines, const double* d_muq_MeV, const double* d_xi,
                                   const int32_t* d_table_idx, int32_t n_T, const double* d_T_MeV, double* d_records_base,
                                   const int64_t* d_out_index, void* stream);
/* Device buffers that other processes on the node can map (CUDA IPC): alloc on the owner (current device), open/close on
 * the peers (peer access is enabled on open), free on the owner after every peer closed. */
int pnjl_ipc_alloc(uint64_t bytes, void** dptr, unsigned char handle[64]);
int pnjl_ipc_open(const unsigned char handle[64], void** dptr);
int pnjl_ipc_close(void* dptr);
int pnjl_ipc_free(void* dptr);

/* T-mu scan with TmuScan.run_tmu_scan semantics (src/pnjl/scans/TmuScan.jl:120-234): line l = (xi[l], T_MeV[l]) marches
 * mu_MeV[0..n_mu) in the given order with a fresh PhaseAwareContinuitySeed tracker per line, the four-candidate seed
 * list, solve() with its automatic fallbacks per candidate, the 1e-4 acceptance / refine / force-promote rules.
 * records: [n_lines][n_mu][32]; rows without any successful candidate carry PNJL_ST_NO_RESULT and NaNs. */
int pnjl_tmu_scan_host(pnjl_handle* h, int64_t n_lines, const double* T_MeV, const double* xi, const int32_t* table_idx,
                       int32_t n_mu, const double* mu_MeV, double* records);
int pnjl_tmu_scan_device(pnjl_handle* h, int64_t n_lines, const double* d_T_MeV, const double* d_xi,
                         const int32_t* d_table_idx, int32_t n_mu, const double* d_mu_MeV, double* d_records,
                         void* stream);

/* Dual-branch scan, DualBranchScan.run_dual_branch_scan (src/pnjl/scans/DualBranchScan.jl:104-182): for line l = (xi[l],
 * T_MeV[l]) the hadron branch marches mu_MeV ascending from the hadron seed and the quark branch marches it descending
 * from the quark seed, each with plain continuity seeding and each stopping at its first non-converged or jumping point
 * (:420-429).  records: [n_lines][2][n_mu][32], branch 0 = hadron, 1 = quark, indexed by the position in mu_MeV; points a
 * branch does not reach carry PNJL_ST_NO_RESULT and NaNs.  Physical-branch selection, the Omega crossing and the merged
 * CSV are host arithmetic (julia_relaxtime_b200/dual_branch.py; DualBranchScan.jl:190-330). */
int pnjl_dual_branch_host(pnjl_handle* h, int64_t n_lines, const double* T_MeV, const double* xi, int32_t n_mu,
                          const double* mu_MeV, double* records);
int pnjl_dual_branch_device(pnjl_handle* h, int64_t n_lines, const double* d_T_MeV, const double* d_xi, int32_t n_mu,
                            const double* d_mu_MeV, double* d_records, void* stream);

/* ---- one-loop integral A and effective couplings (the per-point step after the gap solve) ------------------------
 * build_K_data of scripts/relaxtime/run_gap_transport_scan.jl:297-305:
 *     A_f = OneLoopIntegrals.A(m_f, mu, T, Phi, Phibar, nodes, weights)            src/relaxtime/OneLoopIntegrals.jl:531-543
 *     G_f = calculate_G_from_A(A_f, m_f)                                           src/relaxtime/EffectiveCouplings.jl:56-60
 *     K   = calculate_effective_couplings(G_fm2, K_fm5, G_u, G_s)                  EffectiveCouplings.jl:232-279
 * One 16-double AUX record per point (a Julia Matrix{Float64}(16, n)), offsets PNJL_AUX_* below.  The quadrature rule is
 * the handle's one-loop rule: gauleg(0, 10, 64) = DEFAULT_MOMENTUM_NODES/WEIGHTS (src/integration/GaussLegendre.jl:170)
 * unless replaced with pnjl_set_oneloop_rule (the reference's A takes the rule as an argument). */
#define PNJL_AUX_DOUBLES 16
#define PNJL_AUX_A_U 0
#define PNJL_AUX_A_S 1
#define PNJL_AUX_G_U 2
#define PNJL_AUX_G_S 3
#define PNJL_AUX_K0_PLUS 4
#define PNJL_AUX_K0_MINUS 5
#define PNJL_AUX_K123_PLUS 6
#define PNJL_AUX_K123_MINUS 7
#define PNJL_AUX_K4567_PLUS 8
#define PNJL_AUX_K4567_MINUS 9
#define PNJL_AUX_K8_PLUS 10
#define PNJL_AUX_K8_MINUS 11
#define PNJL_AUX_K08_PLUS 12
#define PNJL_AUX_K08_MINUS 13
#define PNJL_AUX_DETK_PLUS 14
#define PNJL_AUX_DETK_MINUS 15
#define PNJL_MAX_ONELOOP_NODES 512
int pnjl_set_oneloop_rule(pnjl_handle* h, int32_t n_nodes, const double* nodes, const double* weights);
/* n independent states given as arrays (what a caller of A(m, mu, T, Phi, Phibar, ...) has at hand). */
int pnjl_effective_couplings_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* m_u,
                                  const double* m_s, const double* Phi, const double* Phibar, double* aux);
/* From result records on the device (masses, Phi, Phibar, T, mu are read from each record): d_aux [n][16]. */
int pnjl_effective_couplings_device(pnjl_handle* h, int64_t n, const double* d_records, double* d_aux, void* stream);
/* pnjl_scan_lines_host followed by the couplings of every point, one call and one round trip: records [n_lines][n_T][32],
 * aux [n_lines][n_T][16]. */
int pnjl_scan_lines_couplings_host(pnjl_handle* h, int64_t n_lines, const double* muq_MeV, const double* xi,
                                   const int32_t* table_idx, int32_t n_T, const double* T_MeV, double* records,
                                   double* aux);

/* Single Omega-gradient/Jacobian evaluation at given states (test hook for per-iterate parity):
 * FJ: [n][30] = F[5] then J[5][5] row-major. */
int pnjl_eval_fj_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* xi,
                      const double* x /* [n][5] */, double* FJ);

/* F, J and the thermodynamic functions at given states x (no solve): what ThermoDerivatives.jl evaluates around a solution
 * (gap_conditions :88-98, calculate_thermo / calculate_rho :360-420).  out: [n][48] =
 * F[5], J[5][5] row-major, then Omega 30, P 31, rho_norm 32, s 33, eps 34, rho_i[3] 35, n_q[3] 38, n_qbar[3] 41, M_i[3] 44, 0.
 * julia_relaxtime_b200/thermo_derivatives.py builds bulk_viscosity_coefficients / thermo_derivatives / mass_derivatives on it. */
#define PNJL_STATE_DOUBLES 48
int pnjl_eval_state_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* xi,
                         const double* x /* [n][5] */, double* out);

/* The same plus the partial derivatives in (T, mu) at fixed x that the implicit differentiation of the gap equations needs
 * (ThermoDerivatives.jl:80-109: dx/dtheta = -J^{-1} dF/dtheta; :186-250, :342-467), as closed-form quadrature sums over the
 * same mesh (the reference gets them from ForwardDiff): out: [n][64] = the 48 doubles above, then
 * dF/dT[5] 48, dF/dmu[5] 53, ds/dT 58, ds/dmu 59, dn_B/dT 60, dn_B/dmu 61 (n_B = sum_i rho_i / 3), 0, 0.
 * julia_relaxtime_b200/thermo_derivatives.py solves the 5x5 systems and applies the reference's algebra. */
#define PNJL_DERIV_DOUBLES 64
int pnjl_eval_derivs_host(pnjl_handle* h, int64_t n, const double* T_fm, const double* mu_fm, const double* xi,
                          const double* x /* [n][5] */, double* out);

/* Accuracy self-test of the kernels' branch-free FP64 primitives (test hook):
 * which = 0: exp(x) for x in [-708, 0];  1: 1/x;  2: 1/sqrt(x)  (x normal, positive for 1 and 2). */
int pnjl_selftest_math(pnjl_handle* h, int64_t n, const double* x, int32_t which, double* out);

/* Measured launch statistics of the last *_device / *_host call on this handle. */
typedef struct pnjl_stats {
    int64_t kernel_launches;   /* kernels this library launched in the last call */
    double kernel_ms;          /* device time of the solve kernel(s) of the last *_host call (CUDA events) */
    int32_t lanes_per_solve;   /* lanes cooperating on one solve */
    int32_t blocks, threads;   /* launch geometry of the solve kernel */
    int32_t regs_per_thread;   /* cudaFuncGetAttributes */
    int32_t smem_bytes;        /* dynamic shared memory per block */
} pnjl_stats;
int pnjl_get_stats(pnjl_handle* h, pnjl_stats* out);

/* FP64 FMA peak of the device (register-resident DFMA chains; the roofline denominator the north
 * star asks for, since MEASURED_PEAKS.json has no FP64 figure).  *tflops_burst = best single launch;
 * *tflops_sustained (may be NULL) = average over back-to-back launches lasting `seconds`. */
int pnjl_measure_fp64_peak(pnjl_handle* h, double seconds, double* tflops_burst, double* tflops_sustained);

#ifdef __cplusplus
}
#endif
#endif /* PNJL_B200_H */
