# PNJLB200.jl — Julia binding of libpnjl_b200.so (include/pnjl_b200.h) for Julia_RelaxTime.
#
# NOT EXECUTED IN THE BUILD IMAGE: there is no Julia toolchain in this container or on the GPU box, so this file is
# the binding a maintainer adds to the reference (see INTEGRATION.md); its Python twin
# (julia_relaxtime_b200/_lib.py + solver.py + scan.py) is what the tests drive.  Every `ccall` below matches one
# prototype of include/pnjl_b200.h; struct layouts match `pnjl_config` / `pnjl_boundary` field by field.
#
# Drop-in points in the reference:
#   * PNJL.solve(::FixedMu, T_fm, μ_fm; ...)            src/pnjl/solver/ImplicitSolver.jl:211  → PNJLB200.solve
#   * PNJL.solve_multi(::FixedMu, ...)                   src/pnjl/solver/ImplicitSolver.jl:532  → PNJLB200.solve (MultiSeed)
#   * the (xi, muB, T) loop of run_gap_transport_scan.jl scripts/relaxtime/run_gap_transport_scan.jl:407-443
#                                                                                                → PNJLB200.scan_lines
module PNJLB200

using StaticArrays

const LIB = get(ENV, "PNJL_B200_LIB", joinpath(@__DIR__, "..", "julia_relaxtime_b200", "csrc", "_build", "libpnjl_b200.so"))

const REC = 32                       # doubles per result record (PNJL_REC_DOUBLES)
const AUX = 16                       # doubles per couplings record (PNJL_AUX_DOUBLES)
const ABI_VERSION = 12               # PNJL_ABI_VERSION this file was written against (checked in __init__)
const ST_CONVERGED = 1
const ST_ALL_SEEDS_FAILED = 512
const ST_MASS_INVERSION = 32768      # converged on the high-Omega root with M_s <= M_u (flagged, values unchanged)
const SEED_EXPLICIT, SEED_AUTO, SEED_MULTI = Int32(0), Int32(1), Int32(2)

# struct pnjl_config (field order and types as in the header)
struct Config
    hbarc::Cdouble; Lambda::Cdouble; m_ud0::Cdouble; m_s0::Cdouble; G::Cdouble; K::Cdouble
    T0::Cdouble; a0::Cdouble; a1::Cdouble; a2::Cdouble; b3::Cdouble; rho0::Cdouble
    Nc::Int32; p_num::Int32; t_num::Int32
    p_nodes::Ptr{Cdouble}; p_w::Ptr{Cdouble}; c_nodes::Ptr{Cdouble}; c_w::Ptr{Cdouble}
    xtol::Cdouble; ftol::Cdouble; residual_norm_max::Cdouble; phi_tol::Cdouble
    max_iter::Int32; tr_fallback::Int32; auto_multiseed_fallback::Int32
    omega_tie_rel::Cdouble
    device::Int32; lanes_per_solve::Int32
    predict_tol::Cdouble
    isospin_symmetric::Int32
    schedule::Int32
    isotropic_collapse::Int32
end

struct Boundary                      # struct pnjl_boundary
    T_MeV::Ptr{Cdouble}; mu_c_MeV::Ptr{Cdouble}; n::Int32; T_CEP_MeV::Cdouble
end

mutable struct Engine
    handle::Ptr{Cvoid}
    keep::Vector{Any}                # node vectors must outlive pnjl_create only; kept for clarity
end

last_error() = unsafe_string(ccall((:pnjl_last_error, LIB), Cstring, ()))
check(rc, what) = rc == 0 ? nothing : error("$what failed ($rc): $(last_error())")

"""
Layout self-check, run when the module loads: the hand-written `Config` / `Boundary` mirrors must have the size and the
field offsets of the C structs of the library that was found (pnjl_sizeof_config, pnjl_config_field_offset), and the
library must speak the ABI version this file was written against.  A mismatch is an error here, not a silent
mis-read of the options inside `pnjl_create`.
"""
function check_layout()
    v = ccall((:pnjl_abi_version, LIB), Cint, ())
    v == ABI_VERSION || error("libpnjl_b200.so has ABI version $v, PNJLB200.jl was written for $ABI_VERSION")
    sizeof(Config) == ccall((:pnjl_sizeof_config, LIB), Int64, ()) ||
        error("sizeof(Config) = $(sizeof(Config)) differs from the library's pnjl_config")
    sizeof(Boundary) == ccall((:pnjl_sizeof_boundary, LIB), Int64, ()) || error("struct Boundary differs from pnjl_boundary")
    for (i, name) in enumerate(fieldnames(Config))
        off = ccall((:pnjl_config_field_offset, LIB), Int64, (Cstring,), String(name))
        off == Int64(fieldoffset(Config, i)) || error("Config.$name at offset $(fieldoffset(Config, i)), pnjl_config.$name at $off")
    end
    return true
end
__init__() = check_layout()

"""Run-time option of a handle (pnjl_set_option): "schedule", "march_parts", "march_quantum", "isotropic_batch"."""
set_option!(e, key::AbstractString, value::Integer) =
    check(ccall((:pnjl_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), e.handle, String(key), Int64(value)), "pnjl_set_option($key)")

"""
    Engine(; p_num=64, t_num=8, iterations=1000, trust_region_fallback=true, auto_multiseed_fallback=true,
             residual_norm_max=1e-6, device=-1, constants=Main.Constants_PNJL)

One library handle = one GPU + model constants + quadrature rule.  Nodes come from the reference's own
`GaussLegendre.gauleg` (FastGaussQuadrature) so that both paths integrate on bit-identical rules
(Integrals.jl:67-96).
"""
function Engine(; p_num::Int=64, t_num::Int=8, iterations::Int=1000, trust_region_fallback::Bool=true,
                auto_multiseed_fallback::Bool=true, residual_norm_max::Float64=1e-6, device::Int=-1,
                C=Main.Constants_PNJL, gauleg=Main.GaussLegendre.gauleg)
    pn, pw = gauleg(0.0, 10.0, p_num)
    cn, cw = gauleg(0.0, 1.0, t_num)
    cfg = Config(C.ħc_MeV_fm, C.Λ_inv_fm, C.m_ud0_inv_fm, C.m_s0_inv_fm, C.G_fm2, C.K_fm5, C.T0_inv_fm,
                 C.a0, C.a1, C.a2, C.b3, C.ρ0_inv_fm3, Int32(C.N_color), Int32(p_num), Int32(t_num),
                 pointer(pn), pointer(pw), pointer(cn), pointer(cw),
                 1e-9, 1e-9, residual_norm_max, 1e-8, Int32(iterations), Int32(trust_region_fallback),
                 Int32(auto_multiseed_fallback), 1e-12, Int32(device), Int32(0), 1e-4, Int32(1), Int32(0), Int32(1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve pn pw cn cw begin
        check(ccall((:pnjl_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h), "pnjl_create")
    end
    e = Engine(h[], Any[pn, pw, cn, cw])
    finalizer(x -> ccall((:pnjl_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), e)
    return e
end

"""Upload PhaseBoundaryData tables (SeedStrategies.jl:365-371), one per distinct xi."""
function set_boundaries!(e::Engine, tables::Vector)   # tables: Vector of PNJL.PhaseBoundaryData
    bs = [Boundary(pointer(t.T_values), pointer(t.mu_values), Int32(length(t.T_values)), t.T_CEP) for t in tables]
    GC.@preserve tables begin
        check(ccall((:pnjl_set_boundaries, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Boundary}), e.handle,
                    Int32(length(bs)), bs), "pnjl_set_boundaries")
    end
end

"""Independent points.  Returns a `Matrix{Float64}(32, n)`: column i is the record of point i."""
function solve_points(e::Engine, T_fm::Vector{Float64}, mu_fm::Vector{Float64}, xi::Vector{Float64};
                      seed_mode::Int32=SEED_MULTI, seeds::Union{Nothing,Array{Float64,3}}=nothing)
    n = length(T_fm)
    rec = Matrix{Float64}(undef, REC, n)
    n_seeds = seeds === nothing ? Int32(6) : Int32(size(seeds, 2))      # seeds[5, n_seeds, n]
    sp = seeds === nothing ? Ptr{Cdouble}(C_NULL) : pointer(seeds)
    GC.@preserve seeds begin
        check(ccall((:pnjl_solve_points_host, LIB), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                    e.handle, n, T_fm, mu_fm, xi, seed_mode, n_seeds, sp, rec), "pnjl_solve_points_host")
    end
    return rec
end

"""Page-locked result buffer: the kernels write into it while they run (no device->host copy afterwards).  Free with
`free_pinned(a)`; pass it as `out` to the `*!` variants or let `scan_lines(...; pinned=true)` do it."""
function alloc_pinned(dims::Int...)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pnjl_alloc_pinned, LIB), Cint, (UInt64, Ref{Ptr{Cvoid}}), UInt64(8 * prod(dims)), p), "pnjl_alloc_pinned")
    return unsafe_wrap(Array, Ptr{Float64}(p[]), dims; own=false)
end
free_pinned(a::Array{Float64}) = ccall((:pnjl_free_pinned, LIB), Cint, (Ptr{Cvoid},), pointer(a))

"""Continuity lines in run_gap_transport_scan.jl order.  Returns `Array{Float64}(32, n_T, n_lines)`."""
function scan_lines(e::Engine, muq_MeV::Vector{Float64}, xi::Vector{Float64}, table_idx::Vector{Int32},
                    T_MeV::Vector{Float64}; pinned::Bool=false)
    nl, nT = length(muq_MeV), length(T_MeV)
    rec = pinned ? alloc_pinned(REC, nT, nl) : Array{Float64}(undef, REC, nT, nl)
    check(ccall((:pnjl_scan_lines_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, nl, muq_MeV, xi, table_idx, Int32(nT), T_MeV, rec), "pnjl_scan_lines_host")
    return rec
end

"""T-mu scan with TmuScan.run_tmu_scan semantics (src/pnjl/scans/TmuScan.jl:120-234): line l = (xi[l], T_MeV[l]) marches
`mu_MeV` in the given order.  Returns `Array{Float64}(32, n_mu, n_lines)`; rows with `PNJL_ST_NO_RESULT` (bit 16384) are
the reference's all-NaN rows."""
function tmu_scan(e::Engine, T_MeV::Vector{Float64}, xi::Vector{Float64}, table_idx::Vector{Int32}, mu_MeV::Vector{Float64})
    nl, nmu = length(T_MeV), length(mu_MeV)
    rec = Array{Float64}(undef, REC, nmu, nl)
    check(ccall((:pnjl_tmu_scan_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, nl, T_MeV, xi, table_idx, Int32(nmu), mu_MeV, rec), "pnjl_tmu_scan_host")
    return rec
end

"""Both branches of `DualBranchScan.run_dual_branch_scan` (src/pnjl/scans/DualBranchScan.jl:104-182) for the lines
(xi[l], T_MeV[l]).  Returns `Array{Float64}(32, n_mu, 2, n_lines)`: branch 1 = hadron (mu ascending), 2 = quark (mu descending);
entries with status bit 16384 (`PNJL_ST_NO_RESULT`) are the `nothing`s of the reference's branch vectors."""
function dual_branch(e::Engine, T_MeV::Vector{Float64}, xi::Vector{Float64}, mu_MeV::Vector{Float64})
    nl, nmu = length(T_MeV), length(mu_MeV)
    rec = Array{Float64}(undef, REC, nmu, 2, nl)
    check(ccall((:pnjl_dual_branch_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, nl, T_MeV, xi, Int32(nmu), mu_MeV, rec), "pnjl_dual_branch_host")
    return rec
end

"""F (5), J (5x5 row-major) and the thermodynamic functions at given states `x` (5, n) without solving: `Matrix{Float64}(48, n)`
(PNJL_STATE layout: F 1:5, J 6:30, Ω 31, P 32, ρ_norm 33, s 34, ε 35, ρ_i 36:38, n_q 39:41, n_q̄ 42:44, M_i 45:47)."""
function eval_state(e::Engine, T_fm::Vector{Float64}, mu_fm::Vector{Float64}, xi::Vector{Float64}, x::Matrix{Float64})
    n = length(T_fm)
    out = Matrix{Float64}(undef, 48, n)
    check(ccall((:pnjl_eval_state_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, n, T_fm, mu_fm, xi, x, out), "pnjl_eval_state_host")
    return out
end

"""`eval_state` plus the partial derivatives in (T, μ) at fixed `x` that ThermoDerivatives.jl gets from ForwardDiff
(`:80-109`, `:186-262`, `:342-467`), as closed-form quadrature sums: `Matrix{Float64}(64, n)` — rows 1:48 as `eval_state`, then
∂F/∂T 49:53, ∂F/∂μ 54:58, ∂s/∂T 59, ∂s/∂μ 60, ∂n_B/∂T 61, ∂n_B/∂μ 62 (n_B = Σρ_i/3)."""
function eval_derivs(e::Engine, T_fm::Vector{Float64}, mu_fm::Vector{Float64}, xi::Vector{Float64}, x::Matrix{Float64})
    n = length(T_fm)
    out = Matrix{Float64}(undef, 64, n)
    check(ccall((:pnjl_eval_derivs_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, n, T_fm, mu_fm, xi, x, out), "pnjl_eval_derivs_host")
    return out
end

"""dx/dT and dx/dμ along the solution curve, dx/dθ = −J⁻¹ ∂F/∂θ (ThermoDerivatives.jl:80-109), from one column of `eval_derivs`."""
function state_derivatives(col::AbstractVector{Float64})
    J = permutedims(reshape(col[6:30], 5, 5))          # row-major in the record
    return -(J \ col[49:53]), -(J \ col[54:58])
end

# ---- one-loop integral A and effective couplings (build_K_data, run_gap_transport_scan.jl:297-305) -----------------
"""Replace the rule of A (default: DEFAULT_MOMENTUM_NODES / DEFAULT_MOMENTUM_WEIGHTS)."""
function set_oneloop_rule!(e::Engine, nodes::Vector{Float64}, weights::Vector{Float64})
    check(ccall((:pnjl_set_oneloop_rule, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, Int32(length(nodes)), nodes, weights), "pnjl_set_oneloop_rule")
end

"""`Matrix{Float64}(16, n)`: rows A_u, A_s, G_u, G_s, K0±, K123±, K4567±, K8±, K08±, det K± (PNJL_AUX_*)."""
function effective_couplings(e::Engine, T_fm::Vector{Float64}, mu_fm::Vector{Float64}, m_u::Vector{Float64},
                             m_s::Vector{Float64}, Phi::Vector{Float64}, Phibar::Vector{Float64})
    n = length(T_fm)
    aux = Matrix{Float64}(undef, AUX, n)
    check(ccall((:pnjl_effective_couplings_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, n, T_fm, mu_fm, m_u, m_s, Phi, Phibar, aux), "pnjl_effective_couplings_host")
    return aux
end

"""`scan_lines` plus the couplings of every point in one call: `(rec(32, n_T, n_lines), aux(16, n_T, n_lines))`."""
function scan_lines_couplings(e::Engine, muq_MeV::Vector{Float64}, xi::Vector{Float64}, table_idx::Vector{Int32},
                              T_MeV::Vector{Float64})
    nl, nT = length(muq_MeV), length(T_MeV)
    rec = Array{Float64}(undef, REC, nT, nl)
    aux = Array{Float64}(undef, AUX, nT, nl)
    check(ccall((:pnjl_scan_lines_couplings_host, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int32}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                e.handle, nl, muq_MeV, xi, table_idx, Int32(nT), T_MeV, rec, aux), "pnjl_scan_lines_couplings_host")
    return rec, aux
end

"""The script's `build_K_data` NamedTuple from one couplings record."""
function k_data(a::AbstractVector{Float64})
    K = (K0_plus=a[5], K0_minus=a[6], K123_plus=a[7], K123_minus=a[8], K4567_plus=a[9], K4567_minus=a[10],
         K8_plus=a[11], K8_minus=a[12], K08_plus=a[13], K08_minus=a[14], det_K_plus=a[15], det_K_minus=a[16])
    return (K_coeffs=K, A_vals=(u=a[1], d=a[1], s=a[2]))
end

"""Rebuild the reference's `SolverResult` (ImplicitSolver.jl:178-193) from one record."""
function solver_result(r::AbstractVector{Float64}; PNJL=Main.PNJL)
    st = Int(r[22])
    x = SVector{5}(r[1], r[2], r[3], r[4], r[5])
    mu = r[29]
    PNJL.SolverResult(PNJL.FixedMu(), (st & ST_CONVERGED) != 0, collect(x), x, SVector{3}(mu, mu, mu),
                      r[9], r[10], r[11], r[12], r[13], SVector{3}(r[6], r[7], r[8]), Int(r[21]), r[20], r[30])
end

const _ENGINES = Dict{Tuple,Engine}()
engine(p_num, t_num, iterations, tr, ams, rmax) =
    get!(() -> Engine(; p_num, t_num, iterations, trust_region_fallback=tr, auto_multiseed_fallback=ams,
                      residual_norm_max=rmax), _ENGINES, (p_num, t_num, iterations, tr, ams, rmax))

"""
    solve(PNJL.FixedMu(), T_fm, μ_fm; xi, seed_strategy, p_num, t_num, iterations, ...) -> PNJL.SolverResult

Same signature and seed semantics as `PNJL.solve` (ImplicitSolver.jl:211-328): the seed is obtained with the
reference's own `get_seed`, the solve (Newton → trust-region fallback → auto MultiSeed) runs on the GPU.
`error("All seeds failed …")` is raised exactly where the reference raises it (:553,:556).
"""
function solve(mode, T_fm::Real, mu_fm::Real; xi::Real=0.0, seed_strategy=Main.PNJL.DefaultSeed(), p_num::Int=64,
               t_num::Int=8, trust_region_fallback::Bool=true, auto_multiseed_fallback::Bool=true,
               residual_norm_max::Real=1e-6, iterations::Int=1000, PNJL=Main.PNJL)
    e = engine(p_num, t_num, iterations, trust_region_fallback, auto_multiseed_fallback, Float64(residual_norm_max))
    T, mu, x = [Float64(T_fm)], [Float64(mu_fm)], [Float64(xi)]
    multi = seed_strategy isa PNJL.MultiSeed ||
            (seed_strategy isa PNJL.PhaseAwareContinuitySeed && seed_strategy.bootstrap_multiseed &&
             seed_strategy.previous_solution === nothing)
    if multi
        ms = seed_strategy isa PNJL.MultiSeed ? seed_strategy : seed_strategy.bootstrap_strategy
        cand = PNJL.get_all_seeds(ms, [T[1], mu[1]], mode)
        seeds = reshape(reduce(hcat, cand), 5, length(cand), 1)
        rec = solve_points(e, T, mu, x; seed_mode=SEED_EXPLICIT, seeds=seeds)
        (Int(rec[22, 1]) & ST_ALL_SEEDS_FAILED) != 0 && error("All seeds failed to converge to a physical solution")
        return solver_result(view(rec, :, 1); PNJL)
    end
    x0 = Float64.(PNJL.get_seed(seed_strategy, [T[1], mu[1]], mode))
    rec = solve_points(e, T, mu, x; seed_mode=SEED_EXPLICIT, seeds=reshape(x0, 5, 1, 1))
    return solver_result(view(rec, :, 1); PNJL)
end

end # module
