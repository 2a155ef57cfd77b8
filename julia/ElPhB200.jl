# ElPhB200.jl -- Julia shim that routes the hot path of ElPhDynamics through libelph_b200.so.
#
# UNEXECUTED in this repository's environment (no Julia in the image); it documents exactly which
# reference methods bind to which C symbols (include/elph_b200.h).  Usage from the reference driver:
#
#     using ElPhDynamics, ElPhB200
#     model  = ...                       # built by ProcessInputFile as usual (HolsteinModel / SSHModel)
#     gmodel = B200Model(model)          # uploads tables, x stays authoritative on the device
#     P      = B200KPM(gmodel, 20, 0.05, 1.0, 1.0)
#     iters  = evolve!(gmodel, dyn, fa, P)      # same call the driver makes (RunSimulation.jl:62,88)
#     pull_x!(gmodel)                    # before measurements / checkpoints
#
module ElPhB200

using LinearAlgebra, Random
import LinearAlgebra: mul!, ldiv!
using ElPhDynamics.Models: AbstractModel, HolsteinModel, SSHModel
import ElPhDynamics.Models: mulM!, mulMᵀ!, mulMᵀM!, muldMdx!, update_model!
import ElPhDynamics.KPMPreconditioners: setup!
import ElPhDynamics.LangevinDynamics: evolve!, EulerDynamics, RungeKuttaDynamics, HeunsDynamics
import ElPhDynamics.FourierAcceleration: FourierAccelerator

const LIB = get(ENV, "ELPH_B200_LIB", "libelph_b200.so")

# mirrors `elph_config` field for field (include/elph_b200.h)
Base.@kwdef mutable struct ElphConfig
    model::Int32 = 0; index_base::Int32 = 1; device::Int32 = -1; reserved0::Int32 = 0
    Ltau::Int64 = 0; Nsites::Int64 = 0; Nbonds::Int64 = 0; Nph::Int64 = 0
    dtau::Float64 = 0.0
    neighbor_table::Ptr{Int64} = C_NULL
    cosht::Ptr{Float64} = C_NULL; sinht::Ptr{Float64} = C_NULL
    lambda::Ptr{Float64} = C_NULL; lambda2::Ptr{Float64} = C_NULL
    mu::Ptr{Float64} = C_NULL; omega::Ptr{Float64} = C_NULL; omega4::Ptr{Float64} = C_NULL
    t::Ptr{Float64} = C_NULL; alpha::Ptr{Float64} = C_NULL; alpha2::Ptr{Float64} = C_NULL
    checkerboard_perm::Ptr{Int64} = C_NULL; inv_checkerboard_perm::Ptr{Int64} = C_NULL
    phonon_to_bond::Ptr{Int64} = C_NULL; bond_to_phonon::Ptr{Int64} = C_NULL; primary_field::Ptr{Int64} = C_NULL
    cg_tol::Float64 = 1e-5; cg_maxiter::Int64 = 0; cg_kappa_max::Float64 = 1e12
    kpm_n::Int64 = 0; kpm_buf::Float64 = 0.05; kpm_c1::Float64 = 1.0; kpm_c2::Float64 = 1.0
    fa_Q::Ptr{Float64} = C_NULL; fa_M::Ptr{Float64} = C_NULL
end

struct SolveInfo; iters::Int64; residual::Float64; flag::Int32; used_fallback::Int32; pcg_iters::Int64; end
struct KpmInfo; active::Int32; recomputed::Int32; e_min::Float64; e_max::Float64; lambda_lo::Float64; lambda_hi::Float64; total_order::Int64; max_order::Int64; end

check(st, h) = st == 0 || error(unsafe_string(ccall((:elph_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

mutable struct B200Model{T1,T2,T3,T4,M<:AbstractModel{T1,T2,T3,T4}} <: AbstractModel{T1,T2,T3,T4}
    host::M              # the reference model: parameters, rng, solver settings stay here
    h::Ptr{Cvoid}        # elph_handle*
end

function B200Model(m::HolsteinModel)
    cfg = ElphConfig(model=0, Ltau=m.Lτ, Nsites=m.Nsites, Nbonds=m.Nbonds, Nph=m.Nph, dtau=m.Δτ,
        neighbor_table=pointer(m.neighbor_table), cosht=pointer(m.cosht), sinht=pointer(m.sinht),
        lambda=pointer(m.λ), lambda2=pointer(m.λ₂), mu=pointer(m.μ), omega=pointer(m.ω), omega4=pointer(m.ω₄),
        cg_tol=m.solver.tol, cg_maxiter=m.solver.maxiter, cg_kappa_max=m.solver.κmax)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve m check(ccall((:elph_create, LIB), Int32, (Ref{ElphConfig}, Ref{Ptr{Cvoid}}), cfg, h), C_NULL)
    g = B200Model{eltype(m.x),eltype(m.v′),typeof(m.solver),typeof(m.rng),typeof(m)}(m, h[])
    push_x!(g); update_model!(g)
    finalizer(x -> ccall((:elph_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), g)
end
# SSH: phonons on bonds (src/SSHModels.jl:79-314).  The index maps go over as stored (1-based Int64, index_base = 1):
# checkerboard_perm / inv_checkerboard_perm (:166-173), phonon_to_bond / bond_to_phonon (0 = no phonon, :155-160),
# primary_field (:152); t, alpha, alpha2 in the ORIGINAL bond order (:121-133); the engine builds t', cosh, sinh itself.
function B200Model(m::SSHModel)
    cfg = ElphConfig(model=1, Ltau=m.Lτ, Nsites=m.Nsites, Nbonds=m.Nbonds, Nph=m.Nph, dtau=m.Δτ,
        neighbor_table=pointer(m.neighbor_table), mu=pointer(m.μ), omega=pointer(m.ω), omega4=pointer(m.ω₄),
        t=pointer(m.t), alpha=pointer(m.α), alpha2=pointer(m.α₂),
        checkerboard_perm=pointer(m.checkerboard_perm), inv_checkerboard_perm=pointer(m.inv_checkerboard_perm),
        phonon_to_bond=pointer(m.phonon_to_bond), bond_to_phonon=pointer(m.bond_to_phonon), primary_field=pointer(m.primary_field),
        cg_tol=m.solver.tol, cg_maxiter=m.solver.maxiter, cg_kappa_max=m.solver.κmax)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve m check(ccall((:elph_create, LIB), Int32, (Ref{ElphConfig}, Ref{Ptr{Cvoid}}), cfg, h), C_NULL)
    g = B200Model{eltype(m.x),eltype(m.v′),typeof(m.solver),typeof(m.rng),typeof(m)}(m, h[])
    push_x!(g); update_model!(g)
    finalizer(x -> ccall((:elph_destroy, LIB), Int32, (Ptr{Cvoid},), x.h), g)
end

Base.getproperty(g::B200Model, s::Symbol) = s in (:host, :h) ? getfield(g, s) : getproperty(getfield(g, :host), s)

push_x!(g::B200Model) = check(ccall((:elph_set_x, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), g.h, g.host.x), g.h)
pull_x!(g::B200Model) = check(ccall((:elph_get_x, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), g.h, g.host.x), g.h)

# --- operators: src/HolsteinModels.jl:526,569,631,691 ; src/Models.jl:192,215 -----------------------------------
update_model!(g::B200Model) = check(ccall((:elph_update_model, LIB), Int32, (Ptr{Cvoid},), g.h), g.h)
for (jl, c) in ((:mulM!, :elph_mulM), (:mulMᵀ!, :elph_mulMT), (:mulMᵀM!, :elph_mulMTM))
    @eval $jl(y::Vector{Float64}, g::B200Model, v::Vector{Float64}) =
        check(ccall(($(QuoteNode(c)), LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), g.h, v, y), g.h)
end
mul!(y::Vector{Float64}, g::B200Model, v::Vector{Float64}) = mulMᵀM!(y, g, v)     # CG: mul_by_M = false
muldMdx!(d::Vector{Float64}, u::Vector{Float64}, g::B200Model, v::Vector{Float64}) =
    check(ccall((:elph_muldMdx, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), g.h, u, v, d), g.h)

# --- preconditioner: src/KPMPreconditioners.jl:219,259,426 -------------------------------------------------------
struct B200KPM; g::B200Model; end
function B200KPM(g::B200Model, n, buf, c1, c2)
    check(ccall((:elph_kpm_configure, LIB), Int32, (Ptr{Cvoid}, Int64, Float64, Float64, Float64), g.h, n, buf, c1, c2), g.h)
    B200KPM(g)
end
function setup!(P::B200KPM)
    noise = randn(P.g.host.rng, 2 * P.g.host.Nsites)     # the 2N draws of arnoldi_eigenvalue_bounds! (:859-861,:902-904)
    info = Ref{KpmInfo}()
    check(ccall((:elph_kpm_setup, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ref{KpmInfo}), P.g.h, noise, info), P.g.h)
    info[]
end
ldiv!(vout::Vector{Float64}, P::B200KPM, vin::Vector{Float64}) =
    check(ccall((:elph_kpm_apply, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), P.g.h, vin, vout), P.g.h)

# --- solve: src/Models.jl:74-186 returns (iters, residual_error, flag) ------------------------------------------
function ldiv!(x::Vector{Float64}, g::B200Model, b::Vector{Float64}, P=I; maxiter::Int=0)
    info = Ref{SolveInfo}()
    check(ccall((:elph_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ref{SolveInfo}),
                g.h, b, x, P isa B200KPM ? 1 : 0, 1.0, info), g.h)
    info[].iters, info[].residual, Int(info[].flag)
end

# --- measurement solves: update!(Gr, model, P) src/GreensFunctions.jl:201-234, all n_v vectors in one call -------
function update!(est::EstimateGreensFunction, g::B200Model, P=I)
    P isa B200KPM && setup!(P)
    randn!(g.host.rng, est.R)                          # column k = the k-th randn!(model.rng, r1) of the reference loop
    infos = Vector{SolveInfo}(undef, est.nᵥ)
    check(ccall((:elph_Minv_batch, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{SolveInfo}),
                g.h, est.nᵥ, est.R, est.M⁻¹R, P isa B200KPM ? 1 : 0, infos), g.h)
    nothing
end

# --- Green's-function convolutions: setup!(estimator, n1, n2) src/GreensFunctions.jl:239-296 on the device ---------
# (after update! above: R and M⁻¹R are uploaded once per measurement, the four 6-dimensional arrays come back per pair)
function load!(est::EstimateGreensFunction, g::B200Model)
    check(ccall((:elph_greens_load, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), g.h, est.nᵥ, est.R, est.M⁻¹R), g.h)
end
function setup!(est::EstimateGreensFunction, g::B200Model, n₁::Int, n₂::Int)
    est.n₁ = n₁; est.n₂ = n₂
    check(ccall((:elph_greens_setup, LIB), Int32,
                (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                g.h, n₁ - 1, n₂ - 1, est.L₁, est.L₂, est.L₃, est.nₛ, est.GΔ0, est.GΔ0_GΔ0, est.GΔΔ_G00, est.GΔ0_G0Δ), g.h)
    nothing
end

# --- special updates: src/SpecialUpdates.jl:97-160 (reflection), :233-366 (swap); one device call per proposal ------
# kind 0: x[:,i] = -x[:,i] (Holstein); kind 1: swap!(x[:,i], x[:,j]).  Sampling and the acceptance ratio stay here.
function special_proposal!(g::B200Model, kind::Int, i::Int, j::Int, P=I)::Bool
    m, rng = g.host, g.host.rng
    R₊ = randn(rng, m.Ndim); R₋ = randn(rng, m.Ndim)    # refresh_ϕ!(hmc, model, sample_R=true)      HMC.jl:674-677
    a = P isa B200KPM ? randn(rng, 2 * m.Nsites) : Float64[]
    u = rand(rng)                                      # rand(model.rng) < Pf                      SpecialUpdates.jl:146
    acc = Ref{Int32}(0); S₀ = Ref{Float64}(0); S₁ = Ref{Float64}(0); it = Ref{Int64}(0); fl = Ref{Int32}(0)
    check(ccall((:elph_hmc_special_update, LIB), Int32,
                (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64,
                 Ref{Int32}, Ref{Float64}, Ref{Float64}, Ref{Int64}, Ref{Int32}),
                g.h, kind, i - 1, j - 1, R₊, R₋, P isa B200KPM ? a : C_NULL, P isa B200KPM ? 1 : 0, u, acc, S₀, S₁, it, fl), g.h)
    acc[] == 1
end

# --- long-lived host arrays (dyn.η, est.R, ...) can be page-locked once so that their copies run at PCIe rate ------
pin!(g::B200Model, v::Array{Float64}) = check(ccall((:elph_host_register, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), g.h, v, sizeof(v)), g.h)
unpin!(g::B200Model, v::Array{Float64}) = check(ccall((:elph_host_unregister, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), g.h, v), g.h)

# --- dynamics: src/LangevinDynamics.jl:81,162,272 ; noise drawn here in the reference's order ---------------------
method(::EulerDynamics) = Int32(1); method(::RungeKuttaDynamics) = Int32(2); method(::HeunsDynamics) = Int32(3)
function attach!(g::B200Model, fa::FourierAccelerator)   # after update_Q!/update_M! (ProcessInputFile.jl:516-535)
    check(ccall((:elph_set_fourier_acceleration, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), g.h, fa.Q, fa.M), g.h)
end
function evolve!(g::B200Model, dyn, fa::FourierAccelerator, P=I)::Int
    m, rng = g.host, g.host.rng
    usep = P isa B200KPM
    draw(n) = randn(rng, n)
    η = draw(m.Ndof)                                   # randn!(η, model)       :97 / :181 / :287
    two = !(dyn isa EulerDynamics)
    _ = draw(m.Ndim); g1 = draw(m.Ndim)                # wasted randn!(R) then g :100,:360
    a1 = usep ? draw(2 * m.Nsites) : Float64[]         # setup!(P)               :364
    g2 = Float64[]; a2 = Float64[]
    if two
        _ = draw(m.Ndim); g2 = draw(m.Ndim); a2 = usep ? draw(2 * m.Nsites) : Float64[]
    end
    iters = Ref{Int64}(0)
    check(ccall((:elph_langevin_step, LIB), Int32,
                (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32,
                 Ref{Int64}, Ptr{Cvoid}, Ptr{Cvoid}),
                g.h, method(dyn), dyn.Δt, η, g1, two ? g2 : C_NULL, usep ? a1 : C_NULL, (usep && two) ? a2 : C_NULL,
                usep ? 1 : 0, iters, C_NULL, C_NULL), g.h)
    Int(iters[])
end

# --- HMC: update!(model, hmc, fa, P) src/HMC.jl:310-345 -> one whole trajectory on the device (elph_hmc_update) ---------------
# The draws are made here in the reference's order: refresh_v! (randn!(R, model), :655), refresh_ϕ! (R₊, R₋, :675-676), the 2N
# Arnoldi start values of every setup!(P) of the trajectory (Nt + 2 force / action evaluations, call order), and last the
# Metropolis uniform (:453 / :618).  Nothing else draws in between, so taking them up front leaves model.rng's stream as the
# reference's.  fa must have been attached (attach!) after update_M!.  On return x lives on the device (pull_x! before measuring).
import ElPhDynamics.HMC: HybridMonteCarlo, update!
function update!(g::B200Model, hmc::HybridMonteCarlo, fa::FourierAccelerator, P=I)
    hmc.Ndof > 0 || return true, 0.0
    m, rng = g.host, g.host.rng
    usep = P isa B200KPM
    hmc.t = 0
    Rv = randn(rng, m.Ndof)
    R₊ = randn(rng, m.Ndim); R₋ = randn(rng, m.Ndim)
    a = usep ? randn(rng, 2 * m.Nsites * (hmc.Nt + 2)) : Float64[]
    u = rand(rng)
    acc = Ref{Int32}(0); it = Ref{Float64}(0); H₀ = Ref{Float64}(0); H₁ = Ref{Float64}(0); fl = Ref{Int32}(0)
    check(ccall((:elph_hmc_update, LIB), Int32,
                (Ptr{Cvoid}, Float64, Int64, Int64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64,
                 Ref{Int32}, Ref{Float64}, Ref{Float64}, Ref{Float64}, Ref{Int32}),
                g.h, hmc.Δt, hmc.Nt, hmc.Nb, hmc.α, Rv, R₊, R₋, usep ? a : C_NULL, usep ? 1 : 0, u, acc, it, H₀, H₁, fl), g.h)
    hmc.updates += 1
    acc[] == 1, it[]
end

end # module
