# SmoQyElPhB200.jl -- `ccall` shim that keeps SmoQyElPhQMC's exported API for the hot path and forwards the
# arithmetic to libsmoqyelph_b200.so (include/smoqyelph_b200.h).
#
# NOT EXECUTED in the build container (no `julia` binary there): this file is the reference-side binding a
# maintainer would add; every C entry point it uses is exercised by the Python twin (smoqyelphqmc.jl_b200/api.py)
# in tests/.  Type names, type parameters used for dispatch and method signatures follow
# /root/reference/src/SmoQyElPhQMC.jl:60-124; struct internals are free to change (SURVEY.md 8b: no driver reads a
# field of these structs).
module SmoQyElPhB200

using LinearAlgebra, Random
import LinearAlgebra: mul!, lmul!, ldiv!
import SmoQyDQMC
import SmoQyDQMC: FermionPathIntegral, ElectronPhononParameters, hmc_update!, update_chemical_potential!,
                  reflection_update!, swap_update!, radial_update!, make_measurements!,
                  measure_onsite_energy, measure_hopping_energy, measure_bare_hopping_energy, measure_holstein_energy,
                  measure_ssh_energy
import MuTuner
using Checkerboard: checkerboard_decomposition!

export FermionDetMatrix, SymFermionDetMatrix, AsymFermionDetMatrix, KPMPreconditioner, SymKPMPreconditioner, AsymKPMPreconditioner,
       PFFCalculator, EFAPFFHMCUpdater, GreensEstimator

const LIB = get(ENV, "SMOQYELPH_B200_LIB", "libsmoqyelph_b200.so")

struct B200Error <: Exception
    msg::String
end
# status 3: NaN / non-finite residual, Lanczos coefficient or action -- the condition the reference's callers turn into a rejected
# update (src/EFAPFFHMCUpdater.jl:168-187, src/reflection_update.jl:111-127).  Everything else (bad argument, CUDA / NCCL error,
# watchdog time-out) is a B200Error and is never swallowed by this module.
struct B200NumericalInstability <: Exception
    msg::String
end
@inline function check(status::Cint)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:sq_last_error, LIB), Cstring, ()))
    status == 3 && throw(B200NumericalInstability(msg))
    throw(B200Error(msg))
end

# ---- FermionDetMatrix (src/FermionDetMatrix.jl:19-55, 66-111, 137-204) ------------------------------------------
abstract type FermionDetMatrix{T<:Number, E<:AbstractFloat} end

mutable struct CGConfig{E}     # keeps `fdm.cgs.tol` / `fdm.cgs.maxiter` readable (update_chemical_potential.jl:32-33)
    maxiter::Int
    tol::E
end

mutable struct SymFermionDetMatrix{T,E} <: FermionDetMatrix{T,E}
    h::Ptr{Cvoid}; Lτ::Int; N::Int; cgs::CGConfig{E}
end
mutable struct AsymFermionDetMatrix{T,E} <: FermionDetMatrix{T,E}
    h::Ptr{Cvoid}; Lτ::Int; N::Int; cgs::CGConfig{E}
end

function _create(::Type{F}, sym::Bool, fpi::FermionPathIntegral{T,E}; maxiter::Int, tol::E, device::Int) where {F,T,E}
    T <: Real || error("libsmoqyelph_b200 supports real hoppings only (SURVEY.md 9 Q9)")
    (; neighbor_table, N, Lτ) = fpi
    nt = copy(neighbor_table)
    perm, colors = checkerboard_decomposition!(nt)          # src/FermionDetMatrix.jl:95-97
    lo = Int64[first(r) for r in colors]; hi = Int64[last(r) for r in colors]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sq_fdm_create, LIB), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Cdouble, Int64, Cint),
                h, sym, Lτ, N, size(nt, 2), nt, Vector{Int64}(perm), length(colors), lo, hi, tol, maxiter, device))
    fdm = F{T,E}(h[], Lτ, N, CGConfig{E}(maxiter, tol))
    finalizer(f -> ccall((:sq_fdm_destroy, LIB), Cint, (Ptr{Cvoid},), f.h), fdm)
    update!(fdm, fpi)
    return fdm
end
SymFermionDetMatrix(fpi::FermionPathIntegral{T,E}; maxiter::Int = fpi.N * fpi.Lτ, tol::E = 1e-6, device::Int = 0) where {T,E} =
    _create(SymFermionDetMatrix, true, fpi; maxiter, tol, device)
AsymFermionDetMatrix(fpi::FermionPathIntegral{T,E}; maxiter::Int = fpi.N * fpi.Lτ, tol::E = 1e-6, device::Int = 0) where {T,E} =
    _create(AsymFermionDetMatrix, false, fpi; maxiter, tol, device)

Base.size(f::FermionDetMatrix) = (f.Lτ * f.N, f.Lτ * f.N)
Base.size(f::FermionDetMatrix, ::Int) = f.Lτ * f.N
Base.eltype(::FermionDetMatrix{T}) where {T} = T

# update!(fdm, fpi): src/FermionDetMatrix.jl:208-236
function update!(f::FermionDetMatrix{T,E}, fpi::FermionPathIntegral{T,E}) where {T,E}
    GC.@preserve fpi check(ccall((:sq_fdm_update, LIB), Cint, (Ptr{Cvoid}, Ptr{E}, Ptr{T}, Cdouble), f.h, fpi.V, fpi.t, fpi.Δτ))
    return nothing
end

for (name, op) in ((:mul_M!, 0), (:mul_Mt!, 1), (:mul_MtM!, 2), (:mul_MMt!, 3))     # :329-563
    @eval function $name(v′::AbstractVecOrMat{Complex{E}}, f::FermionDetMatrix{T,E}, v::AbstractVecOrMat{Complex{E}}) where {T,E}
        GC.@preserve v′ v check(ccall((:sq_fdm_mul, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Complex{E}}, Ptr{Complex{E}}), f.h, $op, v′, v))
        return nothing
    end
end
mul!(v′::AbstractVecOrMat, f::FermionDetMatrix, v::AbstractVecOrMat) = mul_MtM!(v′, f, v)      # :304-315
lmul!(f::FermionDetMatrix, v::AbstractVecOrMat) = mul_MtM!(v, f, v)                           # :292-301
lmul_M!(f::FermionDetMatrix, v) = mul_M!(v, f, v)
lmul_Mt!(f::FermionDetMatrix, v) = mul_Mt!(v, f, v)

# ---- one Markov chain over several GPUs (no counterpart in the reference; DESIGN.md section 5) -------------------
# Every rank builds the same operator with `device = local_rank`; `allgather` / `bcast` are the caller's MPI wrappers
# (e.g. `x -> MPI.Allgather(x, comm)`, `x -> MPI.bcast(x, 0, comm)`), so this file does not depend on MPI.jl.
function init_sharded_solve!(f::FermionDetMatrix, rank::Int, world::Int; bcast::Function, allgather::Function)
    id = zeros(UInt8, 128)
    rank == 0 && check(ccall((:sq_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    id = bcast(id)
    check(ccall((:sq_fdm_init_slab, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), f.h, rank, world, id))
    handle = zeros(UInt8, 64)
    check(ccall((:sq_fdm_mailbox_create, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), f.h, handle))
    handles = allgather(handle)                       # world x 64 bytes, rank-major
    check(ccall((:sq_fdm_mailbox_open, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), f.h, handles))
    check(ccall((:sq_fdm_set_sharded_solve, LIB), Cint, (Ptr{Cvoid}, Cint), f.h, 1))
    return nothing
end

# ---- KPMPreconditioner (src/KPMPreconditioner.jl:198-284, 554-597) ----------------------------------------------
mutable struct KPMPreconditioner{E}
    h::Ptr{Cvoid}; active::Bool; bounds::NTuple{2,E}
end
# one handle type serves both operator forms (the library picks the Sym / Asym expansion from the operator): the reference's two
# concrete type names (src/KPMPreconditioner.jl:61,132) are kept as aliases for code that dispatches on them
const SymKPMPreconditioner = KPMPreconditioner
const AsymKPMPreconditioner = KPMPreconditioner
function KPMPreconditioner(f::FermionDetMatrix{T,E}; rng::AbstractRNG = Random.default_rng(), rbuf::E = 0.10, n::Int = 20,
                           a1::E = 1.0, a2::E = 1.0) where {T,E}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sq_kpm_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Cdouble, Int64, Cdouble, Cdouble), h, f.h, rbuf, n, a1, a2))
    P = KPMPreconditioner{E}(h[], false, (zero(E), zero(E)))
    finalizer(p -> ccall((:sq_kpm_destroy, LIB), Cint, (Ptr{Cvoid},), p.h), P)
    check(ccall((:sq_kpm_set_seed, LIB), Cint, (Ptr{Cvoid}, UInt64), P.h, rand(rng, UInt64)))    # library-drawn Lanczos starts (solves inside device-resident trajectories)
    update_preconditioner!(P, f, rng)
    return P
end
# the randn!(rng, v) Lanczos start is drawn in Julia so the rng stream matches the reference (:634)
function update_preconditioner!(P::KPMPreconditioner{E}, f::FermionDetMatrix, rng::AbstractRNG) where {E}
    start = randn(rng, E, f.N); act = Ref{Cint}(0); b = zeros(E, 2)
    check(ccall((:sq_kpm_update, LIB), Cint, (Ptr{Cvoid}, Ptr{E}, Ref{Cint}, Ptr{E}), P.h, start, act, b))
    P.active = act[] != 0; P.bounds = (b[1], b[2])
    return nothing
end
update_preconditioner!(P, ignore...) = nothing                                                  # :600
function ldiv!(u′::AbstractVecOrMat{Complex{E}}, P::KPMPreconditioner{E}, u::AbstractVecOrMat{Complex{E}}) where {E}
    GC.@preserve u′ u check(ccall((:sq_kpm_ldiv, LIB), Cint, (Ptr{Cvoid}, Ptr{Complex{E}}, Ptr{Complex{E}}), P.h, u′, u))
    return nothing
end

_kpm_handle(P::KPMPreconditioner) = P.h
_kpm_handle(::UniformScaling) = C_NULL

# ldiv!(x, fdm, b; preconditioner, rng, maxiter, tol) -> (iters, ϵ): src/FermionDetMatrix.jl:248-288
function ldiv!(x::AbstractVecOrMat{Complex{E}}, f::FermionDetMatrix{T,E}, b::AbstractVecOrMat{Complex{E}};
               preconditioner = I, rng::AbstractRNG = Random.default_rng(), maxiter::Int = f.cgs.maxiter, tol::E = f.cgs.tol) where {T,E}
    start = preconditioner isa KPMPreconditioner ? randn(rng, E, f.N) : E[]
    iters = Ref{Int64}(0); ϵ = Ref{Cdouble}(0)
    GC.@preserve x b start check(ccall((:sq_fdm_cg, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Complex{E}}, Ptr{Complex{E}}, Cint, Ptr{Cvoid}, Cint, Ptr{E}, Cdouble, Int64, Ref{Int64}, Ref{Cdouble}),
        f.h, x, b, x === b, _kpm_handle(preconditioner), preconditioner isa KPMPreconditioner, isempty(start) ? C_NULL : pointer(start),
        tol, maxiter, iters, ϵ))
    return Int(iters[]), ϵ[]
end
ldiv!(f::FermionDetMatrix, v::AbstractVecOrMat; kw...) = ldiv!(v, f, v; kw...)

# ---- electron-phonon tables: the fields of SmoQyDQMC.ElectronPhononParameters the path reads ---------------------
mutable struct B200ElPh
    h::Ptr{Cvoid}
    bare_set::Bool          # bare on-site energies / hoppings uploaded (needs the FermionPathIntegral: see _ensure_bare!)
end
const _ELPH_OF_FDM = Dict{Ptr{Cvoid}, WeakRef}()     # operator handle -> its device-side tables (update_chemical_potential! has no other route to them)
# Only what ElectronPhononParameters itself holds is needed here, so PFFCalculator(elph, fdm) keeps the reference's two positional
# arguments (src/PFFCalculator.jl:30-33).  Models the device path does not implement fail HERE, loudly, instead of sampling a wrong action.
function B200ElPh(elph::ElectronPhononParameters{T,E}, f::FermionDetMatrix{T,E}) where {T,E}
    T <: Real || error("libsmoqyelph_b200 supports real hoppings / couplings only (T = $T)")
    ph = elph.phonon_parameters; hol = elph.holstein_parameters_up; ssh = elph.ssh_parameters_up
    disp = elph.dispersion_parameters
    (all(iszero, imag.(ssh.α)) && all(iszero, imag.(ssh.α2)) && all(iszero, imag.(ssh.α3)) && all(iszero, imag.(ssh.α4))) ||
        error("libsmoqyelph_b200: complex SSH couplings are not implemented")
    (elph.holstein_parameters_up === elph.holstein_parameters_dn || elph.holstein_parameters_up.α == elph.holstein_parameters_dn.α) ||
        error("libsmoqyelph_b200: spin-dependent Holstein couplings are not implemented")
    nun = hol.nholstein == 0 ? 1 : hol.Nholstein ÷ hol.nholstein
    phsym = Int32[hol.ph_sym_form[(c - 1) ÷ nun + 1] for c in 1:hol.Nholstein]       # expanded per coupling
    hop_of = zeros(Int64, ssh.Nssh)                                                  # inverse of hopping_to_couplings
    for (hop, cs) in enumerate(ssh.hopping_to_couplings), c in cs; hop_of[c] = hop; end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sq_elph_create, LIB), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Cdouble, Int64, Ptr{E}, Ptr{E}, Ptr{E}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{E}, Ptr{E}, Ptr{E}, Ptr{E},
         Ptr{Int32}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{E}, Ptr{E}, Ptr{E}, Ptr{E}, Ptr{E}, Ptr{E}),
        h, f.h, elph.Δτ, length(ph.Ω), ph.Ω, ph.Ω4, ph.M, hol.Nholstein, Vector{Int64}(hol.coupling_to_phonon), Vector{Int64}(hol.coupling_to_site),
        hol.α, hol.α2, hol.α3, hol.α4, phsym, ssh.Nssh, Matrix{Int64}(ssh.coupling_to_phonon), hop_of, real.(ssh.α), real.(ssh.α2),
        real.(ssh.α3), real.(ssh.α4), C_NULL, C_NULL))
    e = B200ElPh(h[], false)
    finalizer(x -> ccall((:sq_elph_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), e)
    if disp.Ndispersion > 0      # dispersive couplings: bosonic action + eval_derivative_dispersive_action! on the device (src/EFAPFFHMCUpdater.jl:193)
        check(ccall((:sq_elph_set_dispersion, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{E}, Ptr{E}),
                    e.h, disp.Ndispersion, Matrix{Int64}(disp.dispersion_to_phonon), Vector{E}(disp.Ω), Vector{E}(disp.Ω4)))
    end
    _ELPH_OF_FDM[f.h] = WeakRef(e)
    return e
end
# The bare on-site energies (eps - mu) and hoppings are what is left of the path integral once the phonon contribution is taken out --
# the very operation the reference's trajectory performs (`SmoQyDQMC.update!(fpi, elph, x, -1)`, src/EFAPFFHMCUpdater.jl:198).  Done once,
# by the first call that carries the FermionPathIntegral (hmc_update!, the global moves, update_chemical_potential!).
function _ensure_bare!(e::B200ElPh, elph::ElectronPhononParameters{T,E}, fpi::FermionPathIntegral{T,E}) where {T,E}
    e.bare_set && return nothing
    all(iszero, imag.(fpi.t)) || error("libsmoqyelph_b200 supports real hoppings only (SURVEY.md 9 Q9)")
    SmoQyDQMC.update!(fpi, elph, elph.x, -1)
    V0 = Vector{E}(fpi.V[:, 1]); t0 = Vector{E}(real.(fpi.t[:, 1]))
    SmoQyDQMC.update!(fpi, elph, elph.x, +1)
    check(ccall((:sq_elph_set_bare, LIB), Cint, (Ptr{Cvoid}, Ptr{E}, Ptr{E}), e.h, V0, t0))
    e.bare_set = true
    return nothing
end
push_x!(e::B200ElPh, x::Matrix{Float64}) = check(ccall((:sq_elph_set_x, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.h, x))
pull_x!(x::Matrix{Float64}, e::B200ElPh) = check(ccall((:sq_elph_get_x, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), e.h, x))

# ---- PFFCalculator (src/PFFCalculator.jl:9-158) ------------------------------------------------------------------
mutable struct PFFCalculator{E<:AbstractFloat}
    h::Ptr{Cvoid}; elph::B200ElPh
end
function PFFCalculator(elph::ElectronPhononParameters{T,E}, f::FermionDetMatrix{T,E}) where {T,E}      # src/PFFCalculator.jl:30-33
    e = B200ElPh(elph, f)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sq_pff_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}), h, e.h))
    p = PFFCalculator{E}(h[], e)
    finalizer(x -> ccall((:sq_pff_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
    return p
end
# The three methods below take the reference's positional arguments (src/PFFCalculator.jl:56-158).  As there, the operator is whatever the
# caller last set with update!(fdm, fpi); only x travels (it enters Λ and the coupling derivatives).
# sample_pseudofermion_fields!(pff, elph, fdm, rng) -> Sf: the randn!(rng, Φ) draw is made here so that the caller's rng determines Φ.
function sample_pseudofermion_fields!(p::PFFCalculator{E}, elph, f::FermionDetMatrix, rng::AbstractRNG = Random.default_rng()) where {E}
    push_x!(p.elph, elph.x)
    R = randn(rng, Complex{E}, f.Lτ, f.N)
    Sf = Ref{Cdouble}(0)
    GC.@preserve R check(ccall((:sq_pff_sample, LIB), Cint, (Ptr{Cvoid}, Ptr{Complex{E}}, Ref{Cdouble}), p.h, R, Sf))
    return Sf[]
end
function calculate_fermionic_action!(p::PFFCalculator{E}, elph, f, preconditioner, rng::AbstractRNG, tol::E = f.cgs.tol,
                                     maxiter::Int = f.cgs.maxiter) where {E}
    push_x!(p.elph, elph.x)
    start = preconditioner isa KPMPreconditioner ? randn(rng, E, f.N) : E[]
    Sf = Ref{Cdouble}(0); it = Ref{Int64}(0); ϵ = Ref{Cdouble}(0)
    GC.@preserve start check(ccall((:sq_pff_action, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{E}, Cdouble, Int64, Ref{Cdouble}, Ref{Int64}, Ref{Cdouble}),
                p.h, _kpm_handle(preconditioner), isempty(start) ? C_NULL : pointer(start), tol, maxiter, Sf, it, ϵ))
    return Sf[], Int(it[]), ϵ[]
end
function calculate_derivative_fermionic_action!(∂Sf∂x::AbstractMatrix{E}, p::PFFCalculator{E}, elph, f, preconditioner, rng::AbstractRNG,
                                                tol::E = f.cgs.tol, maxiter::Int = f.cgs.maxiter) where {E}
    push_x!(p.elph, elph.x)
    start = preconditioner isa KPMPreconditioner ? randn(rng, E, f.N) : E[]
    Sf = Ref{Cdouble}(0); it = Ref{Int64}(0); ϵ = Ref{Cdouble}(0)
    GC.@preserve ∂Sf∂x start check(ccall((:sq_pff_force, LIB), Cint, (Ptr{Cvoid}, Ptr{E}, Ptr{Cvoid}, Ptr{E}, Cdouble, Int64, Ref{Cdouble}, Ref{Int64}, Ref{Cdouble}),
                p.h, ∂Sf∂x, _kpm_handle(preconditioner), isempty(start) ? C_NULL : pointer(start), tol, maxiter, Sf, it, ϵ))
    return Sf[], Int(it[]), ϵ[]
end

# ---- EFAPFFHMCUpdater (src/EFAPFFHMCUpdater.jl:9-279) -------------------------------------------------------------
mutable struct EFAPFFHMCUpdater{E<:AbstractFloat}
    Nt::Int; Δt::E; δ::E; η::E
    h::Ptr{Cvoid}           # created lazily once the PFFCalculator is known (hmc_update! receives it as a keyword)
end
EFAPFFHMCUpdater(; electron_phonon_parameters::ElectronPhononParameters{T,E}, Nt::Int, Δt::E = π / (2 * Nt), η::E = 0.0,
                 δ::E = 0.05) where {T,E} = EFAPFFHMCUpdater{E}(Nt, Δt, δ, η, C_NULL)

# hmc_update!(elph, updater; ...) -> (accepted, iters_avg): the whole trajectory runs on the device.  x is pushed before
# and pulled after; recenter! other than `identity` is not supported inside the device-resident trajectory.
function hmc_update!(elph::ElectronPhononParameters{T,E}, u::EFAPFFHMCUpdater{E}; fermion_path_integral::FermionPathIntegral{T,E},
                     fermion_det_matrix::FermionDetMatrix{T,E}, pff_calculator::PFFCalculator{E}, rng::AbstractRNG,
                     recenter!::Function = identity, Nt::Int = u.Nt, Δt::E = u.Δt, δ::E = u.δ,
                     tol_action::E = fermion_det_matrix.cgs.tol, tol_force::E = sqrt(fermion_det_matrix.cgs.tol),
                     maxiter::Int = fermion_det_matrix.cgs.maxiter, preconditioner = I) where {T,E}
    recenter! === identity || error("recenter! callbacks are not supported by the device-resident trajectory")
    if u.h == C_NULL || Nt != u.Nt || Δt != u.Δt || δ != u.δ
        u.h == C_NULL || ccall((:sq_hmc_destroy, LIB), Cint, (Ptr{Cvoid},), u.h)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:sq_hmc_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, Cdouble, Cdouble, Cdouble, UInt64),
                    h, pff_calculator.h, Nt, Δt, u.η, δ, UInt64(0)))
        u.h = h[]; u.Nt = Nt; u.Δt = Δt; u.δ = δ
    end
    # every trajectory draws its device-side randoms (Φ, momenta, Lanczos starts, the two uniforms) from a Philox stream keyed by
    # ONE draw from the caller's rng: the rng state passed in determines the trajectory, call after call
    check(ccall((:sq_hmc_set_seed, LIB), Cint, (Ptr{Cvoid}, UInt64), u.h, rand(rng, UInt64)))
    _ensure_bare!(pff_calculator.elph, elph, fermion_path_integral)
    push_x!(pff_calculator.elph, elph.x)
    check(ccall((:sq_elph_refresh_fdm, LIB), Cint, (Ptr{Cvoid},), pff_calculator.elph.h))
    acc = Ref{Cint}(0); info = zeros(Cdouble, 8)
    x0 = copy(elph.x)
    check(ccall((:sq_hmc_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Int64, Ptr{Cdouble}, Int64, Ref{Cint}, Ptr{Cdouble}),
                u.h, _kpm_handle(preconditioner), tol_action, tol_force, maxiter, C_NULL, 0, acc, info))
    if acc[] != 0
        pull_x!(elph.x, pff_calculator.elph)
        SmoQyDQMC.update!(fermion_path_integral, elph, elph.x, x0)     # keep the host-side path integral consistent
    end
    return acc[] != 0, info[1]
end


# ---- global moves (src/reflection_update.jl, src/swap_update.jl, src/radial_update.jl) ---------------------------------
# The mode sampling and the Metropolis test stay in Julia (SmoQyDQMC._sample_phonon_mode*, rand(rng)); the x-mutation, the
# operator refresh and the two action evaluations run on the device.  Rejection restores the device copy of x exactly.
function _global_move!(elph, p::PFFCalculator{E}, mutate_dev!::Function, mutate_host!::Function, logJ::E; fermion_path_integral,
                       fermion_det_matrix, rng, preconditioner, tol::E, maxiter::Int) where {E}
    e = p.elph
    _ensure_bare!(e, elph, fermion_path_integral)
    push_x!(e, elph.x)
    check(ccall((:sq_elph_refresh_fdm, LIB), Cint, (Ptr{Cvoid},), e.h))
    Sf = sample_pseudofermion_fields!(p, elph, fermion_det_matrix, rng)
    Sb = Ref{Cdouble}(0); check(ccall((:sq_elph_bosonic_action, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), e.h, Sb))
    check(ccall((:sq_elph_backup_x, LIB), Cint, (Ptr{Cvoid},), e.h))
    mutate_dev!(e)
    check(ccall((:sq_elph_refresh_fdm, LIB), Cint, (Ptr{Cvoid},), e.h))
    P = 0.0; iters = 0
    try
        Sf′, iters, ϵ = calculate_fermionic_action!(p, elph, fermion_det_matrix, preconditioner, rng, tol, maxiter)
        Sb′ = Ref{Cdouble}(0); check(ccall((:sq_elph_bosonic_action, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), e.h, Sb′))
        P = min(1.0, exp(-((Sf′ + Sb′[]) - (Sf + Sb[])) + logJ))
    catch err
        err isa B200NumericalInstability || rethrow()       # CUDA / argument / watchdog errors are not instabilities
        @warn "Failed to evaluate the fermionic action for the proposed state, update rejected." exception = err
    end
    if rand(rng) < P
        x0 = copy(elph.x)
        mutate_host!(elph.x)
        SmoQyDQMC.update!(fermion_path_integral, elph, elph.x, x0)
        return true, iters
    end
    check(ccall((:sq_elph_restore_x, LIB), Cint, (Ptr{Cvoid},), e.h))
    check(ccall((:sq_elph_refresh_fdm, LIB), Cint, (Ptr{Cvoid},), e.h))
    return false, iters
end

function reflection_update!(elph::ElectronPhononParameters{T,E}, p::PFFCalculator{E}; fermion_path_integral::FermionPathIntegral{T,E},
                            fermion_det_matrix::FermionDetMatrix{T,E}, rng::AbstractRNG, preconditioner = I,
                            tol::E = fermion_det_matrix.cgs.tol, maxiter::Int = fermion_det_matrix.cgs.maxiter, phonon_types = nothing) where {T,E}
    pp = elph.phonon_parameters
    mode = SmoQyDQMC._sample_phonon_mode(rng, pp.nphonon, pp.Nphonon ÷ pp.nphonon, pp.M, phonon_types)
    _global_move!(elph, p, e -> check(ccall((:sq_elph_scale_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), e.h, mode, mode, -1.0)),
                  x -> (@views @. x[mode, :] = -x[mode, :]), zero(E); fermion_path_integral, fermion_det_matrix, rng, preconditioner, tol, maxiter)
end

function swap_update!(elph::ElectronPhononParameters{T,E}, p::PFFCalculator{E}; fermion_path_integral::FermionPathIntegral{T,E},
                      fermion_det_matrix::FermionDetMatrix{T,E}, rng::AbstractRNG, preconditioner = I,
                      tol::E = fermion_det_matrix.cgs.tol, maxiter::Int = fermion_det_matrix.cgs.maxiter, phonon_type_pairs = nothing) where {T,E}
    pp = elph.phonon_parameters
    i, j = SmoQyDQMC._sample_phonon_mode_pair(rng, pp.nphonon, pp.Nphonon ÷ pp.nphonon, pp.M, phonon_type_pairs)
    _global_move!(elph, p, e -> check(ccall((:sq_elph_swap_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int64), e.h, i, j)),
                  x -> SmoQyDQMC.swap!(view(x, i, :), view(x, j, :)), zero(E); fermion_path_integral, fermion_det_matrix, rng, preconditioner, tol, maxiter)
end

function radial_update!(elph::ElectronPhononParameters{T,E}, p::PFFCalculator{E}; fermion_path_integral::FermionPathIntegral{T,E},
                        fermion_det_matrix::FermionDetMatrix{T,E}, rng::AbstractRNG, preconditioner = I,
                        tol::E = fermion_det_matrix.cgs.tol, maxiter::Int = fermion_det_matrix.cgs.maxiter, phonon_id = nothing, σ::E = 1.0) where {T,E}
    pp = elph.phonon_parameters
    Nc = pp.Nphonon ÷ pp.nphonon
    first, last = isnothing(phonon_id) ? (1, pp.Nphonon) : ((phonon_id - 1) * Nc + 1, phonon_id * Nc)
    d = count(isfinite, view(pp.M, first:last)) * fermion_path_integral.Lτ
    γ = randn(rng) * σ / sqrt(d)
    _global_move!(elph, p, e -> check(ccall((:sq_elph_scale_x, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cdouble), e.h, first, last, exp(γ))),
                  x -> (@views @. x[first:last, :] = exp(γ) * x[first:last, :]), d * γ; fermion_path_integral, fermion_det_matrix, rng, preconditioner, tol, maxiter)
end

# ---- GreensEstimator solves + update_chemical_potential! (src/Measurements/GreensEstimator.jl:63-175,
#      src/update_chemical_potential.jl:21-73).  MuTuner stays in Julia. ---------------------------------------------
mutable struct GreensEstimator{E}
    h::Ptr{Cvoid}; Nrv::Int
    n::Int; L::Tuple; Lτ::Int; N::Int      # orbitals per cell, lattice extents, time slices, unit cells (the reference's fields of the same names)
end
function GreensEstimator(f::FermionDetMatrix{T,E}, model_geometry; Nrv::Int = 10, preconditioner = I, rng::AbstractRNG = Random.default_rng(),
                         maxiter::Int = f.cgs.maxiter, tol::E = f.cgs.tol) where {T,E}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sq_greens_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int64, UInt64), h, f.h, Nrv, rand(rng, UInt64)))
    n = model_geometry.unit_cell.n; L = Tuple(model_geometry.lattice.L)
    g = GreensEstimator{E}(h[], Nrv, n, L, f.Lτ, f.N ÷ n)
    finalizer(x -> ccall((:sq_greens_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), g)
    update_greens_estimator!(g, f; preconditioner, rng, maxiter, tol)
    return g
end
function update_greens_estimator!(g::GreensEstimator{E}, f::FermionDetMatrix; preconditioner = I, rng = Random.default_rng(),
                                  maxiter::Int, tol::E) where {E}
    avg = Ref{Cdouble}(0)
    check(ccall((:sq_greens_set_seed, LIB), Cint, (Ptr{Cvoid}, UInt64), g.h, rand(rng, UInt64)))     # the caller's rng keys the random vectors
    check(ccall((:sq_greens_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ref{Cdouble}),
                g.h, _kpm_handle(preconditioner), C_NULL, tol, maxiter, avg))
    return avg[]
end
function _measure(g::GreensEstimator)
    n = zeros(ComplexF64, 1); d = zeros(ComplexF64, 1); N2 = zeros(ComplexF64, 1)
    check(ccall((:sq_greens_measure, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}), g.h, n, d, N2))
    return n[1], d[1], N2[1]
end
# measure_GΔ0!(correlation, g, (a, b)) (src/Measurements/GreensEstimator.jl:177-233): the translation-averaged time-displaced Green's
# function is evaluated on the device (aperiodic extension, (D+1)-dimensional FFT cross-correlation, average over the random vectors)
# and added to `correlation` with the imaginary-time axis last, as add_contraction_to_correlation! does (:718-729).
function measure_GΔ0!(correlation::AbstractArray{Complex{E}}, g::GreensEstimator{E}, orbitals::NTuple{2,Int}; n::Int = g.n, L::NTuple{D,Int} = g.L) where {E,D}
    Lτ = size(correlation, D + 1) - 1
    GΔ0 = zeros(Complex{E}, Lτ + 1, L...)
    dims = collect(Int64, L)
    GC.@preserve GΔ0 dims check(ccall((:sq_greens_measure_GD0, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Cint, Cint, Ptr{Complex{E}}),
                                      g.h, n, D, dims, orbitals[1], orbitals[2], GΔ0))
    @. correlation += $PermutedDimsArray(GΔ0, (2:D+1..., 1))
    return nothing
end
# The four-point contractions (src/Measurements/GreensEstimator.jl:236-606), with or without hopping weights; the correlation functions built on
# them (density.jl, pair.jl, spin.jl) call these exactly as in the reference.
function _contraction!(correlation::AbstractArray{Complex{E}}, g::GreensEstimator{E}, kind::Int, orbitals::NTuple{4,Int}, r1, r2, r3, r4, coef,
                       tΔ = nothing, t0 = nothing, conj_tΔ::Bool = false, conj_t0::Bool = false; n::Int = g.n, L::NTuple{D,Int} = g.L) where {E,D}
    Lτ = size(correlation, D + 1) - 1
    C = zeros(Complex{E}, Lτ + 1, L...)
    dims = collect(Int64, L); orb = collect(Cint, orbitals); r = Int64[r1..., r2..., r3..., r4...]
    if isnothing(tΔ) && isnothing(t0)
        GC.@preserve C dims orb r check(ccall((:sq_greens_measure_contraction, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Int64}, Ptr{Complex{E}}), g.h, kind, n, D, dims, orb, r, C))
    else
        # real hoppings only (conj_tΔ / conj_t0 are no-ops); the weights are materialised as dense (Lτ, L...) arrays
        wΔ = isnothing(tΔ) ? E[] : collect(E, tΔ); w0 = isnothing(t0) ? E[] : collect(E, t0)
        GC.@preserve C dims orb r wΔ w0 check(ccall((:sq_greens_measure_contraction_weighted, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Ptr{Cint}, Ptr{Int64}, Ptr{E}, Ptr{E}, Ptr{Complex{E}}), g.h, kind, n, D, dims, orb, r,
            isnothing(tΔ) ? C_NULL : pointer(wΔ), isnothing(t0) ? C_NULL : pointer(w0), C))
    end
    @. correlation += coef * $PermutedDimsArray(C, (2:D+1..., 1))
    return nothing
end
measure_GΔ0_GΔ0!(corr, g::GreensEstimator, orbitals, r1, r2, r3, r4, coef, t...; kw...) = _contraction!(corr, g, 0, orbitals, r1, r2, r3, r4, coef, t...; kw...)
measure_GΔΔ_G00!(corr, g::GreensEstimator, orbitals, r1, r2, r3, r4, coef, t...; kw...) = _contraction!(corr, g, 1, orbitals, r1, r2, r3, r4, coef, t...; kw...)
measure_G0Δ_GΔ0!(corr, g::GreensEstimator, orbitals, r1, r2, r3, r4, coef, t...; kw...) = _contraction!(corr, g, 2, orbitals, r1, r2, r3, r4, coef, t...; kw...)
# measure_current_correlation! (src/Measurements/Correlations/current.jl:2-151) in terms of the weighted contractions; a bond is given as
# (orbitals = (b, a), displacement); σ = nothing: the spin-summed form
function measure_current_correlation!(CC, g::GreensEstimator, b′, b″, t′, t″, σ = nothing, coef = 1.0; kw...)
    (b, a), r′ = b′; (d, c), r″ = b″
    z = ntuple(_ -> 0, length(r′))
    f1, f2 = isnothing(σ) ? (4.0, 2.0) : (1.0, σ[1] == σ[2] ? 1.0 : 0.0)
    measure_GΔΔ_G00!(CC, g, (a, b, d, c), r′, z, z, r″, +f1 * coef, t′, t″, true, false; kw...)
    measure_GΔΔ_G00!(CC, g, (a, b, c, d), r′, z, r″, z, -f1 * coef, t′, t″, true, true; kw...)
    measure_GΔΔ_G00!(CC, g, (b, a, d, c), z, r′, z, r″, -f1 * coef, t′, t″, false, false; kw...)
    measure_GΔΔ_G00!(CC, g, (b, a, c, d), z, r′, r″, z, +f1 * coef, t′, t″, false, true; kw...)
    f2 == 0 && return nothing
    measure_G0Δ_GΔ0!(CC, g, (b, a, c, d), z, z, r′, r″, -f2 * coef, t′, t″, true, false; kw...)
    measure_G0Δ_GΔ0!(CC, g, (b, a, d, c), r″, z, r′, z, +f2 * coef, t′, t″, true, true; kw...)
    measure_G0Δ_GΔ0!(CC, g, (d, a, b, c), z, r′, z, r″, +f2 * coef, t′, t″, false, false; kw...)
    measure_G0Δ_GΔ0!(CC, g, (c, a, b, d), r″, r′, z, z, -f2 * coef, t′, t″, false, true; kw...)
    return nothing
end
function measure_n(g::GreensEstimator{E}, orbital::Int; n::Int = g.n) where {E}
    out = zeros(Complex{E}, 1)
    check(ccall((:sq_greens_measure_n_orbital, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Complex{E}}), g.h, n, orbital, out))
    return out[1]
end
function measure_double_occ(g::GreensEstimator{E}, orbital::Int; n::Int = g.n) where {E}      # scalar_measurements.jl:98-109
    out = zeros(Complex{E}, 1)
    check(ccall((:sq_greens_measure_double_occ_orbital, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Complex{E}}), g.h, n, orbital, out))
    return out[1]
end
# make_measurements!(measurement_container, fdm, greens_estimator; ...) -> iters  (src/Measurements/make_measurements.jl:19-90), same
# keywords as the reference.  The estimator refresh runs on the device; the global and the local measurements (:92-146, tight-binding and
# electron-phonon energies, tight_binding_measurements.jl:2-133 / electron_phonon_measurements.jl) are accumulated here from this module's
# measure_* methods.  The correlation bodies (:161-913) walk SmoQyDQMC's container and call measure_*_correlation! -- host code of the
# reference package that works unchanged on this GreensEstimator; the integrating package registers it once:
#     SmoQyElPhB200.CORRELATION_MEASUREMENTS[] = (mc, g, mg, tbp, fpi) -> begin
#         SmoQyElPhQMC.make_correlation_measurements!(mc, g, mg, tbp, fpi)
#         SmoQyElPhQMC.make_composite_correlation_measurements!(mc, g, mg, tbp, fpi)
#     end
# A container that requests correlations while no such function is registered is an ERROR, never a silent no-op.
const CORRELATION_MEASUREMENTS = Ref{Union{Nothing,Function}}(nothing)
const PHONON_GREENS_MEASUREMENTS = Ref{Union{Nothing,Function}}(nothing)     # make_phonon_greens_measurements! (pure host code on x)

function make_global_measurements!(gm::AbstractDict, tight_binding_parameters, electron_phonon_parameters, g::GreensEstimator)   # :92-117
    gm["sgn"] += 1.0
    for k in ("sgndetGup", "sgndetGdn", "logdetGup", "logdetGdn", "action_fermionic", "action_total"); gm[k] = NaN; end
    gm["action_bosonic"] += SmoQyDQMC.bosonic_action(electron_phonon_parameters)
    density, docc, N2 = _measure(g)
    gm["density_up"] += density; gm["density_dn"] += density; gm["density"] += 2 * density
    gm["double_occ"] += docc; gm["Nsqrd"] += N2
    gm["chemical_potential"] += tight_binding_parameters.μ
    return nothing
end
function make_local_measurements!(lm::AbstractDict, model_geometry, tight_binding_parameters, electron_phonon_parameters, fermion_path_integral,
                                  g::GreensEstimator)                                                                           # :120-146
    for n in 1:g.n
        density = measure_n(g, n)
        lm["density_up"][n] += density; lm["density_dn"][n] += density; lm["density"][n] += 2 * density
        lm["double_occ"][n] += measure_double_occ(g, n)
    end
    # make_tight_binding_measurements! (tight_binding_measurements.jl:2-40).  The reference adds the modulated hopping energy under the
    # bare_hopping_energy keys as well (SURVEY.md Q8, not propagated: each value goes to its own key).
    for n in 1:g.n
        e = measure_onsite_energy(g, tight_binding_parameters, n)
        lm["onsite_energy_up"][n] += e; lm["onsite_energy_dn"][n] += e; lm["onsite_energy"][n] += 2 * e
    end
    for hid in 1:length(tight_binding_parameters.bond_ids)
        e0 = measure_bare_hopping_energy(g, tight_binding_parameters, model_geometry, hid)
        lm["bare_hopping_energy_up"][hid] += e0; lm["bare_hopping_energy_dn"][hid] += e0; lm["bare_hopping_energy"][hid] += 2 * e0
        e1 = measure_hopping_energy(g, tight_binding_parameters, fermion_path_integral, hid)
        lm["hopping_energy_up"][hid] += e1; lm["hopping_energy_dn"][hid] += e1; lm["hopping_energy"][hid] += 2 * e1
    end
    # make_electron_phonon_measurements! (electron_phonon_measurements.jl:2-58): the phonon energies and moments are SmoQyDQMC's host
    # arithmetic on x; the electron-phonon energies need the estimator
    x = electron_phonon_parameters.x
    hol = electron_phonon_parameters.holstein_parameters_up; ssh = electron_phonon_parameters.ssh_parameters_up
    ph = electron_phonon_parameters.phonon_parameters
    for id in 1:ph.nphonon
        lm["phonon_kin_energy"][id] += SmoQyDQMC.measure_phonon_kinetic_energy(ph, x, electron_phonon_parameters.Δτ, id)
        lm["phonon_pot_energy"][id] += SmoQyDQMC.measure_phonon_potential_energy(ph, x, id)
        for (key, pw) in (("X", 1), ("X2", 2), ("X3", 3), ("X4", 4))
            lm[key][id] += SmoQyDQMC.measure_phonon_position_moment(ph, x, id, pw)
        end
    end
    for id in 1:hol.nholstein
        e = measure_holstein_energy(hol, g, x, id)
        lm["holstein_energy_up"][id] += e; lm["holstein_energy_dn"][id] += e; lm["holstein_energy"][id] += 2 * e
    end
    for id in 1:ssh.nssh
        e = measure_ssh_energy(ssh, g, x, id)
        lm["ssh_energy_up"][id] += e; lm["ssh_energy_dn"][id] += e; lm["ssh_energy"][id] += 2 * e
    end
    return nothing
end
_requests_correlations(mc) = any(k -> hasproperty(mc, k) && !isempty(getproperty(mc, k)),
    (:equaltime_correlations, :time_displaced_correlations, :integrated_correlations,
     :equaltime_composite_correlations, :time_displaced_composite_correlations, :integrated_composite_correlations))
function make_measurements!(measurement_container::NamedTuple, f::FermionDetMatrix{T,E}, g::GreensEstimator{E}; model_geometry, fermion_path_integral,
                            tight_binding_parameters, electron_phonon_parameters, preconditioner = I, rng::AbstractRNG = Random.default_rng(),
                            tol::E = f.cgs.tol, maxiter::Int = f.cgs.maxiter) where {T,E}
    iters = update_greens_estimator!(g, f; preconditioner, rng, tol, maxiter)
    make_global_measurements!(measurement_container.global_measurements, tight_binding_parameters, electron_phonon_parameters, g)
    make_local_measurements!(measurement_container.local_measurements, model_geometry, tight_binding_parameters, electron_phonon_parameters,
                             fermion_path_integral, g)
    if _requests_correlations(measurement_container)
        fn = CORRELATION_MEASUREMENTS[]
        fn === nothing && error("make_measurements!: the measurement container requests correlation measurements, but no correlation driver is " *
                                "registered (set SmoQyElPhB200.CORRELATION_MEASUREMENTS[], see the comment above make_measurements!)")
        fn(measurement_container, g, model_geometry, tight_binding_parameters, fermion_path_integral)
    end
    pg = PHONON_GREENS_MEASUREMENTS[]
    pg === nothing || pg(measurement_container, model_geometry, electron_phonon_parameters)
    return iters
end
# update_chemical_potential!(fdm, greens_estimator; ...) -> iters  (src/update_chemical_potential.jl:21-73): solves and the two scalar
# measurements on the device, MuTuner.update! in Julia, then the shift V += -μ + μ′ on the host path integral AND on the device tables
function update_chemical_potential!(f::FermionDetMatrix{T,E}, g::GreensEstimator{E}; chemical_potential_tuner, tight_binding_parameters,
                                    fermion_path_integral::FermionPathIntegral{T,E}, preconditioner = I, rng::AbstractRNG = Random.default_rng(),
                                    update_greens_estimator::Bool = true, tol::E = f.cgs.tol, maxiter::Int = f.cgs.maxiter) where {T,E}
    iters = 0
    if update_greens_estimator
        iters = update_greens_estimator!(g, f; preconditioner, rng, maxiter, tol)
    end
    μ′ = tight_binding_parameters.μ
    n = real(2 * measure_n(g))
    Nsqrd = real(measure_Nsqrd(g))
    μ = MuTuner.update!(chemical_potential_tuner, n, Nsqrd, one(E))
    tight_binding_parameters.μ = μ
    V = fermion_path_integral.V
    @. V += -μ + μ′
    ref = get(_ELPH_OF_FDM, f.h, nothing)
    if ref !== nothing && ref.value !== nothing && ref.value.bare_set       # (not yet set: _ensure_bare! will read the shifted V later)
        check(ccall((:sq_elph_shift_mu, LIB), Cint, (Ptr{Cvoid}, Cdouble), ref.value.h, μ - μ′))
    end
    update!(f, fermion_path_integral)
    return iters
end

# local measurements (src/Measurements/tight_binding_measurements.jl:58-133): the weights are assembled here exactly as the reference's
# loops imply, the reductions over R, G R run on the device (sq_greens_weighted_density / sq_greens_weighted_bonds)
function _weighted_density(g::GreensEstimator{E}, w::Matrix{E}) where {E}                  # w: (N sites, Lτ)
    out = zeros(Complex{E}, 1)
    GC.@preserve w check(ccall((:sq_greens_weighted_density, LIB), Cint, (Ptr{Cvoid}, Ptr{E}, Ptr{Complex{E}}), g.h, w, out))
    return out[1]
end
function _weighted_bonds(g::GreensEstimator{E}, bonds::Matrix{Int64}, w::Matrix{Complex{E}}) where {E}     # bonds (2, nb) 1-based, w (nb, Lτ)
    out = zeros(Complex{E}, 1)
    GC.@preserve bonds w check(ccall((:sq_greens_weighted_bonds, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Complex{E}}, Ptr{Complex{E}}),
                                     g.h, size(bonds, 2), bonds, w, out))
    return out[1]
end
function measure_onsite_energy(g::GreensEstimator{E}, tight_binding_parameters, orbital::Int; n::Int = g.n, Lτ::Int = g.Lτ) where {E}
    ϵ = tight_binding_parameters.ϵ; μ = tight_binding_parameters.μ
    Nsites = length(ϵ); Ncells = Nsites ÷ n
    w = zeros(E, Nsites, Lτ)
    for i in orbital:n:Nsites
        w[i, :] .= (ϵ[i] - μ) / (Lτ * Ncells)
    end
    return _weighted_density(g, w)
end
function measure_bare_hopping_energy(g::GreensEstimator{E}, tight_binding_parameters, model_geometry, hopping_id::Int; n::Int = g.n, Lτ::Int = g.Lτ) where {E}
    sl = tight_binding_parameters.bond_slices[hopping_id]
    t = tight_binding_parameters.t[sl]
    bonds = Matrix{Int64}(tight_binding_parameters.neighbor_table[:, sl])
    Nsites = length(tight_binding_parameters.ϵ)
    w = Matrix{Complex{E}}(undef, length(t), Lτ)
    for l in 1:Lτ; w[:, l] .= t ./ (Lτ * Nsites); end
    return _weighted_bonds(g, bonds, w)
end
function measure_hopping_energy(g::GreensEstimator{E}, tight_binding_parameters, fermion_path_integral, hopping_id::Int; n::Int = g.n) where {E}
    sl = tight_binding_parameters.bond_slices[hopping_id]
    t = fermion_path_integral.t[sl, :]                                   # the modulated amplitudes t(bond, τ)
    bonds = Matrix{Int64}(tight_binding_parameters.neighbor_table[:, sl])
    Nsites = length(tight_binding_parameters.ϵ); Lτ = size(t, 2)
    return _weighted_bonds(g, bonds, Matrix{Complex{E}}(t ./ (Lτ * Nsites)))
end
# measure_holstein_energy(holstein_parameters, greens_estimator, x, holstein_id)  (src/Measurements/electron_phonon_measurements.jl:60-122).
# As in the reference the density is taken in the unit cell of the phonon (orbital of the first coupling of this id), the phonons of
# the id are the contiguous rows phonon_i:phonon_f of x.  The cubic coupling multiplies x^3 (the reference has x^2 at :115, a typo that
# SURVEY.md Q8 says not to propagate; the two agree for every shipped model, where α3 = 0).
function measure_holstein_energy(holstein_parameters, g::GreensEstimator{E}, x::Matrix{E}, holstein_id::Int) where {E}
    (; coupling_to_site, coupling_to_phonon, ph_sym_form) = holstein_parameters
    N = g.N; n = g.n; Lτ = g.Lτ
    sl = ((holstein_id - 1) * N + 1):(holstein_id * N)
    phs = ph_sym_form[holstein_id]
    orbital_id = mod1(coupling_to_site[first(sl)], n)
    phonon_i = coupling_to_phonon[first(sl)]
    α1 = holstein_parameters.α[sl]; α2 = holstein_parameters.α2[sl]; α3 = holstein_parameters.α3[sl]; α4 = holstein_parameters.α4[sl]
    w = zeros(E, N * n, Lτ)
    shift = zero(E)
    for l in 1:Lτ, u in 1:N
        xv = x[phonon_i + u - 1, l]
        even = α2[u] * xv^2 + α4[u] * xv^4
        odd = α1[u] * xv + α3[u] * xv^3
        w[orbital_id + n * (u - 1), l] += (even + odd) / (N * Lτ)
        phs && (shift += odd / (2 * N * Lτ))
    end
    return _weighted_density(g, w) - shift
end
# measure_ssh_energy(ssh_parameters, greens_estimator, x, ssh_id)  (src/Measurements/electron_phonon_measurements.jl:124-186)
function measure_ssh_energy(ssh_parameters, g::GreensEstimator{E}, x::Matrix{E}, ssh_id::Int) where {E}
    N = g.N; Lτ = g.Lτ
    sl = ((ssh_id - 1) * N + 1):(ssh_id * N)
    bonds = Matrix{Int64}(ssh_parameters.neighbor_table[:, sl])
    c2p = ssh_parameters.coupling_to_phonon[:, sl]
    α1 = real.(ssh_parameters.α[sl]); α2 = real.(ssh_parameters.α2[sl]); α3 = real.(ssh_parameters.α3[sl]); α4 = real.(ssh_parameters.α4[sl])
    w = Matrix{Complex{E}}(undef, N, Lτ)
    for l in 1:Lτ, u in 1:N
        Δx = x[c2p[2, u], l] - x[c2p[1, u], l]
        w[u, l] = -(α1[u] * Δx + α2[u] * Δx^2 + α3[u] * Δx^3 + α4[u] * Δx^4) / (N * Lτ)
    end
    return _weighted_bonds(g, bonds, w)
end
measure_n(g::GreensEstimator) = _measure(g)[1]
measure_double_occ(g::GreensEstimator) = _measure(g)[2]
measure_Nsqrd(g::GreensEstimator) = _measure(g)[3]

end # module
