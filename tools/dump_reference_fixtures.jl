# dump_reference_fixtures.jl -- pins the parity of this repository against the REAL reference.
#
# Run on any machine with Julia and the reference package installed (not possible in the build container: no `julia`, no depot):
#
#     julia --project=<env with SmoQyDQMC + SmoQyElPhQMC> tools/dump_reference_fixtures.jl tests/golden
#
# For the named configurations 1t (honeycomb Holstein L=3, beta=1), 2 (optical SSH chain N=64, beta=16; pass "small" as second argument
# for beta=2) and 3 (bond SSH square 16x16, beta=10; "small": 8x8, beta=1) it writes one directory tests/golden/ref_<tag>/ with
# raw little-endian column-major arrays + manifest.txt (SURVEY.md 8c):
#     inputs   nt (2 x Nh, permuted), perm, color_lo / color_hi, V (N x Ltau), t (Nh x Ltau), x (Nph x Ltau), Phi_in (Ltau x N),
#              the ElectronPhononParameters tables the force reads, the pseudofermion field Phi and its noise R
#     outputs  MtM_Phi = M^T M Phi_in, M_Phi, Mt_Phi; x_cg = [M^T M]^-1 Phi_in at tol 1e-14, iters at 1e-5 / 1e-10 / 1e-14;
#              Sf_sample, Sf (action at tol 1e-14), dSdx (calculate_derivative_fermionic_action! at tol 1e-14)
# tests/test_reference_fixtures.py feeds the inputs to the CPU oracle (oracle/ref_c.c) and to the CUDA library and compares with the outputs
# (M^T M v, CG solutions, forces to 1e-12 / solver accuracy, iteration counts +-1).  Until such a directory is committed the oracle
# is pinned only by the dense-matrix known answers of tests/test_oracle_kat.py ("parity unpinned", DESIGN.md section 2).
using LinearAlgebra
using Random
using SmoQyDQMC
import SmoQyDQMC.LatticeUtilities as lu
using SmoQyElPhQMC
const EQ = SmoQyElPhQMC

function write_array(dir, manifest, name, A::AbstractArray{T}) where {T}
    B = collect(A)
    open(joinpath(dir, name * ".bin"), "w") do io
        write(io, B)
    end
    tn = T <: Complex ? "c16" : (T <: AbstractFloat ? "f8" : (T <: Bool ? "u1" : "i8"))
    println(manifest, name, " ", tn, " ", join(size(B), " "))
end
write_scalar(manifest, name, v) = println(manifest, name, " scalar ", v)

function dump_config(outdir, tag, model_geometry, tight_binding_model, electron_phonon_model; β, Δτ, seed = 1234)
    rng = Xoshiro(seed)
    dir = joinpath(outdir, "ref_" * tag)
    mkpath(dir)
    tbp = TightBindingParameters(tight_binding_model = tight_binding_model, model_geometry = model_geometry, rng = rng)
    elph = ElectronPhononParameters(β = β, Δτ = Δτ, electron_phonon_model = electron_phonon_model, tight_binding_parameters = tbp,
                                    model_geometry = model_geometry, rng = rng)
    fpi = FermionPathIntegral(tight_binding_parameters = tbp, β = β, Δτ = Δτ)
    initialize!(fpi, elph)
    fdm = SymFermionDetMatrix(fpi, maxiter = 50_000, tol = 1e-10)
    Lτ, N = fpi.Lτ, fpi.N
    open(joinpath(dir, "manifest.txt"), "w") do mf
        write_scalar(mf, "beta", β); write_scalar(mf, "dtau", Δτ); write_scalar(mf, "Ltau", Lτ); write_scalar(mf, "N", N)
        write_scalar(mf, "sym", 1)
        write_array(dir, mf, "neighbor_table", Matrix{Int64}(fpi.neighbor_table))
        write_array(dir, mf, "nt", Matrix{Int64}(fdm.checkerboard_neighbor_table))
        write_array(dir, mf, "perm", Vector{Int64}(fdm.checkerboard_perm))
        write_array(dir, mf, "color_lo", Int64[first(r) for r in fdm.checkerboard_colors])
        write_array(dir, mf, "color_hi", Int64[last(r) for r in fdm.checkerboard_colors])
        write_array(dir, mf, "V", Matrix{Float64}(fpi.V)); write_array(dir, mf, "t", Matrix{Float64}(real.(fpi.t)))
        write_array(dir, mf, "x", Matrix{Float64}(elph.x))
        write_array(dir, mf, "eps_bare", Vector{Float64}(tbp.ϵ)); write_scalar(mf, "mu", tbp.μ)
        write_array(dir, mf, "t_bare", Vector{Float64}(real.(tbp.t)))
        write_array(dir, mf, "expV", Matrix{Float64}(fdm.expnΔτV)); write_array(dir, mf, "cosh_t", Matrix{Float64}(fdm.coshΔτt))
        write_array(dir, mf, "sinh_t", Matrix{Float64}(real.(fdm.sinhΔτt)))
        ph = elph.phonon_parameters; hol = elph.holstein_parameters_up; ssh = elph.ssh_parameters_up
        write_array(dir, mf, "Omega", Vector{Float64}(ph.Ω)); write_array(dir, mf, "Omega4", Vector{Float64}(ph.Ω4)); write_array(dir, mf, "M", Vector{Float64}(ph.M))
        write_scalar(mf, "nphonon", ph.nphonon)
        write_array(dir, mf, "hol_phonon", Vector{Int64}(hol.coupling_to_phonon)); write_array(dir, mf, "hol_site", Vector{Int64}(hol.coupling_to_site))
        write_array(dir, mf, "hol_alpha", hcat(Vector{Float64}(hol.α), Vector{Float64}(hol.α2), Vector{Float64}(hol.α3), Vector{Float64}(hol.α4)))
        nun = hol.nholstein == 0 ? 1 : hol.Nholstein ÷ hol.nholstein
        write_array(dir, mf, "hol_phsym", Int64[hol.ph_sym_form[(c - 1) ÷ nun + 1] for c in 1:hol.Nholstein])
        write_array(dir, mf, "ssh_phonon", Matrix{Int64}(ssh.coupling_to_phonon))
        hop_of = zeros(Int64, ssh.Nssh)
        for (hop, cs) in enumerate(ssh.hopping_to_couplings), c in cs; hop_of[c] = hop; end
        write_array(dir, mf, "ssh_hopping", hop_of)
        write_array(dir, mf, "ssh_alpha", hcat(Vector{Float64}(real.(ssh.α)), Vector{Float64}(real.(ssh.α2)), Vector{Float64}(real.(ssh.α3)), Vector{Float64}(real.(ssh.α4))))
        # ---- products
        Φin = randn(rng, ComplexF64, Lτ, N)
        out = similar(Φin)
        write_array(dir, mf, "Phi_in", Φin)
        EQ.mul_MtM!(out, fdm, Φin); write_array(dir, mf, "MtM_Phi", out)
        EQ.mul_M!(out, fdm, Φin); write_array(dir, mf, "M_Phi", out)
        EQ.mul_Mt!(out, fdm, Φin); write_array(dir, mf, "Mt_Phi", out)
        # ---- CG (no preconditioner), zero start (x === b)
        for (tol, nm) in ((1e-5, "5"), (1e-10, "10"), (1e-14, "14"))
            xs = copy(Φin)
            iters, ϵ = ldiv!(xs, fdm, xs, preconditioner = I, rng = rng, maxiter = 50_000, tol = tol)
            write_scalar(mf, "cg_iters_" * nm, iters); write_scalar(mf, "cg_eps_" * nm, ϵ)
            tol == 1e-14 && write_array(dir, mf, "x_cg", xs)
        end
        # ---- pseudofermion action and force
        pff = PFFCalculator(elph, fdm)
        rng2 = Xoshiro(seed + 1)
        R = randn(copy(rng2), ComplexF64, Lτ, N)                  # the very draw sample_pseudofermion_fields! makes (randn!(rng, Φ))
        Sf0 = EQ.sample_pseudofermion_fields!(pff, elph, fdm, rng2)
        write_array(dir, mf, "R", R); write_array(dir, mf, "Phi", pff.Φ); write_scalar(mf, "Sf_sample", Sf0)
        ∂S∂x = zeros(Float64, size(elph.x))
        Sf, iters, ϵ = EQ.calculate_derivative_fermionic_action!(∂S∂x, pff, elph, fdm, I, rng, 1e-14, 50_000)
        write_array(dir, mf, "dSdx", ∂S∂x); write_scalar(mf, "Sf", Sf); write_scalar(mf, "force_iters", iters)
        write_array(dir, mf, "Lambda", Matrix{Float64}(pff.Λ))
    end
    println("wrote ", dir)
end

# ---- config 1t: tutorials/holstein_honeycomb.jl, L = 3, beta = 1 ------------------------------------------------------------
function holstein_honeycomb(outdir; L = 3, β = 1.0, Δτ = 0.05, Ω = 1.0, α = 1.5, μ = 0.0)
    a1 = [+3/2, +√3/2]; a2 = [+3/2, -√3/2]; r1 = [0.0, 0.0]; r2 = [1.0, 0.0]
    unit_cell = lu.UnitCell(lattice_vecs = [a1, a2], basis_vecs = [r1, r2])
    lattice = lu.Lattice(L = [L, L], periodic = [true, true])
    mg = ModelGeometry(unit_cell, lattice)
    bonds = [lu.Bond(orbitals = (1, 2), displacement = d) for d in ([0, 0], [-1, 0], [0, -1])]
    foreach(b -> add_bond!(mg, b), bonds)
    tbm = TightBindingModel(model_geometry = mg, t_bonds = bonds, t_mean = [1.0, 1.0, 1.0], μ = μ, ϵ_mean = [0.0, 0.0])
    epm = ElectronPhononModel(model_geometry = mg, tight_binding_model = tbm)
    for (orb, r) in ((1, r1), (2, r2))
        pid = add_phonon_mode!(electron_phonon_model = epm, phonon_mode = PhononMode(basis_vec = r, Ω_mean = Ω))
        hc = HolsteinCoupling(model_geometry = mg, phonon_id = pid, orbital_id = orb, displacement = [0, 0], α_mean = α, ph_sym_form = true)
        add_holstein_coupling!(electron_phonon_model = epm, holstein_coupling = hc, model_geometry = mg)
    end
    dump_config(outdir, "cfg1t", mg, tbm, epm; β, Δτ)
end

# ---- config 2: examples/ossh_chain.jl ------------------------------------------------------------------------------------------
function ossh_chain(outdir; L = 64, β = 16.0, Δτ = 0.05, Ω = 1.0, α = 0.5, μ = 0.0, tag = "cfg2")
    unit_cell = lu.UnitCell(lattice_vecs = [[1.0]], basis_vecs = [[0.0]])
    lattice = lu.Lattice(L = [L], periodic = [true])
    mg = ModelGeometry(unit_cell, lattice)
    bond = lu.Bond(orbitals = (1, 1), displacement = [1])
    add_bond!(mg, bond)
    tbm = TightBindingModel(model_geometry = mg, t_bonds = [bond], t_mean = [1.0], μ = μ, ϵ_mean = [0.0])
    epm = ElectronPhononModel(model_geometry = mg, tight_binding_model = tbm)
    pid = add_phonon_mode!(electron_phonon_model = epm, phonon_mode = PhononMode(basis_vec = [0.0], Ω_mean = Ω))
    sc = SSHCoupling(model_geometry = mg, tight_binding_model = tbm, phonon_ids = (pid, pid), bond = bond, α_mean = α)
    add_ssh_coupling!(electron_phonon_model = epm, ssh_coupling = sc, tight_binding_model = tbm)
    dump_config(outdir, tag, mg, tbm, epm; β, Δτ)
end

# ---- config 3: examples/bssh_square.jl (frozen M = Inf mode per cell) -------------------------------------------------------------
function bssh_square(outdir; L = 16, β = 10.0, Δτ = 0.05, Ω = 1.0, α = 0.5, μ = 0.0, tag = "cfg3")
    unit_cell = lu.UnitCell(lattice_vecs = [[1.0, 0.0], [0.0, 1.0]], basis_vecs = [[0.0, 0.0]])
    lattice = lu.Lattice(L = [L, L], periodic = [true, true])
    mg = ModelGeometry(unit_cell, lattice)
    bond_px = lu.Bond(orbitals = (1, 1), displacement = [1, 0]); add_bond!(mg, bond_px)
    bond_py = lu.Bond(orbitals = (1, 1), displacement = [0, 1]); add_bond!(mg, bond_py)
    tbm = TightBindingModel(model_geometry = mg, t_bonds = [bond_px, bond_py], t_mean = [1.0, 1.0], μ = μ, ϵ_mean = [0.0])
    epm = ElectronPhononModel(model_geometry = mg, tight_binding_model = tbm)
    px = add_phonon_mode!(electron_phonon_model = epm, phonon_mode = PhononMode(basis_vec = [0.0, 0.0], Ω_mean = Ω))
    py = add_phonon_mode!(electron_phonon_model = epm, phonon_mode = PhononMode(basis_vec = [0.0, 0.0], Ω_mean = Ω))
    pf = add_phonon_mode!(electron_phonon_model = epm, phonon_mode = PhononMode(basis_vec = [0.0, 0.0], Ω_mean = Ω, M = Inf))
    for (pid, bond) in ((px, bond_px), (py, bond_py))
        sc = SSHCoupling(model_geometry = mg, tight_binding_model = tbm, phonon_ids = (pf, pid), bond = bond, α_mean = α)
        add_ssh_coupling!(electron_phonon_model = epm, ssh_coupling = sc, tight_binding_model = tbm)
    end
    dump_config(outdir, tag, mg, tbm, epm; β, Δτ)
end

function main()
    outdir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden"
    small = length(ARGS) >= 2 && ARGS[2] == "small"
    holstein_honeycomb(outdir)
    small ? ossh_chain(outdir; β = 2.0, tag = "cfg2s") : ossh_chain(outdir)
    small ? bssh_square(outdir; L = 8, β = 1.0, tag = "cfg3s") : bssh_square(outdir)
end
main()
