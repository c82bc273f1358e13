# ParticlesMCB200.jl -- reference-side binding of libpmc_b200 (include/pmc_b200.h).
#
# Drop-in for the `Metropolis` entry of ParticlesMC's algorithm list (src/ParticlesMC.jl:246):
#
#     algorithm_list = ((algorithm = MetropolisB200, pool = pool, seed = seed, parallel = false,
#                        sweepstep = length(chains[1])), outputs...)
#
# Everything else (TOML/CLI schema, Atoms/Molecules, Move/Action/Policy objects, Store*/Print* output
# algorithms) stays the reference's.  One Arianna step advances every chain by `sweepstep` trials in a
# single kernel launch; `system.position`, `system.species`, `system.energy[1]` and the Move counters are
# refreshed from the device whenever Arianna is about to run an output algorithm at the same step (`output_due`).
#
# NOT EXECUTED in the build container (no Julia there; see DESIGN.md).  The same call sequence is exercised
# from Python/ctypes in particlesmc_b200/device.py and tests/.  Arianna's algorithm hook names
# (`Arianna.initialise`, `Arianna.make_step!`, `Arianna.finalise`) must be checked against the installed
# Arianna 0.2.x -- Arianna.jl is not vendored with the reference (Project.toml:6,21).
module ParticlesMCB200

using Arianna, ParticlesMC, StaticArrays

const LIB = get(ENV, "PMC_B200_LIB", joinpath(@__DIR__, "..", "particlesmc_b200", "lib", "libpmc_b200.so"))
const PMC_NPAR = 12

# ---- C structs (include/pmc_b200.h) -------------------------------------------------------------------
struct PmcConfig
    device::Int32; mode::Int32; precision::Int32; n_chains::Int32; n_particles::Int32; dim::Int32
    n_species::Int32; model_kind::Int32; molecules::Int32; chain_offset::Int32; threads::Int32
    prefilter::Int32; r0::Int32; r1::Int32; r2::Int32; r3::Int32
end
struct PmcMove
    kind::Int32; species_a::Int32; species_b::Int32; reserved::Int32; probability::Float64; sigma::Float64
end

check(rc) = rc == 0 || error(unsafe_string(ccall((:pmc_last_error, LIB), Cstring, ())))

# ---- model_matrix -> flat parameter block (slots documented in pmc_b200.h) -----------------------------
model_kind(::ParticlesMC.LennardJones) = 1
model_kind(::ParticlesMC.SoftSpheres) = 2
model_kind(::ParticlesMC.SmoothLennardJones) = 3
model_kind(::ParticlesMC.GeneralKG) = 4
flat(m::ParticlesMC.LennardJones) = [m.rcut, m.rcut2, m.ϵ4, m.σ2, m.shift]
flat(m::ParticlesMC.SoftSpheres) = [m.rcut, m.rcut2, m.ϵ, m.σ2, m.shift, Float64(m.ndiv2)]
flat(m::ParticlesMC.SmoothLennardJones) = [m.rcut, m.rcut2, m.ϵ4, m.σ2, 0.0, m.C0, m.C2_σ2, m.C4_σ4]
flat(m::ParticlesMC.GeneralKG) = [m.rcut, m.rcut2, m.ϵ4, m.σ2, m.shift, m.ϵ4bond, m.σ2bond, m.rcut2bond,
                                  m.shiftbond, m.kr02, m.r02]
function flatten(model_matrix)
    ns = size(model_matrix, 1)
    out = zeros(Float64, PMC_NPAR, ns, ns)          # column-major: [slot, j, i] == C [i][j][slot]
    for i in 1:ns, j in 1:ns
        p = flat(model_matrix[i, j])
        out[1:length(p), j, i] .= p
    end
    return out
end

# ---- the algorithm -------------------------------------------------------------------------------------
mutable struct MetropolisB200{P} <: Arianna.AriannaAlgorithm
    ctx::Ptr{Cvoid}
    pool::P
    sweepstep::Int
    dirty::Bool            # device state is ahead of the host copies
end

function MetropolisB200(chains; pool, seed = 1, parallel = false, sweepstep = length(chains[1]), device = 0,
                        kwargs...)
    s = chains[1]
    N, d = length(s), s.d
    ns = size(s.model_matrix, 1)
    molecules = s isa ParticlesMC.Molecules
    mode = 8d * N + 5N + 16384 <= 200 * 1024 ? 0 : 1     # PMC_MODE_CHAINS if the chain fits shared memory
    cfg = Ref(PmcConfig(device, mode, 0, length(chains), N, d, ns, model_kind(s.model_matrix[1, 1]),
                        molecules, 0, 0, 0, 0, 0, 0, 0))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pmc_create, LIB), Cint, (Ref{PmcConfig}, Ref{Ptr{Cvoid}}), cfg, ctx))
    params = flatten(s.model_matrix)
    check(ccall((:pmc_set_model, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], params))
    if molecules                                           # Molecules.bonds -> CSR, 0-based (molecules.jl:40)
        off = Int32[0; cumsum(length.(s.bonds))]
        idx = Int32[j - 1 for b in s.bonds for j in b]
        check(ccall((:pmc_set_bonds, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), ctx[], off, idx))
        st, ln = Int32.(s.start_mol .- 1), Int32.(s.length_mol)    # molecules.jl:28-29, needed by MoleculeFlip
        check(ccall((:pmc_set_molecules, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}), ctx[], length(st), st, ln))
    end
    for (k, c) in enumerate(chains)                        # Vector{SVector{d,Float64}} is contiguous AoS
        box = collect(Float64, c.box)
        GC.@preserve c box check(ccall((:pmc_upload, LIB), Cint,
            (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{Float64}),
            ctx[], k - 1, 1, pointer(reinterpret(Float64, c.position)), pointer(c.species), box, c.temperature))
    end
    check(ccall((:pmc_init_energy, LIB), Cint, (Ptr{Cvoid},), ctx[]))  # "Initial configuration has infinite or NaN energy."
    moves = map(pool) do mv
        a = mv.action
        if a isa ParticlesMC.Displacement
            mv.policy isa ParticlesMC.SimpleGaussian || error("Displacement needs the SimpleGaussian policy")
            PmcMove(0, 0, 0, 0, mv.probability, mv.parameters.σ)
        elseif a isa ParticlesMC.DiscreteSwap
            mv.policy isa ParticlesMC.DoubleUniform || error("DiscreteSwap needs the DoubleUniform policy")
            PmcMove(1, a.species[1], a.species[2], 0, mv.probability, 0.0)
        elseif a isa ParticlesMC.MoleculeFlip
            mv.policy isa ParticlesMC.DoubleUniform || error("MoleculeFlip needs the DoubleUniform policy")
            PmcMove(2, 0, 0, 0, mv.probability, 0.0)
        else
            error("$(typeof(a)) is not on the device path")
        end
    end |> collect
    check(ccall((:pmc_set_moves, LIB), Cint, (Ptr{Cvoid}, Ptr{PmcMove}, Int32), ctx[], moves, length(moves)))
    check(ccall((:pmc_seed, LIB), Cint, (Ptr{Cvoid}, UInt64), ctx[], UInt64(seed)))
    alg = MetropolisB200(ctx[], pool, sweepstep, false)
    finalizer(a -> ccall((:pmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), a.ctx), alg)
    return alg
end

# Does any OTHER algorithm of the simulation (StoreCallbacks, StoreTrajectories, StoreLastFrames, StoreAcceptance,
# PrintTimeSteps: src/ParticlesMC.jl:249-291) run at the current step?  Arianna's `run!` calls, at every step t, the
# `make_step!` of each algorithm whose scheduler contains t, in list order -- Metropolis first (it is pushed first,
# src/ParticlesMC.jl:246), the outputs after it.  The field names `t`, `algorithms`, `schedulers` are those of Arianna
# 0.2.x's `Simulation`; if this installation names them differently the answer is a conservative `true` (the host
# copies are then refreshed after every step: correct, only slower).
function output_due(simulation, alg)
    all(p -> hasproperty(simulation, p), (:t, :algorithms, :schedulers)) || return true
    t = simulation.t
    for (a, sched) in zip(simulation.algorithms, simulation.schedulers)
        a === alg && continue
        t in sched && return true
    end
    return false
end

# One Arianna step: `sweepstep` trials per chain in one launch.  The launch is asynchronous; the host copies
# (`position`, `species`, `energy[1]`, the Move counters) are refreshed only when an output algorithm is about to read
# them at this very step, so a run that stores every 1000 sweeps pays one download per 1000 launches.
function Arianna.make_step!(simulation::Arianna.Simulation, alg::MetropolisB200)
    check(ccall((:pmc_run, LIB), Cint, (Ptr{Cvoid}, Int64), alg.ctx, alg.sweepstep))
    alg.dirty = true
    output_due(simulation, alg) && sync_host!(simulation, alg)
    return nothing
end

# Called before any output algorithm reads the chains (StoreCallbacks / StoreTrajectories / StoreLastFrames /
# StoreAcceptance, src/ParticlesMC.jl:249-291): bring host copies up to date.
function sync_host!(simulation::Arianna.Simulation, alg::MetropolisB200)
    alg.dirty || return
    chains = simulation.chains
    M = length(chains)
    e = Vector{Float64}(undef, M)
    check(ccall((:pmc_energy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), alg.ctx, e))
    for (k, c) in enumerate(chains)
        c.energy[1] = e[k]
        GC.@preserve c check(ccall((:pmc_download, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Int64}),
                                   alg.ctx, k - 1, 1, pointer(reinterpret(Float64, c.position)), pointer(c.species)))
    end
    nm = length(alg.pool)
    calls = Matrix{Int64}(undef, nm, M); acc = similar(calls)
    check(ccall((:pmc_counters, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), alg.ctx, calls, acc))
    for (m, mv) in enumerate(alg.pool)
        mv.total_calls = sum(@view calls[m, :])
        mv.accepted_calls = sum(@view acc[m, :])
    end
    alg.dirty = false
end

Arianna.finalise(alg::MetropolisB200, simulation::Arianna.Simulation) = sync_host!(simulation, alg)

# ---- observables computed on the device (no download of the configurations) -----------------------------
# chain_correlation callback (src/molecules.jl:224-246) of every chain
function device_chain_correlation(alg::MetropolisB200, nchains::Integer)
    out = Vector{Float64}(undef, nchains)
    check(ccall((:pmc_chain_correlation, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), alg.ctx, out))
    return out
end

# counts of energy[1]/N of all chains in nbins equal bins on [emin, emax)
function device_energy_histogram(alg::MetropolisB200, emin, emax, nbins::Integer; per_particle::Bool = true)
    h = zeros(UInt64, nbins)
    check(ccall((:pmc_energy_histogram, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Int32, Int32, Ptr{UInt64}),
                alg.ctx, emin, emax, nbins, per_particle, h))
    return h
end

# raw pair-distance counts behind g(r); species labels 1-based, 0 = any
function device_pair_histogram(alg::MetropolisB200, species_a::Integer, species_b::Integer, rmax, nbins::Integer)
    h = zeros(UInt64, nbins)
    check(ccall((:pmc_pair_histogram, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Float64, Int32, Ptr{UInt64}),
                alg.ctx, species_a, species_b, rmax, nbins, h))
    return h
end

export MetropolisB200, sync_host!, device_chain_correlation, device_energy_histogram, device_pair_histogram

end # module
