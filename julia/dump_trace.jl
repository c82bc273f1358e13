# dump_trace.jl -- record a proposal/decision trace from the REAL reference (ParticlesMC + Arianna) so that
# bit-level accept/reject parity can be pinned on any machine that has Julia (this repo's build container does
# not; until such a trace exists the accept/reject rule is "parity unpinned", see oracle/pmc_oracle.c).
#
#   julia --project=<ParticlesMC checkout> julia/dump_trace.jl config_0.xyz JBB 0.231 2000 trace.bin
#
# Output: little-endian records of include/pmc_b200.h `pmc_trial` (48 bytes) followed by one byte `accepted`
# and one Float64 `system.energy[1]` after the trial -- the exact input of pmc_replay / tests' replay check.
using Arianna, ParticlesMC, Random, ComponentArrays, StaticArrays

# An rng that logs every uniform it hands out, wrapping the chain rng Arianna would use.
struct LoggingRNG{R<:AbstractRNG} <: AbstractRNG
    inner::R
    log::Vector{Float64}
end
Random.rand(r::LoggingRNG, ::Random.SamplerTrivial{Random.CloseOpen01{Float64}}) = (u = rand(r.inner); push!(r.log, u); u)
Random.rand(r::LoggingRNG, sp::Random.Sampler) = rand(r.inner, sp)

function main(config, model, temperature, ntrials, out)
    chains = load_chains(config, args = Dict("temperature" => [temperature], "model" => [model], "list_type" => "LinkedList"))
    system = chains[1]
    rng = LoggingRNG(Xoshiro(10), Float64[])
    action = Displacement(0, zero(system.box), 0.0)
    policy, parameters = SimpleGaussian(), ComponentArray(σ = 0.05)
    open(out, "w") do io
        for t in 1:ntrials
            empty!(rng.log)
            accepted = Arianna.mc_step!(system, action, policy, parameters, rng)   # benchmark/particles_benchmarks.jl:28
            δ = -action.δ                      # mc_step! leaves the action inverted (src/moves.jl:88-90)
            write(io, Int32(0), Int32(0), Int32(action.i - 1), Int32(-1))
            for a in 1:3
                write(io, a <= system.d ? Float64(δ[a]) : 0.0)
            end
            write(io, Float64(rng.log[end]))   # the last uniform drawn is the acceptance uniform
            write(io, UInt8(accepted), Float64(system.energy[1]))
        end
    end
end

main(ARGS[1], ARGS[2], parse(Float64, ARGS[3]), parse(Int, ARGS[4]), ARGS[5])
