# julia_baseline.jl -- the reference's own CPU throughput on BASELINE config 1 (KA N=1000, LinkedList,
# Displacement sigma=0.05), chains over Julia threads exactly as `parallel = true` does it.
# Not runnable in this repo's container (no Julia); bench.py reports the C restatement (oracle/) instead and
# labels it so.  Run where Julia + Arianna exist to fill in the "Julia" cell of the results table:
#
#   julia -t auto --project=<ParticlesMC checkout> bench/julia_baseline.jl
using Arianna, ParticlesMC, Random, StaticArrays, ComponentArrays, Base.Threads

N, ρ, T, sweeps = 1000, 1.2, 1.0, 50
m = 10
L = (N / ρ)^(1 / 3)
position = [SVector(((i + 0.5), (j + 0.5), (k + 0.5)) .* (L / m)) for i in 0:m-1 for j in 0:m-1 for k in 0:m-1]
species = shuffle!(Xoshiro(0), vcat(ones(Int, 800), 2ones(Int, 200)))
chains = [System(copy(position), copy(species), ρ, T, KobAndersen(); list_type = LinkedList) for _ in 1:nthreads()]
pool = (Move(Displacement(0, zero(chains[1].box), 0.0), SimpleGaussian(), ComponentArray(σ = 0.05), 1.0),)
rngs = [Xoshiro(42 + k) for k in 1:nthreads()]
@threads for k in eachindex(chains)          # warm-up / compilation
    Arianna.mc_sweep!(chains[k], pool, rngs[k]; mc_steps = N)
end
t = @elapsed @threads for k in eachindex(chains)
    for _ in 1:sweeps
        Arianna.mc_sweep!(chains[k], pool, rngs[k]; mc_steps = N)   # benchmark/particles_benchmarks.jl:29
    end
end
println("threads=$(nthreads()) attempted moves/s = $(length(chains) * sweeps * N / t)")
