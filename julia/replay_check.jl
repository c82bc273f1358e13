# replay_check.jl -- ONE command that pins bit-level accept/reject parity against the REAL reference (SURVEY 8c: the
# acceptance rule lives in Arianna.jl, which this repo's build container does not have):
#
#   julia --project=<ParticlesMC checkout> julia/replay_check.jl test/config_0.xyz JBB 0.231 2000
#
# 1. runs the reference's own mc_step! (benchmark/particles_benchmarks.jl:28) for `ntrials` Displacement trials and
#    records every proposal (i, delta), the acceptance uniform, the decision and system.energy[1] (julia/dump_trace.jl);
# 2. uploads the SAME initial configuration to libpmc_b200 and replays the recorded proposals through the production
#    sweep kernel (pmc_replay: reference acceptance arithmetic min(1, exp(-dE/T)) > u);
# 3. compares decision for decision and the running energy; exits non-zero on the first difference.
using Arianna, ParticlesMC
include(joinpath(@__DIR__, "ParticlesMCB200.jl"))
using .ParticlesMCB200: LIB, PmcConfig, check, flatten, model_kind

struct PmcTrial
    kind::Int32; move::Int32; i::Int32; j::Int32; d0::Float64; d1::Float64; d2::Float64; u::Float64
end

function main(config, model, temperature, ntrials)
    trace = tempname()
    run(`$(Base.julia_cmd()) --project=$(Base.active_project()) $(joinpath(@__DIR__, "dump_trace.jl")) $config $model $temperature $ntrials $trace`)
    raw = read(trace)
    rec = 48 + 1 + 8
    trials = [reinterpret(PmcTrial, raw[(k - 1) * rec + 1:(k - 1) * rec + 48])[1] for k in 1:ntrials]
    ref_acc = [raw[(k - 1) * rec + 49] for k in 1:ntrials]
    ref_E = [reinterpret(Float64, raw[(k - 1) * rec + 50:(k - 1) * rec + 57])[1] for k in 1:ntrials]
    # the same initial configuration, as dump_trace.jl loads it
    s = load_chains(config, args = Dict("temperature" => [temperature], "model" => [model], "list_type" => "LinkedList"))[1]
    N, d, ns = length(s), s.d, size(s.model_matrix, 1)
    cfg = Ref(PmcConfig(0, 0, 0, 1, N, d, ns, model_kind(s.model_matrix[1, 1]), 0, 0, 0, 0, 0, 0, 0, 0))
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pmc_create, LIB), Cint, (Ref{PmcConfig}, Ref{Ptr{Cvoid}}), cfg, ctx))
    check(ccall((:pmc_set_model, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], flatten(s.model_matrix)))
    box = collect(Float64, s.box)
    GC.@preserve s box check(ccall((:pmc_upload, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{Float64}),
        ctx[], 0, 1, pointer(reinterpret(Float64, s.position)), pointer(s.species), box, s.temperature))
    check(ccall((:pmc_init_energy, LIB), Cint, (Ptr{Cvoid},), ctx[]))
    acc = Vector{UInt8}(undef, ntrials); dE = Vector{Float64}(undef, ntrials)
    check(ccall((:pmc_replay, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{PmcTrial}, Ptr{UInt8}, Ptr{Float64}), ctx[], ntrials, trials, acc, dE))
    E = Ref(0.0)
    check(ccall((:pmc_energy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx[], E))
    ccall((:pmc_destroy, LIB), Cvoid, (Ptr{Cvoid},), ctx[])
    bad = findfirst(acc .!= ref_acc)
    bad === nothing || error("decision of trial $bad differs: reference $(ref_acc[bad]), device $(acc[bad]) (dE = $(dE[bad]))")
    abs(E[] - ref_E[end]) <= 1e-10 * abs(ref_E[end]) || error("running energy differs: reference $(ref_E[end]), device $(E[])")
    println("replay parity OK: $ntrials decisions identical, $(sum(ref_acc)) accepted, final energy $(E[]) (reference $(ref_E[end]))")
end

main(ARGS[1], ARGS[2], parse(Float64, ARGS[3]), parse(Int, ARGS[4]))
