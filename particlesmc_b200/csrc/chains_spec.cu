// chains_spec.cu -- host-side launcher of the speculative chain kernel (chains_spec.cuh).  Separate translation
// unit so that the kernel families compile in parallel.
#include <cuda_runtime.h>

#include "chains.cuh"
#include "chains_spec.cuh"

namespace pmc {

namespace {

int spec_npad(int Npad) { return Npad <= 256 ? 256 : (Npad <= 512 ? 512 : 1024); }

template <bool MIXED, typename F>
cudaError_t spec_dispatch(int dim, int model, int Npad, F &&f) {
    const int np = spec_npad(Npad);
#define PMC_CASE(D, MDL)                                                              \
    if (dim == D && model == MDL) {                                                   \
        if (np == 256) return f(spec::k_chain_sweep_spec<D, MDL, 256, MIXED>);        \
        if (np == 512) return f(spec::k_chain_sweep_spec<D, MDL, 512, MIXED>);        \
        return f(spec::k_chain_sweep_spec<D, MDL, 1024, MIXED>);                      \
    }
    PMC_CASE(3, PMC_MODEL_LJ)
    PMC_CASE(2, PMC_MODEL_LJ)
    PMC_CASE(3, PMC_MODEL_SOFT)
    PMC_CASE(2, PMC_MODEL_SOFT)
    PMC_CASE(3, PMC_MODEL_SMOOTHLJ)
    PMC_CASE(2, PMC_MODEL_SMOOTHLJ)
    PMC_CASE(3, PMC_MODEL_KG)
    PMC_CASE(2, PMC_MODEL_KG)
#undef PMC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace

// 32 packed candidates per lane: N <= 1024, 128 threads (= four speculative trials per round)
bool chain_spec_supported(int Npad, int threads) { return Npad <= 1024 && threads == spec::kSpecThreads; }

size_t chain_spec_smem_bytes(int dim, int Npad, int model, bool mixed) {
    const bool full_par = !mixed && !(model == PMC_MODEL_LJ || model == PMC_MODEL_KG);
    return spec::spec_layout(dim, spec_npad(Npad), full_par, mixed).total;
}

cudaError_t configure_chain_spec(int dim, int model, int Npad, size_t smem, bool mixed) {
    auto set = [&](auto kernel) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); };
    return mixed ? spec_dispatch<true>(dim, model, Npad, set) : spec_dispatch<false>(dim, model, Npad, set);
}

cudaError_t launch_chain_sweep_spec(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st, bool mixed) {
    auto go = [&](auto kernel) {
        kernel<<<M, spec::kSpecThreads, smem, st>>>(a);
        return cudaGetLastError();
    };
    return mixed ? spec_dispatch<true>(dim, model, a.Npad, go) : spec_dispatch<false>(dim, model, a.Npad, go);
}

// ---- local energies through the prefilter (k_chain_energy_fast) --------------------------------------------------
template <typename F>
static cudaError_t energy_dispatch(int dim, int model, int Npad, F &&f) {
    const int np = spec_npad(Npad);
#define PMC_CASE(D, MDL)                                                    \
    if (dim == D && model == MDL) {                                         \
        if (np == 256) return f(spec::k_chain_energy_fast<D, MDL, 256>);    \
        if (np == 512) return f(spec::k_chain_energy_fast<D, MDL, 512>);    \
        return f(spec::k_chain_energy_fast<D, MDL, 1024>);                  \
    }
    PMC_CASE(3, PMC_MODEL_LJ)
    PMC_CASE(2, PMC_MODEL_LJ)
    PMC_CASE(3, PMC_MODEL_SOFT)
    PMC_CASE(2, PMC_MODEL_SOFT)
    PMC_CASE(3, PMC_MODEL_SMOOTHLJ)
    PMC_CASE(2, PMC_MODEL_SMOOTHLJ)
    PMC_CASE(3, PMC_MODEL_KG)
    PMC_CASE(2, PMC_MODEL_KG)
#undef PMC_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_chain_energy_fast(int dim, int model, int M, const EnergyArgs &a, cudaStream_t st) {
    const bool full_par = !(model == PMC_MODEL_LJ || model == PMC_MODEL_KG);
    const size_t smem = spec::energy_layout(dim, spec_npad(a.Npad), full_par).total;
    return energy_dispatch(dim, model, a.Npad, [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<M, spec::kSpecThreads, smem, st>>>(a);
        return cudaGetLastError();
    });
}

}  // namespace pmc
