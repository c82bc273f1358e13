// chains_fast.cu -- host-side launchers of the hand-scheduled chain kernels (chains_fast.cuh).  Separate
// translation unit so that the kernel families compile in parallel.
#include <cuda_runtime.h>

#include <type_traits>

#include "chains.cuh"
#include "chains_fast.cuh"

namespace pmc {

namespace {

template <typename F>
cudaError_t dispatch(int dim, int model, bool, F &&f) {
#define PMC_CASE(D, MDL) \
    if (dim == D && model == MDL) return f(std::integral_constant<int, D>{}, std::integral_constant<int, MDL>{}, std::false_type{});
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

// up to 8 register candidates per thread in 3-D (24 registers), 16 in 2-D (32 registers)
bool chain_fast_supported(int dim, int Npad, int threads) {
    return Npad <= (dim == 2 ? 2 : 1) * fast::kFastMaxCand * fast::kFastThreads && threads == fast::kFastThreads;
}

static int fast_npad(int Npad) { return Npad <= 256 ? 256 : (Npad <= 512 ? 512 : (Npad <= 1024 ? 1024 : 2048)); }

size_t chain_fast_smem_bytes(int dim, int Npad, int model, bool swaps) {
    const bool full_par = !(model == PMC_MODEL_LJ || model == PMC_MODEL_KG);
    return fast::fast_layout(dim, fast_npad(Npad), PMC_MAX_SPECIES, swaps, full_par).total;
}

template <typename F>
static cudaError_t fast_dispatch(int dim, int model, int Npad, bool swaps, F &&f) {
    return dispatch(dim, model, false, [&](auto D, auto MDL, auto) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        const int np = fast_npad(Npad);
        if constexpr (d == 2) {
            if (np == 2048) return swaps ? f(fast::k_chain_sweep_fast<d, mdl, 2048, true>) : f(fast::k_chain_sweep_fast<d, mdl, 2048, false>);
        }
        if (swaps) {
            if (np == 256) return f(fast::k_chain_sweep_fast<d, mdl, 256, true>);
            if (np == 512) return f(fast::k_chain_sweep_fast<d, mdl, 512, true>);
            return f(fast::k_chain_sweep_fast<d, mdl, 1024, true>);
        }
        if (np == 256) return f(fast::k_chain_sweep_fast<d, mdl, 256, false>);
        if (np == 512) return f(fast::k_chain_sweep_fast<d, mdl, 512, false>);
        return f(fast::k_chain_sweep_fast<d, mdl, 1024, false>);
    });
}

size_t chain_mixed_smem_bytes(int dim, int Npad) { return fast::mixed_layout(dim, fast_npad(Npad)).total; }

template <typename F>
static cudaError_t mixed_dispatch(int dim, int model, int Npad, F &&f) {
    return dispatch(dim, model, false, [&](auto D, auto MDL, auto) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        switch (fast_npad(Npad)) {
        case 256: return f(fast::k_chain_sweep_mixed<d, mdl, 256>);
        case 512: return f(fast::k_chain_sweep_mixed<d, mdl, 512>);
        default: return f(fast::k_chain_sweep_mixed<d, mdl, 1024>);
        }
    });
}

cudaError_t launch_chain_sweep_mixed(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st) {
    return mixed_dispatch(dim, model, a.Npad, [&](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<M, fast::kFastThreads, smem, st>>>(a);
        return cudaGetLastError();
    });
}

cudaError_t configure_chain_fast(int dim, int model, int Npad, bool swaps, size_t smem) {
    return fast_dispatch(dim, model, Npad, swaps, [&](auto kernel) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
}

cudaError_t launch_chain_sweep_fast(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st) {
    return fast_dispatch(dim, model, a.Npad, a.any_swap != 0, [&](auto kernel) {
        kernel<<<M, fast::kFastThreads, smem, st>>>(a);
        return cudaGetLastError();
    });
}

}  // namespace pmc
