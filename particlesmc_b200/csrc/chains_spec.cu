// chains_spec.cu -- host-side launcher of the speculative chain kernel (chains_spec.cuh).  Separate translation
// unit so that the kernel families compile in parallel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "chains.cuh"
#include "chains_spec.cuh"

namespace pmc {

namespace {

int spec_npad(int Npad) {
    return Npad <= 256 ? 256 : (Npad <= 512 ? 512 : (Npad <= 1024 ? 1024 : (Npad <= 2048 ? 2048 : (Npad <= 3072 ? 3072 : 4096))));
}

// Shapes built: Atoms N <= 1024 (4 warps; + PMC_MIXED), Atoms N <= 2048 (2-D with 4 warps, 3-D with 8), Molecules
// (GeneralKG, 3-D) N <= 1024 with 4 warps and N <= 4096 with 8.
template <bool MIXED, bool SWAPS, typename F>
cudaError_t spec_dispatch(int dim, int model, int Npad, bool mol, int threads, F &&f) {
    const int np = spec_npad(Npad);
    constexpr int kNwS = (MIXED || SWAPS) ? 4 : PMC_SPEC_NW_SMALL;  // trials in flight for the small shapes
    if (SWAPS && !mol && (MIXED || model == PMC_MODEL_KG)) return cudaErrorInvalidValue;
    if (mol) {  // Molecules: GeneralKG, 3-D; Displacement pools, and pools with MoleculeFlip / DiscreteSwap (SWAPS)
        if (MIXED || dim != 3 || model != PMC_MODEL_KG) return cudaErrorInvalidValue;
        if constexpr (!MIXED) {
            if (threads == 128) {
                if (np == 256) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 256, false, true, 4, SWAPS>);
                if (np == 512) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 512, false, true, 4, SWAPS>);
                if (np == 1024) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 1024, false, true, 4, SWAPS>);
            } else {
                if (np == 2048) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 2048, false, true, 8, SWAPS>);
                if (np == 3072) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 3072, false, true, 8, SWAPS>);
                if (np == 4096) return f(spec::k_chain_sweep_spec<3, PMC_MODEL_KG, 4096, false, true, 8, SWAPS>);
            }
        }
        return cudaErrorInvalidValue;
    }
#define PMC_CASE(D, MDL)                                                                                                \
    if (dim == D && model == MDL) {                                                                                     \
        if constexpr (!SWAPS || MDL != PMC_MODEL_KG) {                                                                  \
            if (np == 256) return f(spec::k_chain_sweep_spec<D, MDL, 256, MIXED, false, kNwS, SWAPS>);                  \
            if (np == 512) return f(spec::k_chain_sweep_spec<D, MDL, 512, MIXED, false, kNwS, SWAPS>);                  \
            if (np == 1024) return f(spec::k_chain_sweep_spec<D, MDL, 1024, MIXED, false, kNwS, SWAPS>);                \
            if constexpr (!MIXED) {                                                                                     \
                if (np == 2048) return f(spec::k_chain_sweep_spec<D, MDL, 2048, false, false, D == 2 ? 4 : 8, SWAPS>);  \
            }                                                                                                           \
        }                                                                                                               \
        return cudaErrorInvalidValue;                                                                                   \
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

int spec_warps(int dim, int Npad, bool mol, bool mixed, bool swaps) {
    if (spec_npad(Npad) <= 1024 && !mol && !mixed && !swaps) return PMC_SPEC_NW_SMALL;
    return spec_npad(Npad) <= 1024 || (dim == 2 && !mol) ? 4 : 8;
}

}  // namespace

bool chain_spec_supported(int dim, int model, int Npad, int threads, bool mol, bool mixed, bool swaps) {
    const int np = spec_npad(Npad);
    if (Npad > np || threads != 32 * spec_warps(dim, Npad, mol, mixed, swaps)) return false;
    if (swaps && (mixed || (!mol && model == PMC_MODEL_KG))) return false;
    if (mixed) return !mol && np <= 1024;
    if (mol) return dim == 3 && model == PMC_MODEL_KG;
    return np <= 2048;
}

size_t chain_spec_smem_bytes(int dim, int Npad, int model, bool mixed, bool mol, bool swaps) {
    const bool full_par = !mixed && (mol || !(model == PMC_MODEL_LJ || model == PMC_MODEL_KG));
    return spec::spec_layout(dim, spec_npad(Npad), full_par, mixed, spec_warps(dim, Npad, mol, mixed, swaps), swaps).total;
}

template <typename F>
static cudaError_t spec_dispatch_rt(int dim, int model, int Npad, bool mixed, bool mol, bool swaps, F &&f) {
    const int th = 32 * spec_warps(dim, Npad, mol, mixed, swaps);
    if (mixed) return spec_dispatch<true, false>(dim, model, Npad, mol, th, f);
    if (swaps) return spec_dispatch<false, true>(dim, model, Npad, mol, th, f);
    return spec_dispatch<false, false>(dim, model, Npad, mol, th, f);
}

cudaError_t configure_chain_spec(int dim, int model, int Npad, size_t smem, bool mixed, bool mol, bool swaps) {
    return spec_dispatch_rt(dim, model, Npad, mixed, mol, swaps, [&](auto kernel) {
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
}

// One CTA per chain, or -- when the chains do not fill a whole number of waves of resident CTAs -- persistent CTAs
// pulling (chain, segment) units from `queue` (chains_spec.cuh).  PMC_SPEC_QUEUE=0 in the environment forces the plain
// launch (A/B measurements).
cudaError_t launch_chain_sweep_spec(int dim, int model, int M, size_t smem, const ChainArgs &a_in, cudaStream_t st, bool mixed,
                                    bool mol, bool swaps, int32_t *queue) {
    const int th = 32 * spec_warps(dim, a_in.Npad, mol, mixed, swaps);
    static const bool queue_on = [] {
        const char *e = std::getenv("PMC_SPEC_QUEUE");
        return !(e && e[0] == '0');
    }();
    return spec_dispatch_rt(dim, model, a_in.Npad, mixed, mol, swaps, [&](auto kernel) {
        ChainArgs a = a_in;
        int grid = M;
        a.queue = nullptr;
        a.n_chains = M;
        a.n_seg = 1;
        a.seg_len = a.n_trials;
        if (queue && queue_on && !a.replay && !a.trace && !a.acc_out && !a.dE_out) {
            int dev = 0, sms = 0, occ = 0;
            cudaError_t e = cudaGetDevice(&dev);
            if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, th, smem);
            if (e != cudaSuccess) return e;
            const long long slots = (long long)sms * occ;
            if (slots > 0 && M > slots && M % slots != 0) {
                // >= 24 waves of units, segments of at least four proposal batches (state load/store stays < 1 %)
                long long nseg = std::min((24 * slots + M - 1) / M, a.n_trials / (4 * spec::kSpecBatch));
                if (nseg >= 2) {
                    long long len = (a.n_trials + nseg - 1) / nseg;
                    len = (len + spec::kSpecBatch - 1) / spec::kSpecBatch * spec::kSpecBatch;
                    nseg = (a.n_trials + len - 1) / len;
                    a.queue = queue;
                    a.n_seg = (int)nseg;
                    a.seg_len = len;
                    grid = (int)std::min<long long>(slots, (long long)M * nseg);
                    e = cudaMemsetAsync(queue, 0, sizeof(int32_t) * ((size_t)M + 1), st);
                    if (e != cudaSuccess) return e;
                }
            }
        }
        kernel<<<grid, th, smem, st>>>(a);
        return cudaGetLastError();
    });
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
