// chains.cuh -- argument block shared by the host API and the chain-mode kernels.
#pragma once
#include <stdint.h>

#include "pmc_b200.h"

namespace pmc {

// Device state of PMC_MODE_CHAINS: M independent systems of N particles each.
//   x    [M][dim][Npad] float64, wrapped into [0, L]      (SoA: coalesced loads, conflict-free smem)
//   img  [M][dim][Npad] int32 image counters              (x_unwrapped = x + img * L)
//   sp   [M][Npad]      uint8 species, 0-based
//   spids/heads [M][Npad] uint16: SpeciesList (src/utils.jl:31-49), ids grouped by species
struct ChainArgs {
    double *x;
    int32_t *img;
    uint8_t *sp;
    uint16_t *spids;
    uint16_t *heads;
    const int32_t *spoff;  // [M][PMC_MAX_SPECIES+1]
    const double *box;     // [M][3]
    const double *temp;    // [M]
    double *energy;        // [M] running energy[1]
    unsigned long long *calls;     // [M][PMC_MAX_MOVES]
    unsigned long long *accepted;  // [M][PMC_MAX_MOVES]
    const double *par;             // [ns][ns][PMC_NPAR]
    const uint16_t *bonds;         // [Npad][PMC_MAX_BONDS], 0xFFFF = none (shared topology)
    const int32_t *mol_start;      // [n_mol] first site of each molecule (MoleculeFlip)
    const int32_t *mol_len;        // [n_mol]
    int32_t n_mol;
    // move pool
    int32_t n_moves;
    int32_t mv_kind[PMC_MAX_MOVES];
    int32_t mv_a[PMC_MAX_MOVES];  // species, 0-based
    int32_t mv_b[PMC_MAX_MOVES];
    double mv_cum[PMC_MAX_MOVES];  // cumulative selection probability (normalised)
    float mv_sigma[PMC_MAX_MOVES];
    int32_t any_swap;
    // rng
    unsigned long long seed;
    unsigned long long t0;  // index of the first trial of this launch
    int32_t chain_offset;
    // shape
    int32_t N, Npad, ns;
    long long n_trials;
    // replay / trace (test hooks)
    const pmc_trial *replay;
    pmc_trial *trace;
    uint8_t *acc_out;
    double *dE_out;
    int32_t exact_exp;  // 1: accept iff min(1, exp(-(e2-e1)/T)) > u (reference form); 0: -dE/T > log u
    // work queue of the speculative kernel (chains_spec.cuh): with more chains than resident CTAs the launch is cut into
    // n_seg segments of seg_len trials per chain and persistent CTAs pull (chain, segment) units from queue[0];
    // queue[1 + chain] counts the finished segments of a chain (a unit waits for its predecessor).  nullptr: one CTA per chain.
    int32_t *queue;
    unsigned long long *stats;  // pmc_work_counters: [0] fp64-evaluated candidates, [1] trial evaluations (nullptr: off)
    int32_t n_chains, n_seg;
    long long seg_len;
};

struct EnergyArgs {
    const double *x;
    const uint8_t *sp;
    const double *box;
    const double *par;
    const uint16_t *bonds;
    double *eloc;  // [M][Npad]
    double *etot;  // [M]
    int32_t N, Npad, ns;
};

// pair-distance histogram of every chain (observable behind g(r)); hist is [nbins] u64, accumulated
cudaError_t launch_chain_pair_histogram(int dim, int M, int N, int Npad, const double *x, const uint8_t *sp,
                                        const double *box, int sa, int sb, double rmax, int nbins,
                                        unsigned long long *hist, cudaStream_t st);
size_t chain_sweep_smem_bytes(int dim, int Npad, int ns, int threads, bool mol, bool any_swap, bool filter);
size_t chain_energy_smem_bytes(int dim, int Npad, int ns, bool mol);
cudaError_t launch_chain_sweep(int dim, int model, bool mol, bool filter, int M, int threads, size_t smem,
                               const ChainArgs &a, cudaStream_t st);
cudaError_t launch_chain_energy(int dim, int model, bool mol, int M, size_t smem, const EnergyArgs &a,
                                cudaStream_t st);
// hand-scheduled kernel for Atoms + Displacement-only pools + cubic boxes + N <= 1024 (chains_fast.cuh)
bool chain_fast_supported(int dim, int Npad, int threads);
size_t chain_fast_smem_bytes(int dim, int Npad, int model, bool swaps);
cudaError_t configure_chain_fast(int dim, int model, int Npad, bool swaps, size_t smem);
cudaError_t launch_chain_sweep_fast(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st);
// speculative kernel, one warp per trial (chains_spec.cuh), cubic boxes: Atoms N <= 2048 with Displacement and
// DiscreteSwap pools, Molecules (GeneralKG, 3-D, Displacement) N <= 4096, PMC_MIXED (Displacement) N <= 1024
bool chain_spec_supported(int dim, int model, int Npad, int threads, bool mol, bool mixed, bool swaps);
size_t chain_spec_smem_bytes(int dim, int Npad, int model, bool mixed, bool mol, bool swaps);
cudaError_t configure_chain_spec(int dim, int model, int Npad, size_t smem, bool mixed, bool mol, bool swaps);
cudaError_t launch_chain_sweep_spec(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st, bool mixed,
                                    bool mol, bool swaps, int32_t *queue);
// local energies through the 8-bit prefilter (Atoms, cubic box, N <= 1024)
cudaError_t launch_chain_energy_fast(int dim, int model, int M, const EnergyArgs &a, cudaStream_t st);
// PMC_MIXED variant of the fast kernel (fp32 pair terms on fixed-point coordinates, fp64 accumulation)
size_t chain_mixed_smem_bytes(int dim, int Npad);
cudaError_t launch_chain_sweep_mixed(int dim, int model, int M, size_t smem, const ChainArgs &a, cudaStream_t st);
cudaError_t configure_chain_kernels(int dim, int model, bool mol, size_t sweep_smem, size_t sweep_smem_filter,
                                    size_t energy_smem);

}  // namespace pmc
