// chains_spec.cuh -- speculative sweep kernel of PMC_MODE_CHAINS (cubic boxes; Atoms up to N = 2048 with Displacement and
// DiscreteSwap pools, Molecules up to N = 4096 with Displacement / MoleculeFlip / DiscreteSwap pools, PMC_MIXED up to
// N = 1024): ONE WARP PER TRIAL, four (N <= 1024) or eight consecutive trials of a chain in flight per round.
//
// k_chain_sweep_fast (chains_fast.cuh) spends one CTA on one trial: the four warps each scan a quarter of the
// candidates, but everything around the scan -- record and position loads, wrapping, the compaction scan, the block
// reduction, the barrier, the commit -- is replicated in all four warps, and the fp64 pass runs with ~21 of 32 lanes.
// Here warp w of the CTA evaluates trial t + w of the SAME chain against the current state, alone: it scans all
// candidates (32 per lane, packed 8-bit coordinates streamed from a shared-memory table), compacts the survivors into its own queue, runs
// the fp64 pass at ~86 % lane utilisation and reduces with shuffles only.  After a barrier, ONE warp retires the four
// results and publishes how many retired -- all trials of the round at once, a group of lanes per trial, because this
// serial section is what bounds the kernel; a second barrier starts the next round (retiring in every warp redundantly
// saves that barrier but costs 165 more instructions per move: measured 3 % slower; letting the warp that arrives last
// retire the round behind a shared-memory arrival counter instead of the first barrier: 4 % slower):
//
//   trial t+w stands  <=>  no earlier trial of this round was ACCEPTED with its particle inside the filter sphere of
//                          t+w, tested on the old AND the new position with the very 8-bit test the scan uses.
//
// The survivors of a trial are defined by that test, so a trial that stands has exactly the survivor queue -- same
// members, same order, same positions -- it would have had if it had been evaluated after the earlier trials were
// committed: its dE is bit-identical to the sequential one.  The first trial that does not stand ends the round; it and
// the ones after it are simply evaluated again in the next round.  The Markov chain is therefore the SEQUENTIAL chain
// of the reference (src/moves.jl:11-20 applied one move at a time), not a checkerboard approximation of it; what
// speculation costs is the re-evaluated trials (about 3 % per pair of trials at 35 % acceptance: 3.7 of 4 trials
// retire per round).
#pragma once
#include <type_traits>

#include "chains.cuh"
#include "chains_fast.cuh"
#include "common.cuh"
#include "rng.cuh"

namespace pmc {
namespace spec {

using namespace pmc::fast;

constexpr int kSpecThreads = 128;
constexpr int kSpecWarps = kSpecThreads / 32;  // speculation depth
constexpr int kSpecBatch = 56;                 // parked proposals (56 x 80 B keeps the CTA under 1/6 of the SM shared memory)
constexpr int kSpecQCap = 256;                 // survivor queue entries per warp (beyond: unqueued fallback)
constexpr int kPubBytes = 64;                  // published result of one trial

// Build-time switches of the schedule (A/B numbers in DESIGN.md section 7):
//   PMC_SPEC_ROTATE  the retiring warp and the proposal-generating warps rotate from round to round (warp w of every CTA
//                    sits on SM sub-partition w % 4; measured +0.5 %).
// Measured and dropped in round 2: minimum image / cutoff as predicated PTX (ptxas turns them back into selects, -1.4 %);
// eight trials in flight at N = 1000 (6.6e8 vs 9.5e8 moves/s: longer rounds, more re-evaluated trials; equal at 512 chains per
// GPU, +8 % only at 256); proposal generation spread over all four warps instead of two (-1.2 %: twice the instructions);
// keeping a conflicting trial by ADDING the four pair terms with the earlier accepted particle at retirement instead of
// evaluating it again (exact, the pair energy is additive: evaluations per move 1.079 -> 1.001, but the in-order
// decide / correct pass lengthens the serial section of every round: -5.6 %, and the bit-for-bit equality of split and
// unsplit launches is lost to the rounding of the corrections).
#ifndef PMC_SPEC_ROTATE
#define PMC_SPEC_ROTATE 1
#endif
#ifndef PMC_SPEC_NW_SMALL
#define PMC_SPEC_NW_SMALL 4  // trials in flight per round for Atoms, fp64, N <= 1024
#endif

__device__ __forceinline__ void lds_u32x4(uint32_t a, uint32_t &v0, uint32_t &v1, uint32_t &v2, uint32_t &v3) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
}
// shared-memory add without a generic pointer (a generic shared pointer makes the compiler re-derive the shared window,
// S2R SR_CgaCtaId + LEA, in front of every use)
__device__ __forceinline__ void reds_add_u32(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double v0, double v1) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}

struct SpecLayout {
    uint32_t x, sp, pk, q, cp, rec, pub, cnt, cnt32, par, rcs, spids, heads, spoff, total;
};
__host__ __device__ inline SpecLayout spec_layout(int dim, int Npad, bool full_par, bool mixed = false, int nw = kSpecWarps,
                                                  bool swaps = false) {
    SpecLayout f;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        uint32_t p = o;
        o += (bytes + 15u) & ~15u;
        return p;
    };
    f.x = take((mixed ? 4u : 8u) * dim * Npad);  // fp64 positions, or the 32-bit fixed-point state of PMC_MIXED
    f.sp = take(Npad);
    f.pk = take(4u * Npad);  // packed 8-bit coordinates (common.cuh), word pk_pos(j) = particle j
    f.q = take(2u * kSpecQCap * nw);
    f.cp = take(32u * PMC_MAX_SPECIES * PMC_MAX_SPECIES);
    f.rec = take((uint32_t)kRecBytes * kSpecBatch);
    f.pub = take(2u * kPubBytes * nw + 16u);  // two alternating sets + the retired count of the round
    f.cnt = take(8u * 2 * PMC_MAX_MOVES);
    f.cnt32 = take(4u * (2 * PMC_MAX_MOVES + 4));  // per-batch counters (native 32-bit shared atomics), folded into cnt; + work counters
    f.par = take(full_par ? 8u * PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR : 0u);
    f.rcs = take(8u * PMC_MAX_SPECIES);
    f.spids = take(swaps ? 2u * (uint32_t)Npad : 0u);  // SpeciesList (src/utils.jl:31-49), DiscreteSwap pools only
    f.heads = take(0u);  // positions inside the lists are searched for when a swap commits (rare), not stored
    f.spoff = take(swaps ? 32u : 0u);                   // species offsets [5] + ~threshold of the swap filter [1]
    f.total = o;
    return f;
}

// Candidate k of a lane is particle 32 * k + lane (consecutive particles sit in different lanes); its bit in the
// survivor mask is KC - 1 - k.  The packed table is stored so that one LDS.128 per lane still fetches four candidates:
// table word 128 * c + 4 * lane + e holds particle 32 * (4 c + e) + lane.
template <int KC>
__device__ __forceinline__ uint32_t cand_index(int b, int lane) {
    return ((uint32_t)(KC - 1 - b) << 5) + (uint32_t)lane;
}
// word of particle j in the packed table
__device__ __forceinline__ uint32_t pk_pos(uint32_t j) { return (j & ~127u) | ((j & 31u) << 2) | ((j >> 5) & 3u); }

// The packed candidates live in a shared-memory table that every warp streams through with LDS.128 (4 candidates per
// load): keeping them in 32 registers per thread instead (96 registers, 5 CTAs per SM) measured slower than this
// (6 CTAs = 24 warps per SM), and a commit is one store instead of a 32-way register select.  Squeezing to 72 registers
// and 32 KB for 7 CTAs (batch 32, queue 128) spills and measured 1 % slower again.
// MIXED = true is the PMC_MIXED arithmetic of k_chain_sweep_mixed (chains_fast.cuh) in the same speculative schedule:
// the chain state is the 32-bit fixed-point representation, pair terms are fp32 on wrapping integer distances,
// accumulation across lanes and trials is fp64; 25 KB per chain, 8 CTAs per SM.
// NW = warps per CTA = trials in flight per round (4 for N <= 1024; 8 for the larger shapes, whose chains are few per SM).
// MOL = true: Molecules -- bonded partners (A.bonds, <= PMC_MAX_BONDS per site) are excluded from the pair pass and
// contribute bond_potential (FENE + bonded LJ, src/models.jl:202-226) in a separate pass of the first lanes.
// NPAD up to 4096: the survivor mask of a lane is NPAD / 1024 words.
// SWAPS = true adds DiscreteSwap trials (src/moves.jl:137-214) and, with MOL, MoleculeFlip trials (src/moves.jl:291-352:
// the species of two sites of one molecule exchanged): positions fixed, four local energies in one pass over
// the survivors of two spheres; an accepted swap ends the round for later swaps (the species lists changed).
template <int DIM, int MODEL, int NPAD, bool MIXED = false, bool MOL = false, int NW = 4, bool SWAPS = false>
__global__ void __launch_bounds__(32 * NW, NPAD <= 1024 ? (NW == 8 ? 3 : (MIXED ? 8 : (MOL ? 5 : 6))) : (NW == 4 ? 4 : 2)) k_chain_sweep_spec(const __grid_constant__ ChainArgs A) {
    constexpr int NT = 32 * NW;
    static_assert(!(MIXED && MOL) && !(MIXED && NW != 4), "PMC_MIXED is implemented for Atoms, N <= 1024");
    static_assert(!(SWAPS && MIXED), "DiscreteSwap / MoleculeFlip pools: fp64");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KC = NPAD / 32;  // candidates per lane: k = 4 * c + e  <->  particle j = 128 * c + 4 * lane + e
    constexpr int NM = (KC + 31) / 32, KCW = KC < 32 ? KC : 32;  // mask words per lane, candidates per word
    static_assert(KC >= 4 && KC % 4 == 0 && (KC <= 32 || KC % 32 == 0) && NM <= 4, "candidates come four per LDS.128, 32 per mask word");
    constexpr int Npad = NPAD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = A.N, gNpad = A.Npad, ns = A.ns;
    constexpr bool kFullPar = !MIXED && (MOL || !(MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG));
    const SpecLayout F = spec_layout(DIM, Npad, kFullPar, MIXED, NW, SWAPS);
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
#ifndef PMC_SPEC_PIN_BASE
#define PMC_SPEC_PIN_BASE 1
#endif
#if PMC_SPEC_PIN_BASE
    // keep the shared-window base in a register: left alone, the compiler re-derives it (S2R SR_CgaCtaId + LEA, tens of
    // cycles of latency) in front of dependent stores on the serial path of every round (+2 %; PMC_MIXED, at its 64
    // registers, measured 0.8 % slower with it)
    if constexpr (!MIXED) asm volatile("" : "+r"(sb));
#endif
    const uint32_t nb8 = 8u * (uint32_t)Npad;  // byte stride between coordinate planes (fp64)
    constexpr uint32_t nb4 = 4u * (uint32_t)NPAD;  // ... of the fixed-point planes (MIXED)
    const uint32_t tail = sb + F.pub + 2u * kPubBytes * NW;  // [0] retired count of the round, [4] work unit, [8] running energy
    // survivor-mask bits of this lane that are particles (index < N): padding never reaches the fp64 pass
    uint32_t vmask[NM];
#pragma unroll
    for (int mw = 0; mw < NM; mw++) {
        vmask[mw] = 0u;
#pragma unroll
        for (int b = 0; b < KCW; b++)
            if (cand_index<KCW>(b, lane) + 1024u * (uint32_t)mw < (uint32_t)N) vmask[mw] |= 1u << b;
    }

  // One CTA per chain (queue == nullptr), or persistent CTAs that pull (chain, segment) units from a global counter: with
  // M chains on S resident CTA slots a plain launch runs ceil(M / S) waves and leaves the last one partly empty (4096
  // chains on 888 slots: 4.61 -> 5 waves, 8 % of the launch); units of a fraction of the launch even that out.  The trial
  // streams are counter-based, so cutting a launch into segments changes nothing (tests: split launches == one launch).
  for (;;) {
    int c = blockIdx.x, seg = 0;
    long long seg_lo = 0, seg_hi = A.n_trials;
    if (A.queue) {
        __syncthreads();  // the previous unit is done with the shared state
        if (tid == 0) sts_u32(tail + 4, (uint32_t)atomicAdd(A.queue, 1));
        __syncthreads();
        const int unit = (int)lds_u32(tail + 4);
        if (unit >= A.n_chains * A.n_seg) break;
        c = unit % A.n_chains;
        seg = unit / A.n_chains;
        seg_lo = (long long)seg * A.seg_len;
        seg_hi = min(A.n_trials, seg_lo + A.seg_len);
        if (seg > 0 && tid == 0) {
            // the predecessor was handed out earlier, to a CTA that is running and waits for nothing but ITS predecessor
            const volatile int32_t *done = A.queue + 1 + c;
            while (*done < seg) __nanosleep(100);
            __threadfence();
        }
        __syncthreads();
    }

    // ---- load chain state -----------------------------------------------------------------------------
    const double L = A.box[c * 3], hL = 0.5 * L;
    const double fscale = 4294967296.0 / L;
    const float r2scale = (float)(L * L * 0x1p-32);  // MIXED: fixed-point r^2 units -> length^2
    double *gx = A.x + (size_t)c * DIM * gNpad;
    {
        if constexpr (MIXED) {
            uint32_t *su = (uint32_t *)(smem_raw + F.x);
            for (int a = 0; a < DIM; a++)
                for (int k = tid; k < Npad; k += NT) su[a * Npad + k] = k < gNpad ? to_fixed32(__ldcg(gx + a * gNpad + k), fscale) : 0u;
        } else {
            double *sx = (double *)(smem_raw + F.x);
            for (int a = 0; a < DIM; a++)
                for (int k = tid; k < Npad; k += NT) sx[a * Npad + k] = k < gNpad ? __ldcg(gx + a * gNpad + k) : 0.0;
        }
        uint8_t *ssp = smem_raw + F.sp;
        const uint8_t *gsp = A.sp + (size_t)c * gNpad;
        for (int k = tid; k < Npad; k += NT) ssp[k] = k < gNpad ? __ldcg(gsp + k) : 0;
        double *scp = (double *)(smem_raw + F.cp);
        if constexpr (kFullPar) {
            double *spar = (double *)(smem_raw + F.par);
            for (int k = tid; k < ns * ns * PMC_NPAR; k += NT) spar[k] = A.par[k];
        }
        for (int k = tid; k < ns * ns; k += NT) {
            if constexpr (MIXED) {  // float table {rc2, eps4|eps, sig2, shift, c0|ndiv2, c2, c4, -}
                float *fcp = (float *)(smem_raw + F.cp);
                const double *p = A.par + k * PMC_NPAR;
                fcp[8 * k + 0] = (float)p[PMC_P_RCUT2];
                fcp[8 * k + 1] = (float)p[PMC_P_EPS];
                fcp[8 * k + 2] = (float)p[PMC_P_SIG2];
                fcp[8 * k + 3] = (float)p[PMC_P_SHIFT];
                fcp[8 * k + 4] = (float)p[5];
                fcp[8 * k + 5] = (float)p[6];
                fcp[8 * k + 6] = (float)p[7];
                fcp[8 * k + 7] = 0.0f;
            } else {
                scp[4 * k + 0] = A.par[k * PMC_NPAR + PMC_P_RCUT2];
                scp[4 * k + 1] = A.par[k * PMC_NPAR + PMC_P_EPS];
                scp[4 * k + 2] = A.par[k * PMC_NPAR + PMC_P_SIG2];
                scp[4 * k + 3] = A.par[k * PMC_NPAR + PMC_P_SHIFT];
            }
        }
        unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
        if (tid < 2 * PMC_MAX_MOVES) {
            scnt[tid] = 0ull;
            ((uint32_t *)(smem_raw + F.cnt32))[tid] = 0u;
        }
        if (tid < 4) ((uint32_t *)(smem_raw + F.cnt32))[2 * PMC_MAX_MOVES + tid] = 0u;
        if constexpr (SWAPS) {
            uint16_t *si_ = (uint16_t *)(smem_raw + F.spids);
            const uint16_t *gi = A.spids + (size_t)c * gNpad;
            for (int k = tid; k < Npad; k += NT) si_[k] = k < gNpad ? __ldcg(gi + k) : 0;
            int *sso = (int *)(smem_raw + F.spoff);
            if (tid <= PMC_MAX_SPECIES) sso[tid] = __ldcg(A.spoff + c * (PMC_MAX_SPECIES + 1) + tid);
            if (tid == 0) {  // one conservative threshold over all species pairs (swap filter: no displacement)
                double rc2 = 0.0;
                for (int k = 0; k < ns * ns; k++) rc2 = fmax(rc2, A.par[k * PMC_NPAR + PMC_P_RCUT2]);
                ((uint32_t *)sso)[PMC_MAX_SPECIES + 1] = neg_thr8(sqrt(rc2) * fscale * 0x1p-24);
                *(double *)(smem_raw + F.spoff + 24) = sqrt(rc2) * fscale * 0x1p-24;  // the same radius in byte units
            }
        }
        if (tid < PMC_MAX_SPECIES) {  // largest cutoff radius per species of the moved particle (filter sphere)
            double rc2 = 0.0;
            for (int b = 0; b < ns; b++) rc2 = fmax(rc2, A.par[((tid < ns ? tid : 0) * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            ((double *)(smem_raw + F.rcs))[tid] = sqrt(rc2);
        }
    }
    for (int j = tid; j < Npad; j += NT) {
        uint32_t u[3] = {0u, 0u, 0u};
#pragma unroll
        for (int a = 0; a < DIM; a++) u[a] = j < gNpad ? to_fixed32(__ldcg(gx + a * gNpad + j), fscale) : 0u;
        ((uint32_t *)(smem_raw + F.pk))[pk_pos((uint32_t)j)] = pack8(u[0], u[1], u[2]);
    }
    const uint32_t pka = sb + F.pk + 16u * (uint32_t)lane;
    const double Tk = A.temp[c];
    if (tid == 0) sts_f64(tail + 8, __ldcg(A.energy + c));  // running energy[1]: lives in shared memory, the retiring warp rotates
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t gchain = (uint32_t)(A.chain_offset + c);
    int32_t *gimg = A.img + (size_t)c * DIM * gNpad;
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * kSpecQCap);
    uint32_t slot = 0;
    const bool dbg_out = A.acc_out != nullptr || A.dE_out != nullptr;

    uint32_t rnd = 0;  // rounds so far: selects the retiring warp
    // species lists are only read by DiscreteSwap proposals (and by replayed ones); MoleculeFlip pools skip their upkeep
    bool keep_lists = A.replay != nullptr, lists_dirty = false;
    if constexpr (SWAPS) {
        for (int k = 0; k < A.n_moves; k++) keep_lists = keep_lists || A.mv_kind[k] == PMC_MOVE_SWAP;
    }
    for (long long tb = seg_lo; tb < seg_hi; tb += kSpecBatch) {
        const int nb = (int)min((long long)kSpecBatch, seg_hi - tb);
        __syncthreads();
        if (tid >= NT - 2 * PMC_MAX_MOVES) {  // fold the counters of the previous batch
            const int k = tid - (NT - 2 * PMC_MAX_MOVES);
            uint32_t *c32 = (uint32_t *)(smem_raw + F.cnt32);
            ((unsigned long long *)(smem_raw + F.cnt))[k] += c32[k];
            c32[k] = 0u;
            if (A.stats && k < 2) {  // work counters of the previous batch
                atomicAdd(A.stats + k, (unsigned long long)c32[2 * PMC_MAX_MOVES + k]);
                c32[2 * PMC_MAX_MOVES + k] = 0u;
            }
        }
        // ---- proposals of trials tb .. tb+nb-1, parked in shared memory (same stream as every other kernel) ----
        // generated by two warps (ceil(batch / 32)); which two alternates from batch to batch (PMC_SPEC_ROTATE)
#if PMC_SPEC_ROTATE
        const int pidx = (((warp - (int)(((tb - seg_lo) / kSpecBatch) * ((kSpecBatch + 31) / 32))) & (NW - 1)) << 5) | lane;
#else
        const int pidx = tid;
#endif
        if (pidx < nb) {
            const long long q = tb + pidx;
            pmc_trial tr;
            unsigned long long flip_pairs = 0ull;  // MoleculeFlip: four parked candidate pairs (site offsets in the molecule)
            if (A.replay) {
                tr = A.replay[(size_t)c * A.n_trials + q];
            } else {
                const unsigned long long t = A.t0 + (unsigned long long)q;
                const Philox4 a = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 0u, k0, k1);
                const Philox4 b = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 1u, k0, k1);
                const double um = (double)a.v[0] * 0x1p-32;
                int m = A.n_moves - 1;
                for (int k = A.n_moves - 2; k >= 0; k--)
                    if (um < A.mv_cum[k]) m = k;
                tr.u = uniform53(a.v[2], a.v[3]);
                tr.move = m;
                tr.kind = A.mv_kind[m];
                if (!SWAPS || tr.kind == PMC_MOVE_DISPLACEMENT) {
                    float z0, z1, z2, z3;
                    box_muller(b.v[0], b.v[1], z0, z1);
                    box_muller(b.v[2], b.v[3], z2, z3);
                    const float sg = A.mv_sigma[m];
                    tr.kind = PMC_MOVE_DISPLACEMENT;
                    tr.i = (int)bounded(a.v[1], (uint32_t)N);
                    tr.j = -1;
                    tr.delta[0] = (double)(sg * z0);
                    tr.delta[1] = (double)(sg * z1);
                    tr.delta[2] = (DIM == 3) ? (double)(sg * z2) : 0.0;
                } else if (MOL && tr.kind == PMC_MOVE_FLIP) {
                    // MoleculeFlip (src/moves.jl:344-352): a molecule uniformly, then ordered pairs of distinct sites until
                    // their species differ.  Species are only known when the trial is evaluated, so four candidate pairs
                    // are parked (same draws as the general kernel, chains.cu); the first unlike one is taken.
                    const int mol = A.n_mol > 0 ? (int)bounded(a.v[1], (uint32_t)A.n_mol) : 0;
                    const int len = A.n_mol > 0 ? __ldg(A.mol_len + mol) : 0;
                    tr.i = A.n_mol > 0 ? __ldg(A.mol_start + mol) : -1;
                    tr.j = -1;
                    const Philox4 c4 = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 2u, k0, k1);
                    const uint32_t w8[8] = {b.v[0], b.v[1], b.v[2], b.v[3], c4.v[0], c4.v[1], c4.v[2], c4.v[3]};
                    unsigned long long packed = 0;
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        uint32_t pa_ = 0, pb_ = 0;
                        if (len >= 2) {
                            pa_ = bounded(w8[2 * q4], (uint32_t)len);
                            pb_ = bounded(w8[2 * q4 + 1], (uint32_t)(len - 1));
                            pb_ += (pb_ >= pa_) ? 1u : 0u;
                        }
                        packed |= (unsigned long long)((pa_ & 0xFFu) | ((pb_ & 0xFFu) << 8)) << (16 * q4);
                    }
                    flip_pairs = len >= 2 && len <= 255 ? packed : 0ull;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                } else {  // slots in the species lists; resolved to particles when the trial is evaluated
                    const int *sso = (const int *)(smem_raw + F.spoff);
                    const int nA = sso[A.mv_a[m] + 1] - sso[A.mv_a[m]], nB = sso[A.mv_b[m] + 1] - sso[A.mv_b[m]];
                    tr.i = (nA > 0 && nB > 0) ? (int)bounded(a.v[1], (uint32_t)nA) : -1;
                    tr.j = (nA > 0 && nB > 0) ? (int)bounded(b.v[0], (uint32_t)nB) : -1;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                }
                if (A.trace) A.trace[(size_t)c * A.n_trials + q] = tr;
            }
            // record: f64 delta[3], f64 thr | s32 dint[3], s32 i | s32 m, pad[3] | u32 ~thr8[4]
            unsigned char *rec = smem_raw + F.rec + (size_t)kRecBytes * pidx;
            double *rd = (double *)rec;
            int *ri = (int *)(rec + 32);
            uint32_t *rt = (uint32_t *)(rec + 64);
            rd[0] = (MOL && SWAPS && tr.kind == PMC_MOVE_FLIP && !A.replay) ? __longlong_as_double((long long)flip_pairs) : tr.delta[0];
            rd[1] = tr.delta[1];
            rd[2] = tr.delta[2];
            rd[3] = A.exact_exp ? tr.u : -Tk * log(tr.u);
            ri[0] = (int)__double2ll_rn(tr.delta[0] * fscale);  // (delta is 0 for swaps and flips)
            ri[1] = (int)__double2ll_rn(tr.delta[1] * fscale);
            ri[2] = (int)__double2ll_rn(tr.delta[2] * fscale);
            ri[3] = tr.i;
            ri[4] = tr.move;
            ri[5] = tr.kind;
            ri[6] = tr.j;
            ri[7] = (tr.kind == PMC_MOVE_SWAP && !A.replay) ? (A.mv_a[tr.move] | (A.mv_b[tr.move] << 8)) : 0;
            const double hd = 0.5 * sqrt(tr.delta[0] * tr.delta[0] + tr.delta[1] * tr.delta[1] + tr.delta[2] * tr.delta[2]);
            const double *rcs = (const double *)(smem_raw + F.rcs);
#pragma unroll
            for (int s = 0; s < PMC_MAX_SPECIES; s++) rt[s] = neg_thr8((rcs[s] + hd) * fscale * 0x1p-24);
        }
        __syncthreads();

        int cur = 0;
        while (cur < nb) {  // one round: up to four consecutive trials, one per warp
            const int nspec = min(NW, nb - cur);
            const uint32_t pa = sb + F.pub + (uint32_t)(kPubBytes * NW) * slot;
            if (warp < nspec) {
                // ---- evaluate trial cur + warp against the current state (this warp alone) ---------------------
                const uint32_t ra = sb + F.rec + (uint32_t)kRecBytes * (uint32_t)(cur + warp);
                double d0, d1, d2, thr;
                int di0, di1, di2, i;
                lds_f64x2(ra, d0, d1);
                lds_f64x2(ra + 16, d2, thr);
                lds_s32x4(ra + 32, di0, di1, di2, i);
                bool is_swap = false;
                if constexpr (SWAPS) {
                    int mv, kind, j, sab;
                    lds_s32x4(ra + 48, mv, kind, j, sab);
                    if (kind == PMC_MOVE_SWAP || (MOL && kind == PMC_MOVE_FLIP)) {
                        is_swap = true;
                        // ---- DiscreteSwap / MoleculeFlip: positions fixed, four local energies in one pass
                        // (swap_particle_species!, src/moves.jl:159-167) ----
                        const uint32_t soa = sb + F.spoff;
                        if (MOL && kind == PMC_MOVE_FLIP && !A.replay) {
                            // the first parked pair whose sites carry different species NOW (src/moves.jl:347-350)
                            const unsigned long long packed = (unsigned long long)__double_as_longlong(d0);
                            const int base = i;
                            i = -1;
                            j = -1;
                            if (packed != 0ull && base >= 0) {
#pragma unroll
                                for (int q4 = 3; q4 >= 0; q4--) {
                                    const int sa_ = base + (int)((packed >> (16 * q4)) & 0xFFull), sb2 = base + (int)((packed >> (16 * q4 + 8)) & 0xFFull);
                                    if (lds_u8(sb + F.sp + (uint32_t)sa_) != lds_u8(sb + F.sp + (uint32_t)sb2)) {
                                        i = sa_;
                                        j = sb2;
                                    }
                                }
                            }
                        } else if (!A.replay && i >= 0) {  // slots -> particles through the species lists (current state)
                            const uint32_t oa = lds_u32(soa + 4u * (uint32_t)(sab & 0xFF)), ob = lds_u32(soa + 4u * (uint32_t)(sab >> 8));
                            i = (int)lds_u16(sb + F.spids + 2u * (oa + (uint32_t)i));
                            j = (int)lds_u16(sb + F.spids + 2u * (ob + (uint32_t)j));
                        }
                        const bool valid = i >= 0 && j >= 0;
                        const uint32_t iu = valid ? (uint32_t)i : 0u, ju = valid ? (uint32_t)j : 0u;
                        const uint32_t xia = sb + F.x + 8u * iu, xja = sb + F.x + 8u * ju;
                        const double xi0 = lds_f64(xia), xi1 = lds_f64(xia + nb8), xi2 = DIM == 3 ? lds_f64(xia + 2 * nb8) : 0.0;
                        const double xj0 = lds_f64(xja), xj1 = lds_f64(xja + nb8), xj2 = DIM == 3 ? lds_f64(xja + 2 * nb8) : 0.0;
                        const uint32_t si = lds_u8(sb + F.sp + iu), sj = lds_u8(sb + F.sp + ju);
                        const uint32_t qi = lds_u32(sb + F.pk + 4u * pk_pos(iu)), qj = lds_u32(sb + F.pk + 4u * pk_pos(ju));
                        const int gthr = (int)lds_u32(soa + 4u * (PMC_MAX_SPECIES + 1));
                        constexpr int NCHUNK = KCW / 4, NCH = NCHUNK >= 4 ? 4 : NCHUNK, CG = NCHUNK / NCH;
                        uint32_t m[NM];
                        int mine = 0;
                        // survivors of either sphere (TWO = true), or of the one sphere qc / cthr that contains both
                        auto scan = [&](auto two, uint32_t qc, int cthr) {
#pragma unroll
                            for (int mw = 0; mw < NM; mw++) {
                                uint32_t mc[NCH];
#pragma unroll
                                for (int h = 0; h < NCH; h++) mc[h] = 0u;
#pragma unroll
                                for (int cc = 0; cc < CG; cc++) {
#pragma unroll
                                    for (int h = 0; h < NCH; h++) {
                                        uint32_t w4[4];
                                        lds_u32x4(pka + 512u * (uint32_t)(mw * 8 + h * CG + cc), w4[0], w4[1], w4[2], w4[3]);
#pragma unroll
                                        for (int e = 0; e < 4; e++) {
                                            const uint32_t t1 = __vabsdiffu4(qc, w4[e]);
                                            int v = __dp4a((int)t1, (int)t1, cthr);
                                            if constexpr (decltype(two)::value) {
                                                const uint32_t t2 = __vabsdiffu4(qj, w4[e]);
                                                v |= __dp4a((int)t2, (int)t2, cthr);
                                            }
                                            mc[h] = __funnelshift_l((uint32_t)v, mc[h], 1);
                                        }
                                    }
                                }
                                uint32_t mm = mc[0];
#pragma unroll
                                for (int h = 1; h < NCH; h++) mm = (mm << (4 * CG)) | mc[h];
                                m[mw] = valid ? mm : 0u;
                                mine += __popc(m[mw]);
                            }
                        };
                        bool one_sphere = false;
                        if constexpr (MOL) one_sphere = kind == PMC_MOVE_FLIP;
                        if (one_sphere) {
                            // MoleculeFlip: the two sites sit a bond length apart, so ONE sphere around their midpoint with
                            // radius rc_max + |x_i - x_j| / 2 holds everything within rc_max of either (triangle inequality)
                            // at half the scan instructions; the midpoint and the half distance come from the wrapping
                            // 32-bit fixed-point difference (minimum image for free), like a Displacement's
                            const uint32_t ui0 = to_fixed32(xi0, fscale), ui1 = to_fixed32(xi1, fscale), ui2 = DIM == 3 ? to_fixed32(xi2, fscale) : 0u;
                            const int e0 = (int)(to_fixed32(xj0, fscale) - ui0), e1 = (int)(to_fixed32(xj1, fscale) - ui1);
                            const int e2 = DIM == 3 ? (int)(to_fixed32(xj2, fscale) - ui2) : 0;
                            const double hd = 0.5 * sqrt((double)e0 * (double)e0 + (double)e1 * (double)e1 + (double)e2 * (double)e2) * 0x1p-24;
                            const double rcu = lds_f64(soa + 24);
                            scan(std::false_type{}, pack8(ui0 + (uint32_t)(e0 >> 1), ui1 + (uint32_t)(e1 >> 1), ui2 + (uint32_t)(e2 >> 1)),
                                 (int)neg_thr8(rcu + hd + 0x1p-20));
                        } else {
                            scan(std::true_type{}, qi, gthr);
                        }
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int t = __shfl_up_sync(0xffffffffu, incl, o);
                            incl += (lane >= o) ? t : 0;
                        }
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        auto pair_e = [&](uint32_t sa, uint32_t sb_, double r2) -> double {
                            if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                                double rc2, eps4, sig2, shift;
                                const uint32_t pp = sb + F.cp + 32u * (sa * (uint32_t)ns + sb_);
                                lds_f64x2(pp, rc2, eps4);
                                lds_f64x2(pp + 16, sig2, shift);
                                return r2 <= rc2 ? lj_core(r2, eps4, sig2) - shift : 0.0;
                            } else {
                                const double *p = (const double *)(smem_raw + F.par) + (sa * (uint32_t)ns + sb_) * PMC_NPAR;
                                return r2 <= p[PMC_P_RCUT2] ? pair_potential<MODEL>(p, r2) : 0.0;
                            }
                        };
                        uint32_t bsi[PMC_MAX_BONDS], bsj[PMC_MAX_BONDS];  // MOL: bonded partners of i and of j (0xFFFF = none)
                        if constexpr (MOL) {
#pragma unroll
                            for (int k = 0; k < PMC_MAX_BONDS; k++) {
                                bsi[k] = (uint32_t)__ldg(A.bonds + (size_t)iu * PMC_MAX_BONDS + k);
                                bsj[k] = (uint32_t)__ldg(A.bonds + (size_t)ju * PMC_MAX_BONDS + k);
                            }
                        }
                        auto sterm = [&](uint32_t k) -> double {
                            double t = 0.0;
                            if (k < (uint32_t)N) {
                                const uint32_t ka = sb + F.x + 8u * k;
                                const double xk0 = lds_f64(ka), xk1 = lds_f64(ka + nb8), xk2 = DIM == 3 ? lds_f64(ka + 2 * nb8) : 0.0;
                                const uint32_t sk = lds_u8(sb + F.sp + k);
                                const uint32_t skn = k == iu ? sj : (k == ju ? si : sk);  // species of k after the exchange
                                bool pair_i = k != iu, pair_j = k != ju;
                                if constexpr (MOL) {  // bonded partners: bond pass below (they may lie outside both spheres)
#pragma unroll
                                    for (int b = 0; b < PMC_MAX_BONDS; b++) {
                                        pair_i = pair_i && k != bsi[b];
                                        pair_j = pair_j && k != bsj[b];
                                    }
                                }
                                if (pair_i) {  // k-term of particle i's local energy: (si, sk) -> (sj, sk')
                                    double r2 = mi_acc(xi0, xk0, L, hL, 0.0);
                                    r2 = mi_acc(xi1, xk1, L, hL, r2);
                                    if constexpr (DIM == 3) r2 = mi_acc(xi2, xk2, L, hL, r2);
                                    t += pair_e(sj, skn, r2) - pair_e(si, sk, r2);
                                }
                                if (pair_j) {  // k-term of particle j's local energy: (sj, sk) -> (si, sk')
                                    double r2 = mi_acc(xj0, xk0, L, hL, 0.0);
                                    r2 = mi_acc(xj1, xk1, L, hL, r2);
                                    if constexpr (DIM == 3) r2 = mi_acc(xj2, xk2, L, hL, r2);
                                    t += pair_e(si, skn, r2) - pair_e(sj, sk, r2);
                                }
                            }
                            return t;
                        };
                        double part = 0.0;
                        if (total <= kSpecQCap) {
                            uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
                            for (int mw = 0; mw < NM; mw++) {
                                uint32_t mm = m[mw];
                                while (mm) {
                                    const int b = 31 - __clz(mm);
                                    mm ^= 1u << b;
                                    sts_u16(wp, cand_index<KCW>(b, lane) + 1024u * (uint32_t)mw);
                                    wp += 2;
                                }
                            }
                            __syncwarp();
                            for (int q = lane; q < total; q += 32) part += sterm(lds_u16(qa + 2u * (uint32_t)q));
                            __syncwarp();
                        } else {
#pragma unroll
                            for (int mw = 0; mw < NM; mw++) {
                                uint32_t mm = m[mw];
                                while (mm) {
                                    const int b = 31 - __clz(mm);
                                    mm ^= 1u << b;
                                    part += sterm(cand_index<KCW>(b, lane) + 1024u * (uint32_t)mw);
                                }
                            }
                        }
                        if constexpr (MOL) {
                            // bond pass: lanes 0..5 own the bonded partners of i, lanes 6..11 those of j (FENE + bonded LJ,
                            // src/models.jl:219-226, with the species pair before and after the exchange)
                            const bool of_j = lane >= PMC_MAX_BONDS;
                            uint32_t b = 0xFFFFu;
#pragma unroll
                            for (int k = 0; k < PMC_MAX_BONDS; k++) {
                                if (lane == k) b = bsi[k];
                                if (lane == PMC_MAX_BONDS + k) b = bsj[k];
                            }
                            if (valid && lane < 2 * PMC_MAX_BONDS && b != 0xFFFFu) {
                                const uint32_t ka = sb + F.x + 8u * b;
                                const double xk0 = lds_f64(ka), xk1 = lds_f64(ka + nb8), xk2 = DIM == 3 ? lds_f64(ka + 2 * nb8) : 0.0;
                                const uint32_t sk = lds_u8(sb + F.sp + b);
                                const uint32_t skn = b == iu ? sj : (b == ju ? si : sk);
                                double r2 = mi_acc(of_j ? xj0 : xi0, xk0, L, hL, 0.0);
                                r2 = mi_acc(of_j ? xj1 : xi1, xk1, L, hL, r2);
                                if constexpr (DIM == 3) r2 = mi_acc(of_j ? xj2 : xi2, xk2, L, hL, r2);
                                const double *spar = (const double *)(smem_raw + F.par);
                                const double *po = spar + ((of_j ? sj : si) * (uint32_t)ns + sk) * PMC_NPAR;
                                const double *pn = spar + ((of_j ? si : sj) * (uint32_t)ns + skn) * PMC_NPAR;
                                part += bond_potential(pn, r2) - bond_potential(po, r2);
                            }
                        }
                        const double dE = warp_sum(part);
                        const bool acc = valid && (A.exact_exp ? accept_exact(dE, Tk, thr) : (dE < thr));
                        if (lane == 0) {
                            const uint32_t pw = pa + (uint32_t)kPubBytes * (uint32_t)warp;
                            sts_f64(pw, dE);
                            // spheres of a swap = its two particles; an invalid trial (empty species list) touches nothing
                            sts_u32x4(pw + 32, qi, valid ? (uint32_t)gthr : 0xFFFFFFFFu, qi, qj);
                            sts_u32x4(pw + 48, iu, (acc ? 1u : 0u) | 2u, ju, (uint32_t)mv);
                            if (A.trace) {
                                pmc_trial *tr = A.trace + (size_t)c * A.n_trials + tb + cur + warp;
                                tr->i = i;
                                tr->j = j;
                            }
                        }
                    }
                }
                if (!is_swap) {
                    const uint32_t si = lds_u8(sb + F.sp + (uint32_t)i);
                    double xo[3] = {0.0, 0.0, 0.0}, xn[3] = {0.0, 0.0, 0.0};
                    uint32_t uo0, uo1, uo2, un0 = 0u, un1 = 0u, un2 = 0u;
                    int wr0, wr1, wr2;  // image-counter increments of the move
                    if constexpr (MIXED) {
                        const uint32_t ua = sb + F.x + 4u * (uint32_t)i;
                        uo0 = lds_u32(ua);
                        uo1 = lds_u32(ua + nb4);
                        uo2 = (DIM == 3) ? lds_u32(ua + 2 * nb4) : 0u;
                        un0 = uo0 + (uint32_t)di0;  // wraps like the box does
                        un1 = uo1 + (uint32_t)di1;
                        un2 = uo2 + (uint32_t)di2;
                        wr0 = (di0 > 0 && un0 < uo0) - (di0 < 0 && un0 > uo0);
                        wr1 = (di1 > 0 && un1 < uo1) - (di1 < 0 && un1 > uo1);
                        wr2 = (di2 > 0 && un2 < uo2) - (di2 < 0 && un2 > uo2);
                    } else {
                        const uint32_t xa = sb + F.x + 8u * (uint32_t)i;
                        xo[0] = lds_f64(xa);
                        xo[1] = lds_f64(xa + nb8);
                        xo[2] = (DIM == 3) ? lds_f64(xa + 2 * nb8) : 0.0;
                        const double t0 = xo[0] + d0, t1 = xo[1] + d1, t2 = xo[2] + d2;
                        xn[0] = wrap_once(t0, L);
                        xn[1] = wrap_once(t1, L);
                        xn[2] = (DIM == 3) ? wrap_once(t2, L) : 0.0;
                        wr0 = (t0 >= L) - (t0 < 0.0);
                        wr1 = (t1 >= L) - (t1 < 0.0);
                        wr2 = (DIM == 3) ? (t2 >= L) - (t2 < 0.0) : 0;
                        uo0 = to_fixed32(xo[0], fscale);
                        uo1 = to_fixed32(xo[1], fscale);
                        uo2 = (DIM == 3) ? to_fixed32(xo[2], fscale) : 0u;
                    }
                    const uint32_t umq = pack8(uo0 + (uint32_t)(di0 >> 1), uo1 + (uint32_t)(di1 >> 1), uo2 + (uint32_t)(di2 >> 1));
                    const int fthr = (int)lds_u32(ra + 64 + 4u * si);
                    // survivor masks: candidate k = 32 * word + k' -> bit KCW-1-k' of its word; every word is built as
                    // independent shift chains over groups of chunks (the funnel shifts of one chain depend on each other)
                    constexpr int NCHUNK = KCW / 4, NCH = NCHUNK >= 4 ? 4 : NCHUNK, CG = NCHUNK / NCH;
                    uint32_t m[NM];
                    int mine = 0;
#pragma unroll
                    for (int mw = 0; mw < NM; mw++) {
                        uint32_t mc[NCH];
#pragma unroll
                        for (int h = 0; h < NCH; h++) mc[h] = 0u;
#pragma unroll
                        for (int cc = 0; cc < CG; cc++) {
#pragma unroll
                            for (int h = 0; h < NCH; h++) {
                                uint32_t w4[4];
                                lds_u32x4(pka + 512u * (uint32_t)(mw * 8 + h * CG + cc), w4[0], w4[1], w4[2], w4[3]);
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    const uint32_t t = __vabsdiffu4(umq, w4[e]);
                                    mc[h] = __funnelshift_l((uint32_t)__dp4a((int)t, (int)t, fthr), mc[h], 1);
                                }
                            }
                        }
                        uint32_t mm = mc[0];
#pragma unroll
                        for (int h = 1; h < NCH; h++) mm = (mm << (4 * CG)) | mc[h];
                        // padding slots and the moved particle itself (always inside its own sphere) leave here, so the
                        // fp64 pass needs no validity test per survivor
                        mm &= vmask[mw];
                        if ((uint32_t)lane == ((uint32_t)i & 31u) && (uint32_t)mw == ((uint32_t)i >> 10))
                            mm &= ~(1u << (KCW - 1 - (int)(((uint32_t)i >> 5) & 31u)));
                        m[mw] = mm;
                        mine += __popc(mm);
                    }
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        incl += (lane >= o) ? t : 0;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    if (A.stats && lane == 0) {
                        reds_add_u32(sb + F.cnt32 + 4u * (2 * PMC_MAX_MOVES), (uint32_t)total);
                        reds_add_u32(sb + F.cnt32 + 4u * (2 * PMC_MAX_MOVES + 1), 1u);
                    }
                    // Everything about this trial except its energy change is known here.  fp64: published now, off the
                    // critical path between the reduction and the barrier (dE and the decision follow after the pass; +1 %);
                    // PMC_MIXED (64 registers) publishes after the pass (early: -2.6 %).
                    // +0: dE, new position | +32: what the conflict test of LATER trials needs | +48: what retiring THIS trial needs
                    auto publish = [&]() {
                        const uint32_t pw = pa + (uint32_t)kPubBytes * (uint32_t)warp;
                        uint32_t qn;
                        if constexpr (MIXED) {
                            sts_u32x4(pw + 16, un0, un1, un2, 0u);
                            qn = pack8(un0, un1, un2);
                        } else {
                            sts_f64(pw + 8, xn[0]);
                            sts_f64x2(pw + 16, xn[1], xn[2]);
                            qn = pack8(to_fixed32(xn[0], fscale), to_fixed32(xn[1], fscale), (DIM == 3) ? to_fixed32(xn[2], fscale) : 0u);
                        }
                        sts_u32x4(pw + 32, umq, (uint32_t)fthr, pack8(uo0, uo1, uo2), qn);
                        sts_u32x4(pw + 48, (uint32_t)i, 0u, (uint32_t)((wr0 + 1) | ((wr1 + 1) << 2) | ((wr2 + 1) << 4)), lds_u32(ra + 48));
                    };
                    if constexpr (!MIXED) {
                        if (lane == 0) publish();
                    }
                    const uint32_t prow = si * (uint32_t)ns;
                    double part = 0.0;
                    uint32_t bi[PMC_MAX_BONDS];  // MOL: bonded partners of i (0xFFFF = none)
                    if constexpr (MOL) {
#pragma unroll
                        for (int k = 0; k < PMC_MAX_BONDS; k++) bi[k] = (uint32_t)__ldg(A.bonds + (size_t)i * PMC_MAX_BONDS + k);
                    }
                    // pair term of candidate j, branch-free (selects) so that two of them interleave in the unrolled loop
                    auto term = [&](uint32_t j) -> double {
                        bool valid = true;
                        if constexpr (MOL) {  // bonded partners are handled by the bond pass below
#pragma unroll
                            for (int k = 0; k < PMC_MAX_BONDS; k++) valid = valid && j != bi[k];
                        }
                        if constexpr (MIXED) {  // fp32 pair terms on wrapping integer distances (minimum image for free)
                            const uint32_t ja = sb + F.x + 4u * j;
                            const uint32_t a0 = lds_u32(ja), a1 = lds_u32(ja + nb4), a2 = (DIM == 3) ? lds_u32(ja + 2 * nb4) : 0u;
                            const float r2o = __uint2float_rn(dist2_u32<DIM>(uo0, uo1, uo2, a0, a1, a2)) * r2scale;
                            const float r2n = __uint2float_rn(dist2_u32<DIM>(un0, un1, un2, a0, a1, a2)) * r2scale;
                            const uint32_t sj = lds_u8(sb + F.sp + j);
                            float rc2, eps, sig2, shift, c0, c2, c4, pad_;
                            const uint32_t pp = sb + F.cp + 32u * (prow + sj);
                            lds_f32x4(pp, rc2, eps, sig2, shift);
                            lds_f32x4(pp + 16, c0, c2, c4, pad_);
                            const float eo = pair_potential_f32<MODEL>(r2o, eps, sig2, shift, c0, c2, c4);
                            const float en = pair_potential_f32<MODEL>(r2n, eps, sig2, shift, c0, c2, c4);
                            const float d = (r2n <= rc2 ? en : 0.0f) - (r2o <= rc2 ? eo : 0.0f);
                            return (double)d;
                        }
                        const uint32_t ja = sb + F.x + 8u * j;
                        const double xj0 = lds_f64(ja), xj1 = lds_f64(ja + nb8);
                        double r2o = mi_acc(xo[0], xj0, L, hL, 0.0), r2n = mi_acc(xn[0], xj0, L, hL, 0.0);
                        r2o = mi_acc(xo[1], xj1, L, hL, r2o);
                        r2n = mi_acc(xn[1], xj1, L, hL, r2n);
                        if constexpr (DIM == 3) {
                            const double xj2 = lds_f64(ja + 2 * nb8);
                            r2o = mi_acc(xo[2], xj2, L, hL, r2o);
                            r2n = mi_acc(xn[2], xj2, L, hL, r2n);
                        }
                        const uint32_t sj = lds_u8(sb + F.sp + j);
                        double uo, un, rc2;
                        if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                            double eps4, sig2, shift;
                            const uint32_t pp = sb + F.cp + 32u * (prow + sj);
                            lds_f64x2(pp, rc2, eps4);
                            lds_f64x2(pp + 16, sig2, shift);
                            uo = lj_core(r2o, eps4, sig2) - shift;
                            un = lj_core(r2n, eps4, sig2) - shift;
                        } else {
                            const double *p = (const double *)(smem_raw + F.par) + (prow + sj) * PMC_NPAR;
                            rc2 = p[PMC_P_RCUT2];
                            uo = pair_potential<MODEL>(p, r2o);
                            un = pair_potential<MODEL>(p, r2n);
                        }
                        const double d = (r2n <= rc2 ? un : 0.0) - (r2o <= rc2 ? uo : 0.0);
                        if constexpr (MOL) return valid ? d : 0.0;
                        return d;
                    };
                    if (total <= kSpecQCap) {
                        // compaction: each lane appends its survivors (ascending candidate index) at its scan offset
                        uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
                        for (int mw = 0; mw < NM; mw++) {
                            uint32_t mm = m[mw];
                            while (mm) {
                                const int b = 31 - __clz(mm);
                                mm ^= 1u << b;
                                sts_u16(wp, cand_index<KCW>(b, lane) + 1024u * (uint32_t)mw);
                                wp += 2;
                            }
                        }
                        __syncwarp();
                        // two survivors per lane and iteration while both exist (the dependent fp64 chains of one pair
                        // term leave the pipe idle; two independent ones overlap), then the remainder one at a time
                        const int nfull = total & ~63;
                        int q = lane;
                        for (; q < nfull; q += 64) {
                            const uint32_t j0 = lds_u16(qa + 2u * (uint32_t)q), j1 = lds_u16(qa + 2u * (uint32_t)q + 64u);
                            const double e0 = term(j0), e1 = term(j1);
                            part += e0;
                            part += e1;
                        }
                        for (; q < total; q += 32) part += term(lds_u16(qa + 2u * (uint32_t)q));
                        __syncwarp();
                    } else {  // tiny boxes where (nearly) every candidate survives: no queue, each lane its own survivors
#pragma unroll
                        for (int mw = 0; mw < NM; mw++) {
                            uint32_t mm = m[mw];
                            while (mm) {
                                const int b = 31 - __clz(mm);
                                mm ^= 1u << b;
                                part += term(cand_index<KCW>(b, lane) + 1024u * (uint32_t)mw);
                            }
                        }
                    }
                    if constexpr (MOL) {  // bond pass: lane k owns bonded partner k of particle i
                        const uint32_t b = lane < PMC_MAX_BONDS ? (uint32_t)__ldg(A.bonds + (size_t)i * PMC_MAX_BONDS + lane) : 0xFFFFu;
                        if (b != 0xFFFFu) {
                            const uint32_t ja = sb + F.x + 8u * b;
                            const double xj0 = lds_f64(ja), xj1 = lds_f64(ja + nb8);
                            double r2o = mi_acc(xo[0], xj0, L, hL, 0.0), r2n = mi_acc(xn[0], xj0, L, hL, 0.0);
                            r2o = mi_acc(xo[1], xj1, L, hL, r2o);
                            r2n = mi_acc(xn[1], xj1, L, hL, r2n);
                            if constexpr (DIM == 3) {
                                const double xj2 = lds_f64(ja + 2 * nb8);
                                r2o = mi_acc(xo[2], xj2, L, hL, r2o);
                                r2n = mi_acc(xn[2], xj2, L, hL, r2n);
                            }
                            const double *p = (const double *)(smem_raw + F.par) + (prow + lds_u8(sb + F.sp + b)) * PMC_NPAR;
                            part += bond_potential(p, r2n) - bond_potential(p, r2o);
                        }
                    }
                    const double dE = warp_sum(part);
                    const bool acc = A.exact_exp ? accept_exact(dE, Tk, thr) : (dE < thr);
                    if (lane == 0) {
                        if constexpr (MIXED) publish();
                        const uint32_t pw = pa + (uint32_t)kPubBytes * (uint32_t)warp;
                        sts_f64(pw, dE);
                        sts_u32(pw + 52, acc ? 1u : 0u);
                    }
                }  // !is_swap
            }
            __syncthreads();
            // ---- retire the round in trial order: ONE warp, the others wait at the second barrier ------------------
#if PMC_SPEC_ROTATE
            const int rw = (int)(rnd & (uint32_t)(NW - 1));
#else
            constexpr int rw = 0;
#endif
            rnd++;
            if (warp == rw) {
                // The retiring warp works on all trials of the round at once: G = 32 / NW lanes per trial.  Lane (w, v) tests
                // trial w against the EARLIER trial v (accepted, old or new position inside w's filter sphere -- the very
                // test that defines w's survivors); one ballot gives the first trial that does not stand, i.e. the number
                // retired.  Accepted standing trials of one round touch different particles outside each other's spheres,
                // so their commits are independent stores issued side by side by the lanes of their groups; only the
                // running energy is summed in trial order (one lane).  The round's critical path is one dependent load /
                // test / ballot / store sequence instead of NW of them back to back.
                constexpr int G = 32 / NW;
                const int w = lane / G, sub = lane % G;
                const uint32_t pw = pa + (uint32_t)kPubBytes * (uint32_t)w;
                uint32_t umq, fthr, qo, qn, iw, fl, wr, mv;
                lds_u32x4(pw + 32, umq, fthr, qo, qn);
                lds_u32x4(pw + 48, iw, fl, wr, mv);
                const bool live = w < nspec;
                const bool wswap = SWAPS && (fl & 2u) != 0u;
                int conflict = 0;
                for (int v = sub; v < w; v += G) {
                    const uint32_t pv = pa + (uint32_t)kPubBytes * (uint32_t)v;
                    uint32_t a_, b_, qov, qnv, iv, flv, c_, d_;
                    lds_u32x4(pv + 32, a_, b_, qov, qnv);
                    lds_u32x4(pv + 48, iv, flv, c_, d_);
                    if (live && (flv & 1u)) {
                        const uint32_t ta = __vabsdiffu4(umq, qov), tb_ = __vabsdiffu4(umq, qnv);
                        conflict |= __dp4a((int)ta, (int)ta, (int)fthr) | __dp4a((int)tb_, (int)tb_, (int)fthr);
                        if (wswap) {  // second sphere of a swap: around its particle j; the species lists changed under it
                            const uint32_t tc = __vabsdiffu4(qn, qov), td = __vabsdiffu4(qn, qnv);
                            conflict |= __dp4a((int)tc, (int)tc, (int)fthr) | __dp4a((int)td, (int)td, (int)fthr);
                            if (flv & 2u) conflict = -1;
                        }
                        if constexpr (MOL) {
                            // a bonded partner that moved (or changed species) changes the bond term even from outside the
                            // pair cutoff sphere (FENE bonds reach r0 > rc): src/molecules.jl:160-176.  A flip has two
                            // particles on either side: iw / wr (= j of w) against iv / c_ (= j of v).
                            const bool vswap = SWAPS && (flv & 2u) != 0u;
#pragma unroll
                            for (int k = 0; k < PMC_MAX_BONDS; k++) {
                                const uint32_t bw = (uint32_t)__ldg(A.bonds + (size_t)iw * PMC_MAX_BONDS + k);
                                if (bw == iv || (vswap && bw == c_)) conflict = -1;
                                if (wswap) {
                                    const uint32_t bw2 = (uint32_t)__ldg(A.bonds + (size_t)wr * PMC_MAX_BONDS + k);
                                    if (bw2 == iv || (vswap && bw2 == c_)) conflict = -1;
                                }
                            }
                        }
                    }
                }
                // what the commit may need is fetched BEFORE the ballots (independent loads, in flight together): lane
                // `sub` < DIM of a group holds coordinate `sub` of its trial's new position, lane 0 the energy changes
                double cval = 0.0;
                uint32_t cval32 = 0u;
                if constexpr (MIXED) {
                    if (sub < DIM) cval32 = lds_u32(pw + 16u + 4u * (uint32_t)sub);
                } else {
                    if (sub < DIM) cval = lds_f64(pw + 8u + 8u * (uint32_t)sub);
                }
                double dEs[NW];
#pragma unroll
                for (int v = 0; v < NW; v++) dEs[v] = lane == 0 ? lds_f64(pa + (uint32_t)kPubBytes * (uint32_t)v) : 0.0;
                const double E0 = lane == 0 ? lds_f64(tail + 8) : 0.0;
                const unsigned cb = __ballot_sync(0xffffffffu, live && conflict < 0);
                const int ndone = cb ? min(nspec, (__ffs((int)cb) - 1) / G) : nspec;
                const bool mine = live && w < ndone;
                const bool acc = mine && (fl & 1u) != 0u;
                const unsigned accb = __ballot_sync(0xffffffffu, acc && sub == 0);
                if (acc && !wswap) {  // one store per lane: coordinates, packed word; rarely an image counter
                    if (sub < DIM) {
                        if constexpr (MIXED) sts_u32(sb + F.x + 4u * iw + (uint32_t)sub * nb4, cval32);
                        else sts_f64(sb + F.x + 8u * iw + (uint32_t)sub * nb8, cval);
                    } else if (sub == 3) {
                        sts_u32(sb + F.pk + 4u * pk_pos(iw), qn);
                        if (wr != 0x15u) {  // some coordinate wrapped around the box
                            const int w0 = (int)(wr & 3u) - 1, w1 = (int)((wr >> 2) & 3u) - 1, w2 = (int)((wr >> 4) & 3u) - 1;
                            if (w0) atomicAdd(&gimg[iw], w0);
                            if (w1) atomicAdd(&gimg[gNpad + iw], w1);
                            if (DIM == 3 && w2) atomicAdd(&gimg[2 * gNpad + iw], w2);
                        }
                    }
                }
                if (mine && sub == (4 % G)) {
                    reds_add_u32(sb + F.cnt32 + 4u * mv, 1u);
                    if (acc) reds_add_u32(sb + F.cnt32 + 4u * (PMC_MAX_MOVES + mv), 1u);
                    if (dbg_out) {
                        if (A.acc_out) A.acc_out[(size_t)c * A.n_trials + tb + cur + w] = acc ? 1 : 0;
                        if (A.dE_out) A.dE_out[(size_t)c * A.n_trials + tb + cur + w] = lds_f64(pw);
                    }
                }
                if (lane == 0) {  // energy[1] += dE in trial order (src/moves.jl:11-20)
                    double E = E0;
#pragma unroll
                    for (int v = 0; v < NW; v++)
                        if (accb & (1u << (G * v))) E += dEs[v];
                    sts_f64(tail + 8, E);
                    sts_u32(tail, (uint32_t)ndone);
                }
                if constexpr (SWAPS) {
                    // accepted swaps: update_species_list! (src/moves.jl:175-179), one at a time -- finding i and j in
                    // their species lists is a search by the whole warp, paid only by accepted swaps
                    unsigned sw = __ballot_sync(0xffffffffu, acc && wswap && sub == 0);
                    while (sw) {
                        const int src = __ffs((int)sw) - 1;
                        sw &= sw - 1u;
                        const uint32_t is_ = __shfl_sync(0xffffffffu, iw, src), js_ = __shfl_sync(0xffffffffu, wr, src);
                        const uint32_t si = lds_u8(sb + F.sp + is_), sj = lds_u8(sb + F.sp + js_);
                        if (!keep_lists) {  // no DiscreteSwap in the pool: nobody reads the lists in this launch
                            lists_dirty = true;
                            __syncwarp();  // every lane has read the two species before lane 0 exchanges them
                            if (lane == 0) {
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + is_), "r"(sj) : "memory");
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + js_), "r"(si) : "memory");
                            }
                            continue;
                        }
                        const uint32_t oi = lds_u32(sb + F.spoff + 4u * si), oj = lds_u32(sb + F.spoff + 4u * sj);
                        const uint32_t ni = lds_u32(sb + F.spoff + 4u * si + 4u) - oi, nj = lds_u32(sb + F.spoff + 4u * sj + 4u) - oj;
                        auto find = [&](uint32_t off, uint32_t n, uint32_t who) -> uint32_t {
                            uint32_t pos = 0;
                            for (uint32_t b0 = 0; b0 < n; b0 += 32) {
                                const uint32_t k = b0 + (uint32_t)lane;
                                const bool hit = k < n && lds_u16(sb + F.spids + 2u * (off + k)) == who;
                                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                                if (bal) pos = b0 + (uint32_t)__ffs((int)bal) - 1u;
                            }
                            return pos;
                        };
                        const uint32_t hi = find(oi, ni, is_), hj = find(oj, nj, js_);
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + is_), "r"(sj) : "memory");
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + js_), "r"(si) : "memory");
                            sts_u16(sb + F.spids + 2u * (oi + hi), js_);
                            sts_u16(sb + F.spids + 2u * (oj + hj), is_);
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
            cur += (int)lds_u32(tail);
            slot ^= 1u;
        }
    }
    __syncthreads();
    {
        if constexpr (MIXED) {
            // back to float64 at the centre of the fixed-point cell: re-quantising it gives the same integer again
            const uint32_t *su = (const uint32_t *)(smem_raw + F.x);
            const double inv = L * 0x1p-32;
            for (int a = 0; a < DIM; a++)
                for (int k = tid; k < gNpad; k += NT) gx[a * gNpad + k] = ((double)su[a * Npad + k] + 0.5) * inv;
        } else {
            const double *sx = (const double *)(smem_raw + F.x);
            for (int a = 0; a < DIM; a++)
                for (int k = tid; k < gNpad; k += NT) gx[a * gNpad + k] = sx[a * Npad + k];
        }
        if constexpr (SWAPS) {
            uint8_t *gsp = A.sp + (size_t)c * gNpad;
            uint16_t *gi = A.spids + (size_t)c * gNpad, *gh = A.heads + (size_t)c * gNpad;
            const uint16_t *si_ = (const uint16_t *)(smem_raw + F.spids);
            const int *sso = (const int *)(smem_raw + F.spoff);
            if constexpr (MOL) {
                // accepted flips of a pool without DiscreteSwap left the lists alone (the reference's Molecules carry
                // none, src/molecules.jl:24-41): rebuild them from the species, ids ascending per species as at upload,
                // so that a later pool with DiscreteSwap finds them consistent.  Counts per species are unchanged.
                if (__syncthreads_or(lists_dirty ? 1 : 0)) {
                    if (warp == 0) {
                        uint16_t *wl = (uint16_t *)(smem_raw + F.spids);
                        for (int sp_ = 0; sp_ < ns; sp_++) {
                            int pos = sso[sp_];
                            for (int b0 = 0; b0 < N; b0 += 32) {
                                const int k = b0 + lane;
                                const bool f = k < N && smem_raw[F.sp + k] == (unsigned char)sp_;
                                const unsigned bal = __ballot_sync(0xffffffffu, f);
                                if (f) wl[pos + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)k;
                                pos += __popc(bal);
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            for (int k = tid; k < gNpad; k += NT) {
                gsp[k] = smem_raw[F.sp + k];
                gi[k] = si_[k];
            }
            for (int k = tid; k < N; k += NT) {  // heads[particle] = its position inside its species list
                const int who = si_[k];
                gh[who] = (uint16_t)(k - sso[smem_raw[F.sp + who]]);
            }
        }
        const unsigned long long *scnt = (const unsigned long long *)(smem_raw + F.cnt);
        const uint32_t *c32 = (const uint32_t *)(smem_raw + F.cnt32);
        if (tid == 0) A.energy[c] = lds_f64(tail + 8);
        if (A.stats && tid < 2) atomicAdd(A.stats + tid, (unsigned long long)c32[2 * PMC_MAX_MOVES + tid]);
        if (tid < A.n_moves) {
            atomicAdd(A.calls + (size_t)c * PMC_MAX_MOVES + tid, scnt[tid] + c32[tid]);
            atomicAdd(A.accepted + (size_t)c * PMC_MAX_MOVES + tid, scnt[PMC_MAX_MOVES + tid] + c32[PMC_MAX_MOVES + tid]);
        }
    }
    if (!A.queue) break;
    __threadfence();  // this segment's state is visible before its completion is
    __syncthreads();
    if (tid == 0) *(volatile int32_t *)(A.queue + 1 + c) = seg + 1;
  }
}

// ==================================================================================================
// Local energies + total through the same 8-bit prefilter (System(): src/atoms.jl:51-52, :81-88).  One CTA per
// chain, one warp per particle i: scan all candidates against ONE sphere (x_i, rc_max(species i) + quantisation
// margin), compact the survivors, fp64 pair terms for survivors only, shuffle reduction.  Same pairs inside the
// cutoff as the direct kernel (k_chain_energy), a third of its instructions.  Atoms, cubic box, N <= 1024.
// ==================================================================================================
struct EnergyLayout {
    uint32_t x, sp, pk, q, cp, par, thr, e, total;
};
__host__ __device__ inline EnergyLayout energy_layout(int dim, int Npad, bool full_par) {
    EnergyLayout f;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        uint32_t p = o;
        o += (bytes + 15u) & ~15u;
        return p;
    };
    f.x = take(8u * dim * Npad);
    f.sp = take(Npad);
    f.pk = take(4u * Npad);
    f.q = take(2u * kSpecQCap * kSpecWarps);
    f.cp = take(32u * PMC_MAX_SPECIES * PMC_MAX_SPECIES);
    f.par = take(full_par ? 8u * PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR : 0u);
    f.thr = take(4u * PMC_MAX_SPECIES);
    f.e = take(8u * Npad);
    f.total = o;
    return f;
}

template <int DIM, int MODEL, int NPAD>
__global__ void __launch_bounds__(kSpecThreads, 5) k_chain_energy_fast(const __grid_constant__ EnergyArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KC = NPAD / 32, Npad = NPAD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, N = A.N, gNpad = A.Npad, ns = A.ns;
    constexpr bool kFullPar = !(MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG);
    const EnergyLayout F = energy_layout(DIM, Npad, kFullPar);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t nb8 = 8u * (uint32_t)Npad;
    const double L = A.box[c * 3], hL = 0.5 * L;
    const double fscale = 4294967296.0 / L;
    const double *gx = A.x + (size_t)c * DIM * gNpad;
    {
        double *sx = (double *)(smem_raw + F.x);
        for (int a = 0; a < DIM; a++)
            for (int k = tid; k < Npad; k += kSpecThreads) sx[a * Npad + k] = k < gNpad ? gx[a * gNpad + k] : 0.0;
        for (int k = tid; k < Npad; k += kSpecThreads) smem_raw[F.sp + k] = k < gNpad ? A.sp[(size_t)c * gNpad + k] : 0;
        for (int j = tid; j < Npad; j += kSpecThreads) {
            uint32_t u[3] = {0u, 0u, 0u};
#pragma unroll
            for (int a = 0; a < DIM; a++) u[a] = j < gNpad ? to_fixed32(gx[a * gNpad + j], fscale) : 0u;
            ((uint32_t *)(smem_raw + F.pk))[pk_pos((uint32_t)j)] = pack8(u[0], u[1], u[2]);
        }
        double *scp = (double *)(smem_raw + F.cp);
        if constexpr (kFullPar) {
            double *spar = (double *)(smem_raw + F.par);
            for (int k = tid; k < ns * ns * PMC_NPAR; k += kSpecThreads) spar[k] = A.par[k];
        }
        for (int k = tid; k < ns * ns; k += kSpecThreads) {
            scp[4 * k + 0] = A.par[k * PMC_NPAR + PMC_P_RCUT2];
            scp[4 * k + 1] = A.par[k * PMC_NPAR + PMC_P_EPS];
            scp[4 * k + 2] = A.par[k * PMC_NPAR + PMC_P_SIG2];
            scp[4 * k + 3] = A.par[k * PMC_NPAR + PMC_P_SHIFT];
        }
        if (tid < PMC_MAX_SPECIES) {  // filter threshold per species of particle i
            double rc2 = 0.0;
            for (int b = 0; b < ns; b++) rc2 = fmax(rc2, A.par[((tid < ns ? tid : 0) * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            ((uint32_t *)(smem_raw + F.thr))[tid] = neg_thr8(sqrt(rc2) * fscale * 0x1p-24);
        }
    }
    __syncthreads();
    const uint32_t pka = sb + F.pk + 16u * (uint32_t)lane;
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * kSpecQCap);
    double *se = (double *)(smem_raw + F.e);
    for (int i = warp; i < N; i += kSpecWarps) {
        const uint32_t xa = sb + F.x + 8u * (uint32_t)i;
        const double x0 = lds_f64(xa), x1 = lds_f64(xa + nb8), x2 = (DIM == 3) ? lds_f64(xa + 2 * nb8) : 0.0;
        const uint32_t si = lds_u8(sb + F.sp + (uint32_t)i);
        const uint32_t uq = lds_u32(sb + F.pk + 4u * pk_pos((uint32_t)i));
        const int fthr = (int)lds_u32(sb + F.thr + 4u * si);
        constexpr int NCHUNK = KC / 4, NCH = NCHUNK >= 4 ? 4 : NCHUNK, CG = NCHUNK / NCH;
        uint32_t mc[NCH];
#pragma unroll
        for (int h = 0; h < NCH; h++) mc[h] = 0u;
#pragma unroll
        for (int cc = 0; cc < CG; cc++) {
#pragma unroll
            for (int h = 0; h < NCH; h++) {
                uint32_t w4[4];
                lds_u32x4(pka + 512u * (uint32_t)(h * CG + cc), w4[0], w4[1], w4[2], w4[3]);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const uint32_t t = __vabsdiffu4(uq, w4[e]);
                    mc[h] = __funnelshift_l((uint32_t)__dp4a((int)t, (int)t, fthr), mc[h], 1);
                }
            }
        }
        uint32_t m = mc[0];
#pragma unroll
        for (int h = 1; h < NCH; h++) m = (m << (4 * CG)) | mc[h];
        const int mine = __popc(m);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            incl += (lane >= o) ? t : 0;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t prow = si * (uint32_t)ns;
        auto term = [&](uint32_t j) -> double {
            const bool valid = j < (uint32_t)N && j != (uint32_t)i;
            const uint32_t ja = sb + F.x + 8u * j;
            double r2 = mi_acc(x0, lds_f64(ja), L, hL, 0.0);
            r2 = mi_acc(x1, lds_f64(ja + nb8), L, hL, r2);
            if constexpr (DIM == 3) r2 = mi_acc(x2, lds_f64(ja + 2 * nb8), L, hL, r2);
            const uint32_t sj = lds_u8(sb + F.sp + j);
            double u, rc2;
            if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                double eps4, sig2, shift;
                const uint32_t pp = sb + F.cp + 32u * (prow + sj);
                lds_f64x2(pp, rc2, eps4);
                lds_f64x2(pp + 16, sig2, shift);
                u = lj_core(r2, eps4, sig2) - shift;
            } else {
                const double *p = (const double *)(smem_raw + F.par) + (prow + sj) * PMC_NPAR;
                rc2 = p[PMC_P_RCUT2];
                u = pair_potential<MODEL>(p, r2);
            }
            return (valid && r2 <= rc2) ? u : 0.0;
        };
        double part = 0.0;
        if (total <= kSpecQCap) {
            uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
            uint32_t mm = m;
            while (mm) {
                const int b = 31 - __clz(mm);
                mm ^= 1u << b;
                sts_u16(wp, cand_index<KC>(b, lane));
                wp += 2;
            }
            __syncwarp();
            for (int q = lane; q < total; q += 32) part += term(lds_u16(qa + 2u * (uint32_t)q));
            __syncwarp();
        } else {
            uint32_t mm = m;
            while (mm) {
                const int b = 31 - __clz(mm);
                mm ^= 1u << b;
                part += term(cand_index<KC>(b, lane));
            }
        }
        const double e = warp_sum(part);
        if (lane == 0) {
            se[i] = e;
            A.eloc[(size_t)c * gNpad + i] = e;
        }
    }
    __syncthreads();
    // deterministic block reduction of se[0..N)
    double s_ = 0.0;
    for (int k = tid; k < N; k += kSpecThreads) s_ += se[k];
    s_ = warp_sum(s_);
    __syncthreads();
    if (lane == 0) se[warp] = s_;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < kSpecWarps; w++) tot += se[w];
        A.etot[c] = tot / 2;
    }
}

}  // namespace spec
}  // namespace pmc
