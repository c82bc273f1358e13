// chains.cu -- PMC_MODE_CHAINS kernels: many independent Metropolis chains, one chain per CTA, the
// whole chain state resident in shared memory for the duration of a launch.
//
// Replaces, for every chain at once, the reference's sequential loop
//   mc_sweep! -> mc_step! -> sample_action! / perform_action! / revert_action!
// (benchmark/particles_benchmarks.jl:28-29; src/moves.jl:57-90 Displacement, :159-207 DiscreteSwap;
//  src/atoms.jl:66-88 and src/molecules.jl:163-215 local energy; src/utils.jl:8-10 acceptance).
//
// Design:
//   * the chain is strictly sequential, so all parallelism inside a chain is over the candidate
//     partners j of the moved particle: thread t owns particles t, t+NT, ... ; one warp-shuffle
//     reduction + ONE __syncthreads per trial;
//   * old and new local energies are evaluated in the same pass over j (shared loads / parameters);
//   * proposals of NT trials are generated in parallel (thread b -> trial b of the batch) from
//     counter-based Philox and parked in shared memory, so no RNG work sits on the serial path;
//   * rejected moves are simply not committed (the reference's add-back x+d-d is a no-op physically).
#include <cuda_runtime.h>

#include <type_traits>

#include "chains.cuh"
#include "common.cuh"
#include "rng.cuh"

namespace pmc {

namespace {

constexpr int kMaxThreads = 256;  // CTA size limit of the sweep kernel
constexpr int kMaxWarps = kMaxThreads / 32;
constexpr int kBatch = 32;    // proposals generated per batch (parked in shared memory)
constexpr int kRegCand = 8;   // FILTER: candidates per thread whose fixed-point coordinates live in registers

// One parked proposal (what sample_action! draws + the acceptance threshold), 80 bytes.
struct __align__(16) TrialRec {
    double delta[3];
    double thr;                       // -T*log(u), or u itself in exact_exp mode
    int dint[3];                      // FILTER: delta in fixed-point units
    int i;                            // particle i, or slot ka in the species-A list for swaps
    int j;                            // slot kb for swaps, -2 for displacements
    int m;                            // pool index
    int pad[2];                       // MoleculeFlip: four parked (site a, site b) pairs, 8 bits each
    uint32_t thr_t[PMC_MAX_SPECIES];  // FILTER: (rc_s + |delta|/2)^2 in fixed-point units
    int flip;                         // 1: MoleculeFlip whose sites are resolved from `pad` when the trial executes
    int pad2[3];
};

// Fixed-size part of the CTA state: statically allocated so every access is a constant shared-memory
// offset (no pointer registers on the serial path).
struct __align__(16) SweepStatic {
    TrialRec trial[kBatch];
    double par[PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR];
    double red[2][kMaxWarps];
    double rcs[PMC_MAX_SPECIES];            // FILTER: largest cutoff radius per species of the moved particle
    unsigned long long cnt[2][PMC_MAX_MOVES];  // calls, accepted
    uint32_t thr_u[PMC_MAX_SPECIES + 4];    // FILTER: cutoff^2 in fixed-point units per species; [4] = global
    int spoff[PMC_MAX_SPECIES + 4];
};

// N-dependent part, carved from dynamic shared memory.
struct SweepDyn {
    double *x;        // [DIM][Npad] wrapped positions (source of truth)
    uint32_t *u;      // [DIM][Npad] FILTER: positions as 32-bit fixed-point fractions of L
    uint16_t *wq;     // [nwarp][qcap] FILTER: per-warp queues of surviving candidates
    uint16_t *spids;  // [Npad] SpeciesList ids grouped by species (swaps)
    uint16_t *heads;  // [Npad]
    uint16_t *bonds;  // [Npad][PMC_MAX_BONDS] (MOL only)
    uint8_t *sp;      // [Npad]
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int queue_cap(int Npad, int NT) { return (Npad + NT - 1) / NT * 32; }

__host__ __device__ inline size_t carve_sweep(SweepDyn &s, unsigned char *base, int dim, int Npad, int NT, bool mol,
                                              bool any_swap, bool filter) {
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char *p = base + o;
        o += align_up(bytes, 16);
        return p;
    };
    s.x = (double *)take(sizeof(double) * dim * Npad);
    s.u = (uint32_t *)take(filter ? sizeof(uint32_t) * dim * Npad : 0);
    s.wq = (uint16_t *)take(filter ? sizeof(uint16_t) * (NT / 32) * queue_cap(Npad, NT) : 0);
    s.spids = (uint16_t *)take(any_swap ? sizeof(uint16_t) * Npad : 0);
    s.heads = (uint16_t *)take(any_swap ? sizeof(uint16_t) * Npad : 0);
    s.bonds = (uint16_t *)take(mol ? sizeof(uint16_t) * Npad * PMC_MAX_BONDS : 0);
    s.sp = (uint8_t *)take(Npad);
    return o;
}

template <int DIM>
__device__ __forceinline__ double dist2(const double *__restrict__ sx, int Npad, int j, const double (&xi)[3],
                                        const double (&L)[3]) {
    double r2 = mi_sq(xi[0], sx[j], L[0]);
    r2 += mi_sq(xi[1], sx[Npad + j], L[1]);
    if constexpr (DIM == 3) r2 += mi_sq(xi[2], sx[2 * Npad + j], L[2]);
    return r2;
}

// Squared nearest-image separation in fixed-point units: coordinates are 32-bit fractions of the (cubic)
// box, so the wrapping subtraction IS the minimum image; hi32(d*d) summed over axes is r^2 * 2^32 / L^2.
template <int DIM>
__device__ __forceinline__ uint32_t dist2_fixed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t a0, uint32_t a1, uint32_t a2) {
    int d = (int)(c0 - a0);
    uint32_t r = (uint32_t)__mulhi(d, d);
    d = (int)(c1 - a1);
    r += (uint32_t)__mulhi(d, d);
    if constexpr (DIM == 3) {
        d = (int)(c2 - a2);
        r += (uint32_t)__mulhi(d, d);
    }
    return r;
}

__device__ __forceinline__ uint32_t to_fixed(double x, double scale) {  // scale = 2^32 / L
    return (uint32_t)__double2ull_rd(x * scale);
}

__device__ __forceinline__ uint32_t fixed_threshold(double r2_scaled) {  // r^2 * 2^32 / L^2, + rounding slack
    const double t = r2_scaled * (1.0 + 1e-9) + 64.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

template <bool MOL>
__device__ __forceinline__ bool bonded_to(const uint16_t (&bi)[PMC_MAX_BONDS], int j) {
    if constexpr (!MOL) return false;
    bool b = false;
#pragma unroll
    for (int k = 0; k < PMC_MAX_BONDS; k++) b |= (bi[k] == (uint16_t)j);
    return b;
}

// Contribution of candidate j to e2 - e1 of a Displacement of particle i (old position xo, new xn).
template <int DIM, int MODEL, bool MOL>
__device__ __forceinline__ double displacement_term(const SweepDyn &S, int Npad, int j, const double (&xo)[3],
                                                    const double (&xn)[3], const double (&L)[3],
                                                    const double *__restrict__ prow,
                                                    const uint16_t (&bi)[PMC_MAX_BONDS]) {
    const double *p = prow + S.sp[j] * PMC_NPAR;
    const double r2o = dist2<DIM>(S.x, Npad, j, xo, L);
    const double r2n = dist2<DIM>(S.x, Npad, j, xn, L);
    if (bonded_to<MOL>(bi, j)) return bond_potential(p, r2n) - bond_potential(p, r2o);
    const double rc2 = p[PMC_P_RCUT2];
    double t = 0.0;
    if (r2o <= rc2) t -= pair_potential<MODEL>(p, r2o);
    if (r2n <= rc2) t += pair_potential<MODEL>(p, r2n);
    return t;
}

// Contribution of particle k to e2 - e1 of a DiscreteSwap between i (species si) and j (species sj):
// the k-terms of both local energies before and after the exchange (src/moves.jl:159-167).
template <int DIM, int MODEL, bool MOL>
__device__ __forceinline__ double swap_term(const SweepDyn &S, const double *__restrict__ par, int Npad, int ns, int k,
                                            int i, int j, int si, int sj, const double (&xi)[3], const double (&xj)[3],
                                            const double (&L)[3], const uint16_t (&bi)[PMC_MAX_BONDS],
                                            const uint16_t (&bj)[PMC_MAX_BONDS]) {
    const int sk_old = S.sp[k];
    int sk_new = sk_old;
    if (k == i) sk_new = sj;
    if (k == j) sk_new = si;
    double t = 0.0;
    if (k != i) {  // term of particle i's local energy
        const double r2 = dist2<DIM>(S.x, Npad, k, xi, L);
        const double *po = par + (si * ns + sk_old) * PMC_NPAR;
        const double *pn = par + (sj * ns + sk_new) * PMC_NPAR;
        if (bonded_to<MOL>(bi, k)) {
            t += bond_potential(pn, r2) - bond_potential(po, r2);
        } else {
            if (r2 <= po[PMC_P_RCUT2]) t -= pair_potential<MODEL>(po, r2);
            if (r2 <= pn[PMC_P_RCUT2]) t += pair_potential<MODEL>(pn, r2);
        }
    }
    if (k != j) {  // term of particle j's local energy
        const double r2 = dist2<DIM>(S.x, Npad, k, xj, L);
        const double *po = par + (sj * ns + sk_old) * PMC_NPAR;
        const double *pn = par + (si * ns + sk_new) * PMC_NPAR;
        if (bonded_to<MOL>(bj, k)) {
            t += bond_potential(pn, r2) - bond_potential(po, r2);
        } else {
            if (r2 <= po[PMC_P_RCUT2]) t -= pair_potential<MODEL>(po, r2);
            if (r2 <= pn[PMC_P_RCUT2]) t += pair_potential<MODEL>(pn, r2);
        }
    }
    return t;
}

// ------------------------------------------------------------------------------------------------
// Sweep kernel.  FILTER = true (cubic boxes): every candidate first goes through an integer
// fixed-point distance test on the ALU / integer-FMA pipes; only survivors (a conservative superset of
// the pairs inside the cutoff) are compacted per warp and evaluated in fp64.  The result is the same set
// of fp64 pair terms as FILTER = false, which visits all candidates in fp64.
// ------------------------------------------------------------------------------------------------
template <int DIM, int MODEL, bool MOL, bool FILTER, int NTMAX>
__global__ void __launch_bounds__(NTMAX, NTMAX <= 128 ? 5 : 2) k_chain_sweep(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SweepStatic T;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    const int c = blockIdx.x;
    const int N = A.N, Npad = A.Npad, ns = A.ns;
    SweepDyn S;
    carve_sweep(S, smem_raw, DIM, Npad, NT, MOL, A.any_swap != 0, FILTER);

    // ---- load chain state into shared memory --------------------------------------------------
    const double L[3] = {A.box[c * 3 + 0], A.box[c * 3 + 1], A.box[c * 3 + 2]};
    const double fscale = 4294967296.0 / L[0];  // FILTER: fixed-point units per unit length
    double *gx = A.x + (size_t)c * DIM * Npad;
    for (int k = tid; k < DIM * Npad; k += NT) {
        const double v = gx[k];
        S.x[k] = v;
        if constexpr (FILTER) S.u[k] = to_fixed(v, fscale);
    }
    uint8_t *gsp = A.sp + (size_t)c * Npad;
    for (int k = tid; k < Npad; k += NT) S.sp[k] = gsp[k];
    for (int k = tid; k < ns * ns * PMC_NPAR; k += NT) T.par[k] = A.par[k];
    if (A.any_swap) {
        const uint16_t *gi = A.spids + (size_t)c * Npad, *gh = A.heads + (size_t)c * Npad;
        for (int k = tid; k < Npad; k += NT) {
            S.spids[k] = gi[k];
            S.heads[k] = gh[k];
        }
    }
    if (tid <= PMC_MAX_SPECIES) T.spoff[tid] = A.spoff[c * (PMC_MAX_SPECIES + 1) + tid];
    if (tid < 2 * PMC_MAX_MOVES) (&T.cnt[0][0])[tid] = 0ull;
    if constexpr (MOL) {
        for (int k = tid; k < Npad * PMC_MAX_BONDS; k += NT) S.bonds[k] = A.bonds[k];
    }
    // FILTER: the first kRegCand * NT candidates keep their fixed-point coordinates in REGISTERS of the
    // thread that owns them (candidate j = k * NT + tid lives in slot k of thread tid), so the scan has no
    // loads or address arithmetic; candidates beyond that (large N) are scanned from shared memory.
    uint32_t myu[kRegCand][DIM];
    if constexpr (FILTER) {
#pragma unroll
        for (int k = 0; k < kRegCand; k++) {
            const int j = k * NT + tid;
#pragma unroll
            for (int a = 0; a < DIM; a++) myu[k][a] = j < Npad ? to_fixed(gx[a * Npad + j], fscale) : 0u;
        }
        if (tid <= PMC_MAX_SPECIES) {  // conservative cutoffs per species of the moved particle; [4] = global
            double rc2 = 0.0;
            for (int a = 0; a < ns; a++)
                for (int b = 0; b < ns; b++)
                    if (tid == PMC_MAX_SPECIES || a == tid) rc2 = fmax(rc2, A.par[(a * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            T.thr_u[tid] = fixed_threshold(rc2 * (fscale / L[0]));
            if (tid < PMC_MAX_SPECIES) T.rcs[tid] = sqrt(rc2);
        }
    }
    const double Tk = A.temp[c];
    double E = A.energy[c];
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t gchain = (uint32_t)(A.chain_offset + c);
    int32_t *gimg = A.img + (size_t)c * DIM * Npad;
    uint16_t *myq = FILTER ? S.wq + warp * queue_cap(Npad, NT) : nullptr;
    const unsigned lt_mask = (1u << lane) - 1u;

    int last_i = -1;  // most recently committed particle and its position (forwarded, see below)
    double last_x[3] = {0.0, 0.0, 0.0};
    uint32_t last_u[3] = {0u, 0u, 0u};
    int slot = 0;

    for (long long tb = 0; tb < A.n_trials; tb += kBatch) {
        const int nb = (int)min((long long)kBatch, A.n_trials - tb);
        __syncthreads();  // previous batch fully consumed; first pass: state loaded
        // ---- proposals of trials tb .. tb+nb-1, one per thread --------------------------------
        for (int t_ = tid; t_ < nb; t_ += NT) {
            const long long q = tb + t_;
            pmc_trial tr;
            unsigned long long flip_pairs = 0ull;
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
                if (tr.kind == PMC_MOVE_DISPLACEMENT) {
                    float z0, z1, z2, z3;
                    box_muller(b.v[0], b.v[1], z0, z1);
                    box_muller(b.v[2], b.v[3], z2, z3);
                    const float sg = A.mv_sigma[m];
                    tr.i = (int)bounded(a.v[1], (uint32_t)N);
                    tr.j = -1;
                    tr.delta[0] = (double)(sg * z0);
                    tr.delta[1] = (double)(sg * z1);
                    tr.delta[2] = (DIM == 3) ? (double)(sg * z2) : 0.0;
                } else if (tr.kind == PMC_MOVE_FLIP) {
                    // MoleculeFlip (src/moves.jl:344-352): a molecule uniformly, then ordered pairs of distinct sites
                    // until their species differ.  Species are only known when the trial executes, so four candidate
                    // pairs are parked; the first unlike one is taken (none: the trial is a rejected no-op).
                    const int mol = A.n_mol > 0 ? (int)bounded(a.v[1], (uint32_t)A.n_mol) : 0;
                    const int len = A.n_mol > 0 ? A.mol_len[mol] : 0;
                    tr.i = A.n_mol > 0 ? A.mol_start[mol] : -1;
                    tr.j = -1;
                    const Philox4 c4 = philox4x32_10((uint32_t)t, (uint32_t)(t >> 32), gchain, 2u, k0, k1);
                    const uint32_t w[8] = {b.v[0], b.v[1], b.v[2], b.v[3], c4.v[0], c4.v[1], c4.v[2], c4.v[3]};
                    unsigned long long packed = 0;
                    for (int q = 0; q < 4; q++) {
                        uint32_t pa = 0, pb = 0;
                        if (len >= 2) {
                            pa = bounded(w[2 * q], (uint32_t)len);
                            pb = bounded(w[2 * q + 1], (uint32_t)(len - 1));
                            pb += (pb >= pa) ? 1u : 0u;
                        }
                        packed |= (unsigned long long)((pa & 0xFFu) | ((pb & 0xFFu) << 8)) << (16 * q);
                    }
                    flip_pairs = len >= 2 && len <= 255 ? packed : 0ull;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                } else {  // slots in the species lists; resolved to particles when the trial executes
                    const int nA = T.spoff[A.mv_a[m] + 1] - T.spoff[A.mv_a[m]];
                    const int nB = T.spoff[A.mv_b[m] + 1] - T.spoff[A.mv_b[m]];
                    tr.i = (nA > 0 && nB > 0) ? (int)bounded(a.v[1], (uint32_t)nA) : -1;
                    tr.j = (nA > 0 && nB > 0) ? (int)bounded(b.v[0], (uint32_t)nB) : -1;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                }
                if (A.trace) A.trace[(size_t)c * A.n_trials + q] = tr;
            }
            TrialRec &R = T.trial[t_];
            R.thr = A.exact_exp ? tr.u : -Tk * log(tr.u);
            R.m = tr.move;
            R.i = tr.i;
            R.j = (tr.kind == PMC_MOVE_DISPLACEMENT) ? -2 : tr.j;  // -2 marks a displacement
            R.pad[0] = (tr.kind == PMC_MOVE_FLIP && !A.replay) ? (int)(uint32_t)flip_pairs : 0;
            R.pad[1] = (tr.kind == PMC_MOVE_FLIP && !A.replay) ? (int)(uint32_t)(flip_pairs >> 32) : 0;
            R.flip = (tr.kind == PMC_MOVE_FLIP && !A.replay) ? 1 : 0;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                R.delta[a] = tr.delta[a];
                if constexpr (FILTER) R.dint[a] = (int)__double2ll_rn(tr.delta[a] * fscale);
            }
            if constexpr (FILTER) {
                // one sphere around the midpoint of old and new position covers both cutoff spheres
                const double hd = 0.5 * sqrt(tr.delta[0] * tr.delta[0] + tr.delta[1] * tr.delta[1] + tr.delta[2] * tr.delta[2]);
                for (int sp_ = 0; sp_ < PMC_MAX_SPECIES; sp_++) {
                    const double r = T.rcs[sp_ < ns ? sp_ : 0] + hd;
                    R.thr_t[sp_] = fixed_threshold(r * r * (fscale / L[0]));
                }
            }
        }
        __syncthreads();

        // ---- the serial chain: one trial at a time ---------------------------------------------
        for (int b = 0; b < nb; b++) {
            const TrialRec &R = T.trial[b];
            const int m = R.m;
            const bool is_disp = R.j == -2;
            double part = 0.0, dE;
            bool acc;
            if (is_disp) {
                const int i = R.i;
                double xo[3] = {0.0, 0.0, 0.0}, xn[3] = {0.0, 0.0, 0.0};
                int w[3] = {0, 0, 0};
                // The owner thread (i % NT) commits accepted positions after the barrier of the
                // previous trial; every other thread only ever reads x[i] here, at the start of a
                // trial.  A commit becomes visible at the NEXT barrier, so the only unordered read
                // is "same particle as the most recent commit" -- served from registers instead.
                const bool fwd = (i == last_i);
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    xo[a] = fwd ? last_x[a] : S.x[a * Npad + i];
                    xn[a] = wrap1(xo[a] + R.delta[a], L[a], w[a]);
                }
                const int si = S.sp[i];
                uint16_t bi[PMC_MAX_BONDS];
                if constexpr (MOL) {
#pragma unroll
                    for (int k = 0; k < PMC_MAX_BONDS; k++) bi[k] = S.bonds[i * PMC_MAX_BONDS + k];
                }
                const double *prow = T.par + si * ns * PMC_NPAR;
                uint32_t un[3] = {0u, 0u, 0u};
                if constexpr (FILTER) {
                    uint32_t um[3] = {0u, 0u, 0u};
#pragma unroll
                    for (int a = 0; a < DIM; a++) {
                        const uint32_t uo = fwd ? last_u[a] : S.u[a * Npad + i];
                        un[a] = uo + (uint32_t)R.dint[a];        // wraps like the box does
                        um[a] = uo + (uint32_t)(R.dint[a] >> 1);  // midpoint of old and new
                    }
                    const uint32_t thr = R.thr_t[si];
                    // Branch-free scan of all candidates against ONE sphere (midpoint, rc + |delta|/2).  Self
                    // and padding entries are not excluded here (the fp64 pass drops them).  Per candidate:
                    // 3 wrapping subtractions (= minimum image), 3 mul-hi accumulate, compare, vote, compact.
                    int cnt = 0;
#pragma unroll
                    for (int k = 0; k < kRegCand; k++) {
                        if (k * NT < Npad) {  // uniform
                            bool pass = dist2_fixed<DIM>(um[0], um[1], um[2], myu[k][0], myu[k][1], DIM == 3 ? myu[k][DIM - 1] : 0u) <= thr;
                            if constexpr (MOL) pass |= bonded_to<MOL>(bi, k * NT + tid);
                            const unsigned mask = __ballot_sync(0xffffffffu, pass);
                            if (pass) myq[cnt + __popc(mask & lt_mask)] = (uint16_t)(k * NT + tid);
                            cnt += __popc(mask);
                        }
                    }
                    for (int j0 = kRegCand * NT + tid; j0 < Npad + tid; j0 += NT) {  // large N: the rest from smem
                        const int jc = min(j0, Npad - 1);
                        bool pass = dist2_fixed<DIM>(um[0], um[1], um[2], S.u[jc], S.u[Npad + jc], DIM == 3 ? S.u[(DIM - 1) * Npad + jc] : 0u) <= thr;
                        if constexpr (MOL) pass |= bonded_to<MOL>(bi, j0);
                        const unsigned mask = __ballot_sync(0xffffffffu, pass);
                        if (pass) myq[cnt + __popc(mask & lt_mask)] = (uint16_t)j0;
                        cnt += __popc(mask);
                    }
                    __syncwarp();
                    for (int q = lane; q < cnt; q += 32) {
                        const int j = myq[q];
                        if (j < N && j != i) part += displacement_term<DIM, MODEL, MOL>(S, Npad, j, xo, xn, L, prow, bi);
                    }
                    __syncwarp();
                } else {
                    for (int j = tid; j < N; j += NT)
                        if (j != i) part += displacement_term<DIM, MODEL, MOL>(S, Npad, j, xo, xn, L, prow, bi);
                }
                // block-wide sum, one barrier; every thread ends with the same bits.  `slot` alternates so a
                // warp racing ahead into the next trial cannot overwrite partials still being read.
                part = warp_sum(part);
                if (lane == 0) T.red[slot][warp] = part;
                __syncthreads();
                dE = T.red[slot][0];
                for (int w_ = 1; w_ < nwarp; w_++) dE += T.red[slot][w_];
                slot ^= 1;
                acc = A.exact_exp ? accept_exact(dE, Tk, R.thr) : (dE < R.thr);
                if (acc) {
                    if (tid == i % NT) {
#pragma unroll
                        for (int a = 0; a < DIM; a++) {
                            S.x[a * Npad + i] = xn[a];
                            if (w[a] != 0) atomicAdd(&gimg[a * Npad + i], w[a]);
                            if constexpr (FILTER) {
                                const uint32_t v = to_fixed(xn[a], fscale);
                                S.u[a * Npad + i] = v;
                                const int ki = i / NT;
#pragma unroll
                                for (int k = 0; k < kRegCand; k++)
                                    if (k == ki) myu[k][a] = v;
                            }
                        }
                    }
                    last_i = i;
#pragma unroll
                    for (int a = 0; a < DIM; a++) {
                        last_x[a] = xn[a];
                        last_u[a] = un[a];
                    }
                    E += dE;
                }
            } else {
                // DiscreteSwap: i from the species-A list, j from the species-B list; positions fixed,
                // four local energies folded into one pass (src/moves.jl:159-167).
                const int ka = R.i, kb = R.j;
                int i = -1, j = -1;
                if (A.replay) {
                    i = ka;
                    j = kb;
                } else if (R.flip) {
                    const unsigned long long packed = (unsigned long long)(uint32_t)R.pad[0] | ((unsigned long long)(uint32_t)R.pad[1] << 32);
                    if (packed != 0ull && ka >= 0) {
                        for (int q = 3; q >= 0; q--) {  // first unlike pair wins
                            const int sa = ka + (int)((packed >> (16 * q)) & 0xFFull), sb = ka + (int)((packed >> (16 * q + 8)) & 0xFFull);
                            if (S.sp[sa] != S.sp[sb]) {
                                i = sa;
                                j = sb;
                            }
                        }
                    }
                } else if (ka >= 0) {
                    i = S.spids[T.spoff[A.mv_a[m]] + ka];
                    j = S.spids[T.spoff[A.mv_b[m]] + kb];
                }
                if (i >= 0 && j >= 0) {
                    double xi[3] = {0.0, 0.0, 0.0}, xj[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int a = 0; a < DIM; a++) {
                        xi[a] = (i == last_i) ? last_x[a] : S.x[a * Npad + i];
                        xj[a] = (j == last_i) ? last_x[a] : S.x[a * Npad + j];
                    }
                    const int si = S.sp[i], sj = S.sp[j];
                    uint16_t bi[PMC_MAX_BONDS], bj[PMC_MAX_BONDS];
                    if constexpr (MOL) {
#pragma unroll
                        for (int k = 0; k < PMC_MAX_BONDS; k++) {
                            bi[k] = S.bonds[i * PMC_MAX_BONDS + k];
                            bj[k] = S.bonds[j * PMC_MAX_BONDS + k];
                        }
                    }
                    if constexpr (FILTER) {
                        uint32_t ui[3] = {0u, 0u, 0u}, uj[3] = {0u, 0u, 0u};
#pragma unroll
                        for (int a = 0; a < DIM; a++) {
                            ui[a] = (i == last_i) ? last_u[a] : S.u[a * Npad + i];
                            uj[a] = (j == last_i) ? last_u[a] : S.u[a * Npad + j];
                        }
                        const uint32_t thr = T.thr_u[PMC_MAX_SPECIES];
                        int cnt = 0;
                        for (int k_ = tid; k_ < Npad + tid; k_ += NT) {
                            const int kc = min(k_, Npad - 1);
                            const uint32_t a0 = S.u[kc], a1 = S.u[Npad + kc], a2 = DIM == 3 ? S.u[(DIM - 1) * Npad + kc] : 0u;
                            bool pass = min(dist2_fixed<DIM>(ui[0], ui[1], ui[2], a0, a1, a2),
                                            dist2_fixed<DIM>(uj[0], uj[1], uj[2], a0, a1, a2)) <= thr;
                            if constexpr (MOL) pass |= bonded_to<MOL>(bi, k_) || bonded_to<MOL>(bj, k_);
                            const unsigned mask = __ballot_sync(0xffffffffu, pass);
                            if (pass) myq[cnt + __popc(mask & lt_mask)] = (uint16_t)k_;
                            cnt += __popc(mask);
                        }
                        __syncwarp();
                        for (int q = lane; q < cnt; q += 32) {
                            const int k = myq[q];
                            if (k < N) part += swap_term<DIM, MODEL, MOL>(S, T.par, Npad, ns, k, i, j, si, sj, xi, xj, L, bi, bj);
                        }
                        __syncwarp();
                    } else {
                        for (int k = tid; k < N; k += NT)
                            part += swap_term<DIM, MODEL, MOL>(S, T.par, Npad, ns, k, i, j, si, sj, xi, xj, L, bi, bj);
                    }
                }
                part = warp_sum(part);
                if (lane == 0) T.red[slot][warp] = part;
                __syncthreads();
                dE = T.red[slot][0];
                for (int w_ = 1; w_ < nwarp; w_++) dE += T.red[slot][w_];
                slot ^= 1;
                acc = (i >= 0 && j >= 0) && (A.exact_exp ? accept_exact(dE, Tk, R.thr) : (dE < R.thr));
                if (acc) {
                    if (tid == 0) {
                        const uint8_t si = S.sp[i], sj = S.sp[j];
                        S.sp[i] = sj;
                        S.sp[j] = si;
                        if (A.any_swap) {  // update_species_list! (src/moves.jl:175-179)
                            const uint16_t hi = S.heads[i], hj = S.heads[j];
                            S.spids[T.spoff[si] + hi] = (uint16_t)j;
                            S.spids[T.spoff[sj] + hj] = (uint16_t)i;
                            S.heads[i] = hj;
                            S.heads[j] = hi;
                        }
                    }
                    E += dE;
                    __syncthreads();  // species and species lists are read by every thread
                }
                if (A.trace && tid == 0) {
                    pmc_trial *tr = A.trace + (size_t)c * A.n_trials + tb + b;
                    tr->i = i;
                    tr->j = j;
                }
            }
            if (tid == 0) {
                T.cnt[0][m] += 1ull;
                T.cnt[1][m] += acc ? 1ull : 0ull;
                if (A.acc_out) A.acc_out[(size_t)c * A.n_trials + tb + b] = acc ? 1 : 0;
                if (A.dE_out) A.dE_out[(size_t)c * A.n_trials + tb + b] = dE;
            }
        }
    }
    __syncthreads();

    // ---- write the chain state back -------------------------------------------------------------
    for (int k = tid; k < DIM * Npad; k += NT) gx[k] = S.x[k];
    if (A.any_swap) {
        for (int k = tid; k < Npad; k += NT) gsp[k] = S.sp[k];
        uint16_t *gi = A.spids + (size_t)c * Npad, *gh = A.heads + (size_t)c * Npad;
        for (int k = tid; k < Npad; k += NT) {
            gi[k] = S.spids[k];
            gh[k] = S.heads[k];
        }
    }
    if (tid == 0) A.energy[c] = E;
    if (tid < A.n_moves) {
        A.calls[(size_t)c * PMC_MAX_MOVES + tid] += T.cnt[0][tid];
        A.accepted[(size_t)c * PMC_MAX_MOVES + tid] += T.cnt[1][tid];
    }
}

// ------------------------------------------------------------------------------------------------
// Energy kernel: local energies of every particle of every chain + total = sum/2
// (src/atoms.jl:51-52, :81-88; src/molecules.jl:89-90, :206-215)
// ------------------------------------------------------------------------------------------------
template <int DIM, int MODEL, bool MOL>
__global__ void k_chain_energy(const __grid_constant__ EnergyArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    const int c = blockIdx.x, N = A.N, Npad = A.Npad, ns = A.ns;
    double *sx = (double *)smem_raw;
    double *se = sx + DIM * Npad;
    double *spar = se + Npad;
    uint8_t *ssp = (uint8_t *)(spar + ns * ns * PMC_NPAR);
    const double *gx = A.x + (size_t)c * DIM * Npad;
    for (int k = tid; k < DIM * Npad; k += NT) sx[k] = gx[k];
    for (int k = tid; k < Npad; k += NT) ssp[k] = A.sp[(size_t)c * Npad + k];
    for (int k = tid; k < ns * ns * PMC_NPAR; k += NT) spar[k] = A.par[k];
    const double L[3] = {A.box[c * 3 + 0], A.box[c * 3 + 1], A.box[c * 3 + 2]};
    __syncthreads();
    for (int i = warp; i < N; i += nwarp) {
        double xi[3] = {sx[i], sx[Npad + i], DIM == 3 ? sx[2 * Npad + i] : 0.0};
        const double *prow = spar + ssp[i] * ns * PMC_NPAR;
        uint16_t bi[PMC_MAX_BONDS];
        if constexpr (MOL) {
#pragma unroll
            for (int k = 0; k < PMC_MAX_BONDS; k++) bi[k] = A.bonds[i * PMC_MAX_BONDS + k];
        }
        double e = 0.0;
        for (int j = lane; j < N; j += 32) {
            if (j == i) continue;
            const double *p = prow + ssp[j] * PMC_NPAR;
            const double r2 = dist2<DIM>(sx, Npad, j, xi, L);
            if (bonded_to<MOL>(bi, j)) {
                e += bond_potential(p, r2);
            } else if (r2 <= p[PMC_P_RCUT2]) {
                e += pair_potential<MODEL>(p, r2);
            }
        }
        e = warp_sum(e);
        if (lane == 0) {
            se[i] = e;
            A.eloc[(size_t)c * Npad + i] = e;
        }
    }
    __syncthreads();
    // deterministic block reduction of se[0..N)
    double s = 0.0;
    for (int k = tid; k < N; k += NT) s += se[k];
    __syncthreads();
    s = warp_sum(s);
    if (lane == 0) se[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < nwarp; w++) tot += se[w];
        A.etot[c] = tot / 2;
    }
}

// ------------------------------------------------------------------------------------------------
// Pair-distance histogram (raw counts of g(r)): one CTA per chain, positions in shared memory, all i < j,
// per-CTA histogram in shared memory flushed with one atomicAdd per bin.  Integer counts: deterministic.
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_chain_pair_histogram(const double *__restrict__ x, const uint8_t *__restrict__ sp,
                                       const double *__restrict__ box, int N, int Npad, int sa, int sb, double rmax,
                                       int nbins, unsigned long long *hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sx = (double *)smem_raw;
    unsigned int *sh = (unsigned int *)(sx + DIM * Npad);
    uint8_t *ssp = (uint8_t *)(sh + nbins);
    const int tid = threadIdx.x, NT = blockDim.x, c = blockIdx.x;
    for (int k = tid; k < DIM * Npad; k += NT) sx[k] = x[(size_t)c * DIM * Npad + k];
    for (int k = tid; k < Npad; k += NT) ssp[k] = sp[(size_t)c * Npad + k];
    for (int k = tid; k < nbins; k += NT) sh[k] = 0u;
    const double L[3] = {box[c * 3 + 0], box[c * 3 + 1], box[c * 3 + 2]};
    const double inv_dr = (double)nbins / rmax, rmax2 = rmax * rmax;
    __syncthreads();
    for (int i = 0; i < N - 1; i++) {
        const double xi[3] = {sx[i], sx[Npad + i], DIM == 3 ? sx[2 * Npad + i] : 0.0};
        const int si = ssp[i];
        for (int j = i + 1 + tid; j < N; j += NT) {
            const int sj = ssp[j];
            const bool match = (sa < 0 && sb < 0) || (sa < 0 && (si == sb || sj == sb)) || (sb < 0 && (si == sa || sj == sa)) ||
                               (si == sa && sj == sb) || (si == sb && sj == sa);
            if (!match) continue;
            const double r2 = dist2<DIM>(sx, Npad, j, xi, L);
            if (r2 < rmax2) {
                int bin = (int)(sqrt(r2) * inv_dr);
                bin = bin < nbins ? bin : nbins - 1;
                atomicAdd(&sh[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < nbins; k += NT)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

template <typename F>
cudaError_t dispatch(int dim, int model, bool mol, F &&f) {
#define PMC_CASE(D, MDL, ML) \
    if (dim == D && model == MDL && mol == ML) return f(std::integral_constant<int, D>{}, std::integral_constant<int, MDL>{}, std::integral_constant<bool, ML>{});
    PMC_CASE(3, PMC_MODEL_LJ, false)
    PMC_CASE(2, PMC_MODEL_LJ, false)
    PMC_CASE(3, PMC_MODEL_SOFT, false)
    PMC_CASE(2, PMC_MODEL_SOFT, false)
    PMC_CASE(3, PMC_MODEL_SMOOTHLJ, false)
    PMC_CASE(2, PMC_MODEL_SMOOTHLJ, false)
    PMC_CASE(3, PMC_MODEL_KG, false)
    PMC_CASE(2, PMC_MODEL_KG, false)
    PMC_CASE(3, PMC_MODEL_KG, true)
    PMC_CASE(2, PMC_MODEL_KG, true)
#undef PMC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace

size_t chain_sweep_smem_bytes(int dim, int Npad, int, int threads, bool mol, bool any_swap, bool filter) {
    SweepDyn s;
    return carve_sweep(s, nullptr, dim, Npad, threads, mol, any_swap, filter);
}

size_t chain_energy_smem_bytes(int dim, int Npad, int ns, bool) {
    return sizeof(double) * ((size_t)dim * Npad + Npad + (size_t)ns * ns * PMC_NPAR) + Npad + 16;
}

cudaError_t configure_chain_kernels(int dim, int model, bool mol, size_t sweep_smem, size_t sweep_smem_filter,
                                    size_t energy_smem) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        constexpr bool ml = decltype(ML)::value;
        const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        cudaError_t e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, false, 128>, attr, (int)sweep_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, false, 256>, attr, (int)sweep_smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, true, 128>, attr, (int)sweep_smem_filter);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, true, 256>, attr, (int)sweep_smem_filter);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_chain_energy<d, mdl, ml>, attr, (int)energy_smem);
        return e;
    });
}

// Two register budgets per kernel: CTAs of <= 128 threads are compiled for 5 resident CTAs per SM (the
// shared-memory residency at N = 1000), larger CTAs for 2.
cudaError_t launch_chain_sweep(int dim, int model, bool mol, bool filter, int M, int threads, size_t smem,
                               const ChainArgs &a, cudaStream_t st) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        constexpr bool ml = decltype(ML)::value;
        if (threads <= 128) {
            if (filter)
                k_chain_sweep<d, mdl, ml, true, 128><<<M, threads, smem, st>>>(a);
            else
                k_chain_sweep<d, mdl, ml, false, 128><<<M, threads, smem, st>>>(a);
        } else {
            if (filter)
                k_chain_sweep<d, mdl, ml, true, 256><<<M, threads, smem, st>>>(a);
            else
                k_chain_sweep<d, mdl, ml, false, 256><<<M, threads, smem, st>>>(a);
        }
        return cudaGetLastError();
    });
}

cudaError_t launch_chain_pair_histogram(int dim, int M, int N, int Npad, const double *x, const uint8_t *sp,
                                        const double *box, int sa, int sb, double rmax, int nbins,
                                        unsigned long long *hist, cudaStream_t st) {
    const size_t smem = sizeof(double) * (size_t)dim * Npad + sizeof(unsigned int) * nbins + Npad + 16;
    cudaError_t e;
    if (dim == 3) {
        e = cudaFuncSetAttribute(k_chain_pair_histogram<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_chain_pair_histogram<3><<<M, 256, smem, st>>>(x, sp, box, N, Npad, sa, sb, rmax, nbins, hist);
    } else {
        e = cudaFuncSetAttribute(k_chain_pair_histogram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_chain_pair_histogram<2><<<M, 256, smem, st>>>(x, sp, box, N, Npad, sa, sb, rmax, nbins, hist);
    }
    return cudaGetLastError();
}

cudaError_t launch_chain_energy(int dim, int model, bool mol, int M, size_t smem, const EnergyArgs &a,
                                cudaStream_t st) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        k_chain_energy<decltype(D)::value, decltype(MDL)::value, decltype(ML)::value><<<M, 256, smem, st>>>(a);
        return cudaGetLastError();
    });
}

}  // namespace pmc
