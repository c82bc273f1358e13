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
constexpr int kMinBlocks = 4;     // resident CTAs per SM the register allocation is sized for
constexpr int kMaxWarps = kMaxThreads / 32;
constexpr int kBatch = 64;  // proposals generated per batch (parked in shared memory)

struct SweepSmem {
    double *x;        // [DIM][Npad] wrapped positions (source of truth)
    double *par;      // [ns*ns*PMC_NPAR]
    double *red;      // [2][kMaxWarps]
    double *delta;    // [kBatch][3]
    double *thr;      // [kBatch]  -T*log(u)  (or u itself in exact_exp mode)
    unsigned long long *cnt;  // [2][PMC_MAX_MOVES] calls, accepted
    uint32_t *u;      // [DIM][Npad] FILTER: positions as 32-bit fixed-point fractions of L
    int *dint;        // [kBatch][3] FILTER: delta in the same fixed-point units
    uint32_t *thr_u;  // [PMC_MAX_SPECIES+1] FILTER: cutoff^2 in fixed-point units per species of i; last = global
    int *ti;          // [kBatch]  particle i, or slot ka for swaps
    int *tj;          // [kBatch]  slot kb for swaps
    int *tm;          // [kBatch]  pool index
    int *spoff;       // [PMC_MAX_SPECIES+1]
    uint16_t *wq;     // [nwarp][qcap] FILTER: per-warp queues of surviving candidates
    uint16_t *spids;  // [Npad]
    uint16_t *heads;  // [Npad]
    uint16_t *bonds;  // [Npad][PMC_MAX_BONDS] (MOL only)
    uint8_t *sp;      // [Npad]
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ inline int queue_cap(int Npad, int NT) { return (Npad + NT - 1) / NT * 32; }

__host__ __device__ inline size_t carve_sweep(SweepSmem &s, unsigned char *base, int dim, int Npad, int ns, int NT,
                                              bool mol, bool any_swap, bool filter) {
    size_t o = 0;
    auto take = [&](size_t bytes) {
        unsigned char *p = base + o;
        o += align_up(bytes, 8);
        return p;
    };
    s.x = (double *)take(sizeof(double) * dim * Npad);
    s.par = (double *)take(sizeof(double) * ns * ns * PMC_NPAR);
    s.red = (double *)take(sizeof(double) * 2 * kMaxWarps);
    s.delta = (double *)take(sizeof(double) * 3 * kBatch);
    s.thr = (double *)take(sizeof(double) * kBatch);
    s.cnt = (unsigned long long *)take(sizeof(unsigned long long) * 2 * PMC_MAX_MOVES);
    s.u = (uint32_t *)take(filter ? sizeof(uint32_t) * dim * Npad : 0);
    s.dint = (int *)take(filter ? sizeof(int) * 3 * kBatch : 0);
    s.thr_u = (uint32_t *)take(sizeof(uint32_t) * 8);
    s.ti = (int *)take(sizeof(int) * kBatch);
    s.tj = (int *)take(sizeof(int) * kBatch);
    s.tm = (int *)take(sizeof(int) * kBatch);
    s.spoff = (int *)take(sizeof(int) * 8);
    s.wq = (uint16_t *)take(filter ? sizeof(uint16_t) * (NT / 32) * queue_cap(Npad, NT) : 0);
    s.spids = (uint16_t *)take(any_swap ? sizeof(uint16_t) * Npad : 0);
    s.heads = (uint16_t *)take(any_swap ? sizeof(uint16_t) * Npad : 0);
    s.bonds = (uint16_t *)take(mol ? sizeof(uint16_t) * Npad * PMC_MAX_BONDS : 0);
    s.sp = (uint8_t *)take(Npad);
    return align_up(o, 16);
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
__device__ __forceinline__ uint32_t dist2_fixed(const uint32_t *__restrict__ su, int Npad, int j, const uint32_t (&ui)[3]) {
    const int dx = (int)(ui[0] - su[j]), dy = (int)(ui[1] - su[Npad + j]);
    uint32_t r = (uint32_t)__mulhi(dx, dx) + (uint32_t)__mulhi(dy, dy);
    if constexpr (DIM == 3) {
        const int dz = (int)(ui[2] - su[2 * Npad + j]);
        r += (uint32_t)__mulhi(dz, dz);
    }
    return r;
}

__device__ __forceinline__ uint32_t to_fixed(double x, double scale) {  // scale = 2^32 / L
    return (uint32_t)__double2ull_rd(x * scale);
}

// Block-wide sum with one barrier; every thread returns the same bits.  `slot` alternates per trial
// so that a warp racing ahead into the next trial cannot overwrite partials still being read.
__device__ __forceinline__ double block_sum(double v, double *red, int slot, int lane, int warp, int nwarp) {
    v = warp_sum(v);
    if (lane == 0) red[slot * kMaxWarps + warp] = v;
    __syncthreads();
    double s = red[slot * kMaxWarps];
    for (int w = 1; w < nwarp; w++) s += red[slot * kMaxWarps + w];
    return s;
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
__device__ __forceinline__ double displacement_term(const SweepSmem &S, int Npad, int j, const double (&xo)[3],
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
__device__ __forceinline__ double swap_term(const SweepSmem &S, int Npad, int ns, int k, int i, int j, int si, int sj,
                                            const double (&xi)[3], const double (&xj)[3], const double (&L)[3],
                                            const uint16_t (&bi)[PMC_MAX_BONDS], const uint16_t (&bj)[PMC_MAX_BONDS]) {
    const int sk_old = S.sp[k];
    int sk_new = sk_old;
    if (k == i) sk_new = sj;
    if (k == j) sk_new = si;
    double t = 0.0;
    if (k != i) {  // term of particle i's local energy
        const double r2 = dist2<DIM>(S.x, Npad, k, xi, L);
        const double *po = S.par + (si * ns + sk_old) * PMC_NPAR;
        const double *pn = S.par + (sj * ns + sk_new) * PMC_NPAR;
        if (bonded_to<MOL>(bi, k)) {
            t += bond_potential(pn, r2) - bond_potential(po, r2);
        } else {
            if (r2 <= po[PMC_P_RCUT2]) t -= pair_potential<MODEL>(po, r2);
            if (r2 <= pn[PMC_P_RCUT2]) t += pair_potential<MODEL>(pn, r2);
        }
    }
    if (k != j) {  // term of particle j's local energy
        const double r2 = dist2<DIM>(S.x, Npad, k, xj, L);
        const double *po = S.par + (sj * ns + sk_old) * PMC_NPAR;
        const double *pn = S.par + (si * ns + sk_new) * PMC_NPAR;
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
// fixed-point distance test on the ALU/FMA-int pipes; only survivors (a conservative superset of the
// pairs inside the cutoff) are compacted per warp and evaluated in fp64.  The result is the same set of
// fp64 pair terms as FILTER = false, which visits all candidates in fp64.
// ------------------------------------------------------------------------------------------------
template <int DIM, int MODEL, bool MOL, bool FILTER>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) k_chain_sweep(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    const int c = blockIdx.x;
    const int N = A.N, Npad = A.Npad, ns = A.ns;
    SweepSmem S;
    carve_sweep(S, smem_raw, DIM, Npad, ns, NT, MOL, A.any_swap != 0, FILTER);

    // ---- load chain state into shared memory --------------------------------------------------
    double L[3] = {A.box[c * 3 + 0], A.box[c * 3 + 1], A.box[c * 3 + 2]};
    const double fscale = 4294967296.0 / L[0];  // FILTER: fixed-point units per unit length
    double *gx = A.x + (size_t)c * DIM * Npad;
    for (int k = tid; k < DIM * Npad; k += NT) {
        const double v = gx[k];
        S.x[k] = v;
        if constexpr (FILTER) S.u[k] = to_fixed(v, fscale);
    }
    uint8_t *gsp = A.sp + (size_t)c * Npad;
    for (int k = tid; k < Npad; k += NT) S.sp[k] = gsp[k];
    for (int k = tid; k < ns * ns * PMC_NPAR; k += NT) S.par[k] = A.par[k];
    if (A.any_swap) {
        const uint16_t *gi = A.spids + (size_t)c * Npad, *gh = A.heads + (size_t)c * Npad;
        for (int k = tid; k < Npad; k += NT) {
            S.spids[k] = gi[k];
            S.heads[k] = gh[k];
        }
    }
    if (tid <= PMC_MAX_SPECIES) S.spoff[tid] = A.spoff[c * (PMC_MAX_SPECIES + 1) + tid];
    if (tid < 2 * PMC_MAX_MOVES) S.cnt[tid] = 0ull;
    if constexpr (MOL) {
        for (int k = tid; k < Npad * PMC_MAX_BONDS; k += NT) S.bonds[k] = A.bonds[k];
    }
    if constexpr (FILTER) {
        if (tid <= PMC_MAX_SPECIES) {  // conservative integer cutoffs: per species of i, and global (swaps)
            double rc2 = 0.0;
            for (int a = 0; a < ns; a++)
                for (int b = 0; b < ns; b++)
                    if (tid == PMC_MAX_SPECIES || a == tid) rc2 = fmax(rc2, A.par[(a * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            const double t = rc2 * (1.0 + 1e-9) * (fscale / L[0]) + 64.0;  // r^2 * 2^32 / L^2, + rounding slack
            S.thr_u[tid] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        }
    }
    const double T = A.temp[c];
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
            if (A.replay) {
                tr = A.replay[(size_t)c * A.n_trials + q];
                S.thr[t_] = A.exact_exp ? tr.u : -T * log(tr.u);
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
                S.thr[t_] = A.exact_exp ? tr.u : -T * log(tr.u);
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
                } else {  // slots in the species lists; resolved to particles when the trial executes
                    const int nA = S.spoff[A.mv_a[m] + 1] - S.spoff[A.mv_a[m]];
                    const int nB = S.spoff[A.mv_b[m] + 1] - S.spoff[A.mv_b[m]];
                    tr.i = (nA > 0 && nB > 0) ? (int)bounded(a.v[1], (uint32_t)nA) : -1;
                    tr.j = (nA > 0 && nB > 0) ? (int)bounded(b.v[0], (uint32_t)nB) : -1;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                }
                if (A.trace) A.trace[(size_t)c * A.n_trials + q] = tr;
            }
            S.tm[t_] = tr.move;
            S.ti[t_] = tr.i;
            S.tj[t_] = (tr.kind == PMC_MOVE_SWAP) ? tr.j : -2;  // -2 marks a displacement
#pragma unroll
            for (int a = 0; a < 3; a++) {
                S.delta[3 * t_ + a] = tr.delta[a];
                if constexpr (FILTER) S.dint[3 * t_ + a] = (int)__double2ll_rn(tr.delta[a] * fscale);
            }
        }
        __syncthreads();

        // ---- the serial chain: one trial at a time ---------------------------------------------
        for (int b = 0; b < nb; b++) {
            const int m = S.tm[b];
            const bool is_disp = S.tj[b] == -2;
            double part = 0.0, dE;
            bool acc;
            if (is_disp) {
                const int i = S.ti[b];
                double xo[3] = {0.0, 0.0, 0.0}, xn[3] = {0.0, 0.0, 0.0};
                int w[3] = {0, 0, 0};
                // The owner thread (i % NT) commits accepted positions after the barrier of the
                // previous trial; every other thread only ever reads x[i] here, at the start of a
                // trial.  A commit becomes visible at the NEXT barrier, so the only unordered read
                // is "same particle as the most recent commit" -- served from registers instead.
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    xo[a] = (i == last_i) ? last_x[a] : S.x[a * Npad + i];
                    xn[a] = wrap1(xo[a] + S.delta[3 * b + a], L[a], w[a]);
                }
                const int si = S.sp[i];
                uint16_t bi[PMC_MAX_BONDS];
                if constexpr (MOL) {
#pragma unroll
                    for (int k = 0; k < PMC_MAX_BONDS; k++) bi[k] = S.bonds[i * PMC_MAX_BONDS + k];
                }
                const double *prow = S.par + si * ns * PMC_NPAR;
                uint32_t un[3] = {0u, 0u, 0u};
                if constexpr (FILTER) {
                    uint32_t uo[3] = {0u, 0u, 0u};
#pragma unroll
                    for (int a = 0; a < DIM; a++) {
                        uo[a] = (i == last_i) ? last_u[a] : S.u[a * Npad + i];
                        un[a] = uo[a] + (uint32_t)S.dint[3 * b + a];  // wraps like the box does
                    }
                    const uint32_t thr = S.thr_u[si];
                    int cnt = 0;
                    for (int j0 = 0; j0 < Npad; j0 += NT) {
                        const int j = j0 + tid;
                        bool pass = false;
                        if (j < N && j != i) {
                            pass = dist2_fixed<DIM>(S.u, Npad, j, uo) <= thr || dist2_fixed<DIM>(S.u, Npad, j, un) <= thr ||
                                   bonded_to<MOL>(bi, j);
                        }
                        const unsigned mask = __ballot_sync(0xffffffffu, pass);
                        if (pass) myq[cnt + __popc(mask & lt_mask)] = (uint16_t)j;
                        cnt += __popc(mask);
                    }
                    __syncwarp();
                    for (int q = lane; q < cnt; q += 32)
                        part += displacement_term<DIM, MODEL, MOL>(S, Npad, myq[q], xo, xn, L, prow, bi);
                    __syncwarp();
                } else {
                    for (int j = tid; j < N; j += NT)
                        if (j != i) part += displacement_term<DIM, MODEL, MOL>(S, Npad, j, xo, xn, L, prow, bi);
                }
                dE = block_sum(part, S.red, slot, lane, warp, nwarp);
                slot ^= 1;
                acc = A.exact_exp ? accept_exact(dE, T, S.thr[b]) : (dE < S.thr[b]);
                if (acc) {
                    if (tid == i % NT) {
#pragma unroll
                        for (int a = 0; a < DIM; a++) {
                            S.x[a * Npad + i] = xn[a];
                            if constexpr (FILTER) S.u[a * Npad + i] = to_fixed(xn[a], fscale);
                            if (w[a] != 0) atomicAdd(&gimg[a * Npad + i], w[a]);
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
                const int ka = S.ti[b], kb = S.tj[b];
                int i = -1, j = -1;
                if (A.replay) {
                    i = ka;
                    j = kb;
                } else if (ka >= 0) {
                    i = S.spids[S.spoff[A.mv_a[m]] + ka];
                    j = S.spids[S.spoff[A.mv_b[m]] + kb];
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
                        const uint32_t thr = S.thr_u[PMC_MAX_SPECIES];
                        int cnt = 0;
                        for (int k0_ = 0; k0_ < Npad; k0_ += NT) {
                            const int k = k0_ + tid;
                            bool pass = false;
                            if (k < N) {
                                pass = (k != i && dist2_fixed<DIM>(S.u, Npad, k, ui) <= thr) ||
                                       (k != j && dist2_fixed<DIM>(S.u, Npad, k, uj) <= thr) || bonded_to<MOL>(bi, k) ||
                                       bonded_to<MOL>(bj, k);
                            }
                            const unsigned mask = __ballot_sync(0xffffffffu, pass);
                            if (pass) myq[cnt + __popc(mask & lt_mask)] = (uint16_t)k;
                            cnt += __popc(mask);
                        }
                        __syncwarp();
                        for (int q = lane; q < cnt; q += 32)
                            part += swap_term<DIM, MODEL, MOL>(S, Npad, ns, myq[q], i, j, si, sj, xi, xj, L, bi, bj);
                        __syncwarp();
                    } else {
                        for (int k = tid; k < N; k += NT)
                            part += swap_term<DIM, MODEL, MOL>(S, Npad, ns, k, i, j, si, sj, xi, xj, L, bi, bj);
                    }
                }
                dE = block_sum(part, S.red, slot, lane, warp, nwarp);
                slot ^= 1;
                acc = (i >= 0 && j >= 0) && (A.exact_exp ? accept_exact(dE, T, S.thr[b]) : (dE < S.thr[b]));
                if (acc) {
                    if (tid == 0) {
                        const uint8_t si = S.sp[i], sj = S.sp[j];
                        S.sp[i] = sj;
                        S.sp[j] = si;
                        if (A.any_swap) {  // update_species_list! (src/moves.jl:175-179)
                            const uint16_t hi = S.heads[i], hj = S.heads[j];
                            S.spids[S.spoff[si] + hi] = (uint16_t)j;
                            S.spids[S.spoff[sj] + hj] = (uint16_t)i;
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
                S.cnt[m] += 1ull;
                S.cnt[PMC_MAX_MOVES + m] += acc ? 1ull : 0ull;
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
        A.calls[(size_t)c * PMC_MAX_MOVES + tid] += S.cnt[tid];
        A.accepted[(size_t)c * PMC_MAX_MOVES + tid] += S.cnt[PMC_MAX_MOVES + tid];
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

size_t chain_sweep_smem_bytes(int dim, int Npad, int ns, int threads, bool mol, bool any_swap, bool filter) {
    SweepSmem s;
    return carve_sweep(s, nullptr, dim, Npad, ns, threads, mol, any_swap, filter);
}

size_t chain_energy_smem_bytes(int dim, int Npad, int ns, bool) {
    return sizeof(double) * ((size_t)dim * Npad + Npad + (size_t)ns * ns * PMC_NPAR) + Npad + 16;
}

cudaError_t configure_chain_kernels(int dim, int model, bool mol, size_t sweep_smem, size_t sweep_smem_filter,
                                    size_t energy_smem) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        constexpr bool ml = decltype(ML)::value;
        cudaError_t e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sweep_smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_chain_sweep<d, mdl, ml, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sweep_smem_filter);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_chain_energy<d, mdl, ml>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)energy_smem);
    });
}

cudaError_t launch_chain_sweep(int dim, int model, bool mol, bool filter, int M, int threads, size_t smem,
                               const ChainArgs &a, cudaStream_t st) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        constexpr bool ml = decltype(ML)::value;
        if (filter)
            k_chain_sweep<d, mdl, ml, true><<<M, threads, smem, st>>>(a);
        else
            k_chain_sweep<d, mdl, ml, false><<<M, threads, smem, st>>>(a);
        return cudaGetLastError();
    });
}

cudaError_t launch_chain_energy(int dim, int model, bool mol, int M, size_t smem, const EnergyArgs &a,
                                cudaStream_t st) {
    return dispatch(dim, model, mol, [&](auto D, auto MDL, auto ML) {
        k_chain_energy<decltype(D)::value, decltype(MDL)::value, decltype(ML)::value><<<M, 256, smem, st>>>(a);
        return cudaGetLastError();
    });
}

}  // namespace pmc
