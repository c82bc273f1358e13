// chains_fast.cuh -- hand-scheduled sweep kernel for the dominant PMC_MODE_CHAINS case:
// Atoms (no bonds), Displacement-only pools, cubic box, N <= kFastCand * kFastThreads (= 1024).
// Same algorithm and the same fp64 pair terms as k_chain_sweep<FILTER = true> in chains.cu (which remains the
// general kernel: swaps, molecules, larger N, non-cubic boxes); what changes is how the work is issued:
//
//   * ALL shared-memory traffic goes through explicit 32-bit shared addresses (ld.shared / st.shared), one
//     base register + constant offsets, instead of ~20 generic pointers the compiler kept re-deriving;
//   * each thread keeps the fixed-point coordinates of its 8 candidates in registers, so the candidate scan
//     is 3 subtract + 3 mul-hi + compare per candidate with no loads;
//   * survivors are compacted with ONE warp prefix sum per trial (per-thread bit mask -> shuffle scan)
//     instead of one ballot + popc + store sequence per candidate;
//   * no forwarding registers: every thread performs the (identical) commit stores, so each thread's own
//     program order makes committed positions visible to it without a second barrier.
//
// The kernel is issue-bound (ncu: profiles/), so instruction count per trial is the figure of merit -- but not the only
// one: a 96-thread variant (11 candidates per thread, 21 % fewer instructions per trial) measured 11 % SLOWER than
// 128 threads at N = 1000 because three warps per chain hide less latency (7 x 3 vs 6 x 4 warps per SM).
#pragma once
#include "chains.cuh"
#include "common.cuh"
#include "rng.cuh"

namespace pmc {
namespace fast {

constexpr int kFastThreads = 128;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastMaxCand = 8; // candidates per thread (registers) at the largest padded size
constexpr int kFastBatch = 32;  // parked proposals
constexpr int kRecBytes = 80;   // trial record stride

// ---- explicit shared-memory accessors (32-bit shared-window addresses) ---------------------------------
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void lds_f64x2(uint32_t a, double &v0, double &v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a) : "memory");
}
__device__ __forceinline__ void lds_s32x4(uint32_t a, int &v0, int &v1, int &v2, int &v3) {
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// Dynamic shared-memory layout (byte offsets from the base), all 16-byte aligned.
struct FastLayout {
    uint32_t x, sp, q, cp, rec, red, cnt, par, rcs, spids, heads, spoff, total;
};
__host__ __device__ inline FastLayout fast_layout(int dim, int Npad, int ns, bool swaps, bool full_par) {
    FastLayout f;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        uint32_t p = o;
        o += (bytes + 15u) & ~15u;
        return p;
    };
    f.x = take(8u * dim * Npad);
    f.sp = take(Npad);
    f.q = take(2u * (uint32_t)Npad);                       // per-warp survivor queues (worst case: all pass)
    f.cp = take(32u * ns * ns);                            // compact LJ pair table {rc2, eps4, sig2, shift}
    f.rec = take((uint32_t)kRecBytes * kFastBatch);        // parked proposals
    f.red = take(8u * 2 * kFastWarps);
    f.cnt = take(8u * 2 * PMC_MAX_MOVES);
    f.par = take(full_par ? 8u * ns * ns * PMC_NPAR : 0u); // full parameter table (non-LJ models only: 7 CTAs/SM need < 32 KB)
    f.rcs = take(8u * PMC_MAX_SPECIES);                    // largest cutoff radius per species of the moved particle
    f.spids = take(swaps ? 2u * (uint32_t)Npad : 0u);      // SpeciesList (src/utils.jl:31-49), DiscreteSwap only
    f.heads = take(swaps ? 2u * (uint32_t)Npad : 0u);
    f.spoff = take(32u);                                   // species offsets [5] + fixed-point global cutoff [1]
    f.total = o;
    return f;
}

// floor(x * scale) mod 2^32 in ONE instruction: DFMA.RM against 2^52 leaves the integer in the low mantissa word
// (the product is exact inside the fma, so this is the exact floor; DMUL + F2I.U64.FLOOR costs three issue slots more)
__device__ __forceinline__ uint32_t to_fixed32(double x, double scale) {
    return (uint32_t)__double2loint(__fma_rd(x, scale, 4503599627370496.0));
}

// wrapped-coordinate nearest image, squared, accumulated.  |a - L| == L - a and the square drops the sign, so this
// is fmin(a, L - a)^2 bit for bit, issued as compare + subtract + select (fmin() costs six instructions in fp64:
// DSETP.MIN + two selects + NaN fix-up + moves).
__device__ __forceinline__ double mi_acc(double xi, double xj, double L, double hL, double acc) {
    const double a = xi - xj;
    const double r = fabs(a) > hL ? fabs(a) - L : a;
    return fma(r, r, acc);
}
// x in [-L, 2L) -> [0, L): the two sequential folds of the displacement move as predicated adds
__device__ __forceinline__ double wrap_once(double x, double L) {
    asm("{ .reg .pred p, q; setp.ge.f64 p, %0, %1; @p sub.f64 %0, %0, %1; setp.lt.f64 q, %0, 0d0000000000000000; @q add.f64 %0, %0, %1; }"
        : "+d"(x)
        : "d"(L));
    return x;
}

// NPAD (compile time): padded particle count of the shared-memory planes, one of 256 / 512 / 1024 (/ 2048 in 2-D); makes every
// shared-memory offset a constant and fixes the number of register-resident candidates per thread.
template <int DIM, int MODEL, int NPAD, bool SWAPS, int NT = kFastThreads>
__global__ void __launch_bounds__(NT, SWAPS ? 6 : 7) k_chain_sweep_fast(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kThreads = NT, kWarps = NT / 32;  // CTA size is a compile-time constant of this kernel
    constexpr int kFastCand = NPAD / kThreads;
    constexpr int kImgThread = kThreads > 32 ? 32 : 0, kCntThread = kThreads > 64 ? 64 : 0;  // bookkeeping lanes
    static_assert(NPAD % NT == 0 && NT % 32 == 0 && NT <= 128, "NPAD must be a multiple of the CTA size");
    constexpr int Npad = NPAD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const int N = A.N, gNpad = A.Npad, ns = A.ns;  // gNpad: stride of the GLOBAL arrays (multiple of 32)
    constexpr bool kFullPar = !(MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG);
    const FastLayout F = fast_layout(DIM, Npad, PMC_MAX_SPECIES, SWAPS, kFullPar);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t nb8 = 8u * (uint32_t)Npad;  // byte stride between coordinate planes

    // ---- load chain state -----------------------------------------------------------------------------
    const double L = A.box[c * 3], hL = 0.5 * L;
    const double fscale = 4294967296.0 / L;
    double *gx = A.x + (size_t)c * DIM * gNpad;
    {
        double *sx = (double *)(smem_raw + F.x);
        for (int a = 0; a < DIM; a++)
            for (int k = tid; k < Npad; k += kThreads) sx[a * Npad + k] = k < gNpad ? gx[a * gNpad + k] : 0.0;
        uint8_t *ssp = smem_raw + F.sp;
        const uint8_t *gsp = A.sp + (size_t)c * gNpad;
        for (int k = tid; k < Npad; k += kThreads) ssp[k] = k < gNpad ? gsp[k] : 0;
        double *spar = (double *)(smem_raw + F.par), *scp = (double *)(smem_raw + F.cp);
        if constexpr (kFullPar)
            for (int k = tid; k < ns * ns * PMC_NPAR; k += kThreads) spar[k] = A.par[k];
        for (int k = tid; k < ns * ns; k += kThreads) {
            scp[4 * k + 0] = A.par[k * PMC_NPAR + PMC_P_RCUT2];
            scp[4 * k + 1] = A.par[k * PMC_NPAR + PMC_P_EPS];
            scp[4 * k + 2] = A.par[k * PMC_NPAR + PMC_P_SIG2];
            scp[4 * k + 3] = A.par[k * PMC_NPAR + PMC_P_SHIFT];
        }
        unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
        if (tid < 2 * PMC_MAX_MOVES) scnt[tid] = 0ull;
        int *sso = (int *)(smem_raw + F.spoff);
        if (tid <= PMC_MAX_SPECIES) sso[tid] = A.spoff[c * (PMC_MAX_SPECIES + 1) + tid];
        if (tid == 0) {  // conservative fixed-point cutoff over all species pairs (swap filter)
            double rc2 = 0.0;
            for (int k = 0; k < ns * ns; k++) rc2 = fmax(rc2, A.par[k * PMC_NPAR + PMC_P_RCUT2]);
            ((uint32_t *)sso)[PMC_MAX_SPECIES + 1] = neg_thr8(sqrt(rc2) * fscale * 0x1p-24);
        }
        if constexpr (SWAPS) {
            uint16_t *si_ = (uint16_t *)(smem_raw + F.spids), *sh_ = (uint16_t *)(smem_raw + F.heads);
            const uint16_t *gi = A.spids + (size_t)c * gNpad, *gh = A.heads + (size_t)c * gNpad;
            for (int k = tid; k < Npad; k += kThreads) {
                si_[k] = k < gNpad ? gi[k] : 0;
                sh_[k] = k < gNpad ? gh[k] : 0;
            }
        }
    }
    // fixed-point coordinates of this thread's candidates j = k * 128 + tid, k = 0..7 (registers)
    uint32_t myq[kFastCand];
#pragma unroll
    for (int k = 0; k < kFastCand; k++) {
        const int j = k * kThreads + tid;
        uint32_t u[3] = {0u, 0u, 0u};
#pragma unroll
        for (int a = 0; a < DIM; a++) u[a] = j < gNpad ? to_fixed32(gx[a * gNpad + j], fscale) : 0u;
        myq[k] = pack8(u[0], u[1], u[2]);
    }
    if (tid < PMC_MAX_SPECIES) {  // largest cutoff radius per species of the moved particle (filter sphere)
        double rc2 = 0.0;
        for (int b = 0; b < ns; b++) rc2 = fmax(rc2, A.par[((tid < ns ? tid : 0) * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
        ((double *)(smem_raw + F.rcs))[tid] = sqrt(rc2);
    }
    const double Tk = A.temp[c];
    double E = A.energy[c];
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t gchain = (uint32_t)(A.chain_offset + c);
    int32_t *gimg = A.img + (size_t)c * DIM * gNpad;
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * kFastCand * 32);
    uint32_t slot = 0;

    for (long long tb = 0; tb < A.n_trials; tb += kFastBatch) {
        const int nb = (int)min((long long)kFastBatch, A.n_trials - tb);
        __syncthreads();
        // ---- proposals of trials tb .. tb+nb-1 (one warp generates, records parked in shared memory) -----
        if (tid < nb) {
            const long long q = tb + tid;
            pmc_trial tr;
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
                } else {  // slots in the species lists; resolved to particles when the trial executes
                    const int *sso = (const int *)(smem_raw + F.spoff);
                    const int nA = sso[A.mv_a[m] + 1] - sso[A.mv_a[m]], nB = sso[A.mv_b[m] + 1] - sso[A.mv_b[m]];
                    tr.i = (nA > 0 && nB > 0) ? (int)bounded(a.v[1], (uint32_t)nA) : -1;
                    tr.j = (nA > 0 && nB > 0) ? (int)bounded(b.v[0], (uint32_t)nB) : -1;
                    tr.delta[0] = tr.delta[1] = tr.delta[2] = 0.0;
                }
                if (A.trace) A.trace[(size_t)c * A.n_trials + q] = tr;
            }
            // record layout: f64 delta[3], f64 thr | s32 dint[3], s32 i | s32 m, kind, j, pool species | u32 thr_t[4]
            unsigned char *rec = smem_raw + F.rec + (size_t)kRecBytes * tid;
            double *rd = (double *)rec;
            int *ri = (int *)(rec + 32);
            uint32_t *rt = (uint32_t *)(rec + 64);
            rd[0] = tr.delta[0];
            rd[1] = tr.delta[1];
            rd[2] = tr.delta[2];
            rd[3] = A.exact_exp ? tr.u : -Tk * log(tr.u);
            ri[0] = (int)__double2ll_rn(tr.delta[0] * fscale);
            ri[1] = (int)__double2ll_rn(tr.delta[1] * fscale);
            ri[2] = (int)__double2ll_rn(tr.delta[2] * fscale);
            ri[3] = tr.i;
            ri[4] = tr.move;
            ri[5] = tr.kind;
            ri[6] = tr.j;
            ri[7] = (tr.kind == PMC_MOVE_SWAP && !A.replay) ? (A.mv_a[tr.move] | (A.mv_b[tr.move] << 8)) : 0;
            // one sphere around the midpoint of old and new position covers both cutoff spheres
            const double hd = 0.5 * sqrt(tr.delta[0] * tr.delta[0] + tr.delta[1] * tr.delta[1] + tr.delta[2] * tr.delta[2]);
            const double *rcs = (const double *)(smem_raw + F.rcs);
#pragma unroll
            for (int s = 0; s < PMC_MAX_SPECIES; s++) {
                rt[s] = neg_thr8((rcs[s] + hd) * fscale * 0x1p-24);
            }
        }
        __syncthreads();

        // ---- the serial chain -------------------------------------------------------------------------------
        for (int b = 0; b < nb; b++) {
            const uint32_t ra = sb + F.rec + (uint32_t)kRecBytes * (uint32_t)b;
            double d0, d1, d2, thr;
            int di0, di1, di2, i;
            lds_f64x2(ra, d0, d1);
            lds_f64x2(ra + 16, d2, thr);
            lds_s32x4(ra + 32, di0, di1, di2, i);
            if constexpr (SWAPS) {
                int mv, kind, j, sab;
                lds_s32x4(ra + 48, mv, kind, j, sab);
                if (kind == PMC_MOVE_SWAP) {
                    // ---- DiscreteSwap: positions fixed, four local energies in one pass (src/moves.jl:159-167) ----
                    const uint32_t soa = sb + F.spoff;
                    if (!A.replay && i >= 0) {  // slots -> particles through the species lists
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
                    // list slots are read BEFORE the barrier: the commit below overwrites them (read-modify-write)
                    const uint32_t hi = lds_u16(sb + F.heads + 2u * iu), hj = lds_u16(sb + F.heads + 2u * ju);
                    const uint32_t oi = lds_u32(soa + 4u * si), oj = lds_u32(soa + 4u * sj);
                    const uint32_t ui0 = to_fixed32(xi0, fscale), ui1 = to_fixed32(xi1, fscale), ui2 = DIM == 3 ? to_fixed32(xi2, fscale) : 0u;
                    const uint32_t uj0 = to_fixed32(xj0, fscale), uj1 = to_fixed32(xj1, fscale), uj2 = DIM == 3 ? to_fixed32(xj2, fscale) : 0u;
                    const int gthr = (int)lds_u32(soa + 4u * (PMC_MAX_SPECIES + 1));
                    const uint32_t uiq = pack8(ui0, ui1, ui2), ujq = pack8(uj0, uj1, uj2);
                    uint32_t m8 = 0;
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < kFastCand; k++) {  // survivor of either sphere: bit kFastCand-1-k
                            const uint32_t t1 = __vabsdiffu4(uiq, myq[k]), t2 = __vabsdiffu4(ujq, myq[k]);
                            const int v = __dp4a((int)t1, (int)t1, gthr) | __dp4a((int)t2, (int)t2, gthr);
                            m8 = __funnelshift_l((uint32_t)v, m8, 1);
                        }
                    }
                    const int mine = __popc(m8);
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        incl += (lane >= o) ? t : 0;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
                    for (int k = 0; k < kFastCand; k++) {
                        if (m8 & (1u << (kFastCand - 1 - k))) {
                            sts_u16(wp, (uint32_t)(k * kThreads + tid));
                            wp += 2;
                        }
                    }
                    __syncwarp();
                    double part = 0.0;
                    auto pair_e = [&](uint32_t sa, uint32_t sb_, double r2) -> double {
                        if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                            double rc2, eps4, sig2, shift;
                            const uint32_t pa = sb + F.cp + 32u * (sa * (uint32_t)ns + sb_);
                            lds_f64x2(pa, rc2, eps4);
                            lds_f64x2(pa + 16, sig2, shift);
                            return r2 <= rc2 ? lj_core(r2, eps4, sig2) - shift : 0.0;
                        } else {
                            const double *p = (const double *)(smem_raw + F.par) + (sa * (uint32_t)ns + sb_) * PMC_NPAR;
                            return r2 <= p[PMC_P_RCUT2] ? pair_potential<MODEL>(p, r2) : 0.0;
                        }
                    };
                    for (int q = lane; q < total; q += 32) {
                        const uint32_t k = lds_u16(qa + 2u * (uint32_t)q);
                        if (k < (uint32_t)N) {
                            const uint32_t ka = sb + F.x + 8u * k;
                            const double xk0 = lds_f64(ka), xk1 = lds_f64(ka + nb8), xk2 = DIM == 3 ? lds_f64(ka + 2 * nb8) : 0.0;
                            const uint32_t sk = lds_u8(sb + F.sp + k);
                            const uint32_t skn = k == iu ? sj : (k == ju ? si : sk);  // species of k after the exchange
                            if (k != iu) {  // k-term of particle i's local energy: (si, sk) -> (sj, sk')
                                double r2 = mi_acc(xi0, xk0, L, hL, 0.0);
                                r2 = mi_acc(xi1, xk1, L, hL, r2);
                                if constexpr (DIM == 3) r2 = mi_acc(xi2, xk2, L, hL, r2);
                                part += pair_e(sj, skn, r2) - pair_e(si, sk, r2);
                            }
                            if (k != ju) {  // k-term of particle j's local energy: (sj, sk) -> (si, sk')
                                double r2 = mi_acc(xj0, xk0, L, hL, 0.0);
                                r2 = mi_acc(xj1, xk1, L, hL, r2);
                                if constexpr (DIM == 3) r2 = mi_acc(xj2, xk2, L, hL, r2);
                                part += pair_e(si, skn, r2) - pair_e(sj, sk, r2);
                            }
                        }
                    }
                    __syncwarp();
                    part = warp_sum(part);
                    const uint32_t rda = sb + F.red + 32u * slot;
                    if (lane == 0) sts_f64(rda + 8u * (uint32_t)warp, part);
                    __syncthreads();
                    double s0, s1, s2 = 0.0, s3 = 0.0;
                    lds_f64x2(rda, s0, s1);
                    if constexpr (kWarps > 2) lds_f64x2(rda + 16, s2, s3);
                    const double dE = kWarps == 4 ? ((s0 + s1) + s2) + s3 : (kWarps == 3 ? (s0 + s1) + s2 : (kWarps == 2 ? s0 + s1 : s0));
                    slot ^= 1u;
                    const bool acc = valid && (A.exact_exp ? accept_exact(dE, Tk, thr) : (dE < thr));
                    if (acc) {  // every thread performs the identical stores (see the displacement commit)
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + iu), "r"(sj) : "memory");
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(sb + F.sp + ju), "r"(si) : "memory");
                        sts_u16(sb + F.spids + 2u * (oi + hi), ju);  // update_species_list! (src/moves.jl:175-179)
                        sts_u16(sb + F.spids + 2u * (oj + hj), iu);
                        sts_u16(sb + F.heads + 2u * iu, hj);
                        sts_u16(sb + F.heads + 2u * ju, hi);
                        E += dE;
                    }
                    if (tid == kCntThread) {
                        unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
                        scnt[mv] += 1ull;
                        scnt[PMC_MAX_MOVES + mv] += acc ? 1ull : 0ull;
                        if (A.acc_out) A.acc_out[(size_t)c * A.n_trials + tb + b] = acc ? 1 : 0;
                        if (A.dE_out) A.dE_out[(size_t)c * A.n_trials + tb + b] = dE;
                        if (A.trace) {
                            pmc_trial *tr = A.trace + (size_t)c * A.n_trials + tb + b;
                            tr->i = i;
                            tr->j = j;
                        }
                    }
                    continue;
                }
            }
            const uint32_t xa = sb + F.x + 8u * (uint32_t)i;
            double xo[3], xn[3];
            xo[0] = lds_f64(xa);
            xo[1] = lds_f64(xa + nb8);
            xo[2] = (DIM == 3) ? lds_f64(xa + 2 * nb8) : 0.0;
            const uint32_t si = lds_u8(sb + F.sp + (uint32_t)i);
            xn[0] = xo[0] + d0;
            xn[1] = xo[1] + d1;
            xn[2] = xo[2] + d2;
#pragma unroll
            for (int a = 0; a < DIM; a++) xn[a] = wrap_once(xn[a], L);
            // filter sphere: midpoint of old and new in fixed point, radius from the parked record
            const uint32_t um0 = to_fixed32(xo[0], fscale) + (uint32_t)(di0 >> 1);
            const uint32_t um1 = to_fixed32(xo[1], fscale) + (uint32_t)(di1 >> 1);
            const uint32_t um2 = (DIM == 3) ? to_fixed32(xo[2], fscale) + (uint32_t)(di2 >> 1) : 0u;
            const int fthr = (int)lds_u32(ra + 64 + 4u * si);
            const uint32_t umq = pack8(um0, um1, um2);
            uint32_t m8 = 0;
#pragma unroll
            for (int k = 0; k < kFastCand; k++) {  // survivor: bit kFastCand-1-k
                const uint32_t t = __vabsdiffu4(umq, myq[k]);
                m8 = __funnelshift_l((uint32_t)__dp4a((int)t, (int)t, fthr), m8, 1);
            }
            // compaction: warp prefix sum of the per-thread survivor counts
            const int mine = __popc(m8);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                incl += (lane >= o) ? t : 0;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
            for (int k = 0; k < kFastCand; k++) {
                if (m8 & (1u << (kFastCand - 1 - k))) {
                    sts_u16(wp, (uint32_t)(k * kThreads + tid));
                    wp += 2;
                }
            }
            __syncwarp();
            // fp64 pass over this warp's survivors
            double part = 0.0;
            const uint32_t prow = si * (uint32_t)ns;
            for (int q = lane; q < total; q += 32) {
                const uint32_t j = lds_u16(qa + 2u * (uint32_t)q);
                if (j < (uint32_t)N && j != (uint32_t)i) {
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
                    if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                        double rc2, eps4, sig2, shift;
                        const uint32_t pa = sb + F.cp + 32u * (prow + sj);
                        lds_f64x2(pa, rc2, eps4);
                        lds_f64x2(pa + 16, sig2, shift);
                        const double uo = lj_core(r2o, eps4, sig2) - shift;
                        const double un = lj_core(r2n, eps4, sig2) - shift;
                        part += (r2n <= rc2 ? un : 0.0) - (r2o <= rc2 ? uo : 0.0);
                    } else {
                        const double *p = (const double *)(smem_raw + F.par) + (prow + sj) * PMC_NPAR;
                        const double rc2 = p[PMC_P_RCUT2];
                        if (r2o <= rc2) part -= pair_potential<MODEL>(p, r2o);
                        if (r2n <= rc2) part += pair_potential<MODEL>(p, r2n);
                    }
                }
            }
            __syncwarp();
            // block-wide sum, one barrier, identical bits in every thread; slots alternate between trials
            part = warp_sum(part);
            const uint32_t rda = sb + F.red + 32u * slot;
            if (lane == 0) sts_f64(rda + 8u * (uint32_t)warp, part);
            __syncthreads();
            double s0, s1, s2 = 0.0, s3 = 0.0;
            lds_f64x2(rda, s0, s1);
            if constexpr (kWarps > 2) lds_f64x2(rda + 16, s2, s3);
            const double dE = kWarps == 4 ? ((s0 + s1) + s2) + s3 : (kWarps == 3 ? (s0 + s1) + s2 : (kWarps == 2 ? s0 + s1 : s0));
            slot ^= 1u;
            const bool acc = A.exact_exp ? accept_exact(dE, Tk, thr) : (dE < thr);
            if (acc) {
                // every thread stores the (identical) committed position: its own later reads are ordered after
                // its own store, so no forwarding registers and no second barrier are needed
                sts_f64(xa, xn[0]);
                sts_f64(xa + nb8, xn[1]);
                if constexpr (DIM == 3) sts_f64(xa + 2 * nb8, xn[2]);
                E += dE;
                if (tid == (i % kThreads)) {  // owner refreshes its register copy
                    const int ki = i / kThreads;
                    const uint32_t f = pack8(to_fixed32(xn[0], fscale), to_fixed32(xn[1], fscale), DIM == 3 ? to_fixed32(xn[2], fscale) : 0u);
#pragma unroll
                    for (int k = 0; k < kFastCand; k++) myq[k] = k == ki ? f : myq[k];
                }
                // image counters (which way did the coordinate wrap) and move counters are kept by threads of
                // different warps: every warp waits at the next barrier, so the per-trial bookkeeping is spread out
                if (tid == kImgThread) {
                    const double t0 = xo[0] + d0, t1 = xo[1] + d1, t2 = xo[2] + d2;
                    const int w0 = (t0 >= L) - (t0 < 0.0), w1 = (t1 >= L) - (t1 < 0.0), w2 = (t2 >= L) - (t2 < 0.0);
                    if (w0) atomicAdd(&gimg[i], w0);
                    if (w1) atomicAdd(&gimg[gNpad + i], w1);
                    if (DIM == 3 && w2) atomicAdd(&gimg[2 * gNpad + i], w2);
                }
            }
            if (tid == kCntThread) {
                unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
                const int m = (int)lds_u32(ra + 48);
                scnt[m] += 1ull;
                scnt[PMC_MAX_MOVES + m] += acc ? 1ull : 0ull;
                if (A.acc_out) A.acc_out[(size_t)c * A.n_trials + tb + b] = acc ? 1 : 0;
                if (A.dE_out) A.dE_out[(size_t)c * A.n_trials + tb + b] = dE;
            }
        }
    }
    __syncthreads();
    {
        const double *sx = (const double *)(smem_raw + F.x);
        for (int a = 0; a < DIM; a++)
            for (int k = tid; k < gNpad; k += kThreads) gx[a * gNpad + k] = sx[a * Npad + k];
        if constexpr (SWAPS) {
            uint8_t *gsp = A.sp + (size_t)c * gNpad;
            uint16_t *gi = A.spids + (size_t)c * gNpad, *gh = A.heads + (size_t)c * gNpad;
            const uint16_t *si_ = (const uint16_t *)(smem_raw + F.spids), *sh_ = (const uint16_t *)(smem_raw + F.heads);
            for (int k = tid; k < gNpad; k += kThreads) {
                gsp[k] = smem_raw[F.sp + k];
                gi[k] = si_[k];
                gh[k] = sh_[k];
            }
        }
        const unsigned long long *scnt = (const unsigned long long *)(smem_raw + F.cnt);
        if (tid == 0) A.energy[c] = E;
        if (tid < A.n_moves) {
            A.calls[(size_t)c * PMC_MAX_MOVES + tid] += scnt[tid];
            A.accepted[(size_t)c * PMC_MAX_MOVES + tid] += scnt[PMC_MAX_MOVES + tid];
        }
    }
}


// ==================================================================================================
// PMC_MIXED: float32 pair terms, float64 accumulation (north star: "reported separately at 1e-6").
// During a launch the chain state IS the 32-bit fixed-point representation (resolution L / 2^32 ~ 2e-9):
// proposals are applied in fixed point, squared distances come from the same wrapping integer arithmetic as
// the prefilter (minimum image for free, relative resolution ~2e-8), the pair potential is evaluated in fp32
// (MUFU.RCP instead of a 5-instruction fp64 reciprocal), per-lane partial sums are fp32 (<= 3 terms) and
// everything across lanes / warps / trials is accumulated in fp64.  No fp64 positions in shared memory:
// 18 KB per chain instead of 31 KB.
// ==================================================================================================
struct MixedLayout {
    uint32_t u, sp, q, cp, rec, red, cnt, rcs, total;
};
__host__ __device__ inline MixedLayout mixed_layout(int dim, int Npad) {
    MixedLayout f;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        uint32_t p = o;
        o += (bytes + 15u) & ~15u;
        return p;
    };
    f.u = take(4u * dim * Npad);
    f.sp = take(Npad);
    f.q = take(2u * (uint32_t)Npad);
    f.cp = take(32u * PMC_MAX_SPECIES * PMC_MAX_SPECIES);  // float table {rc2, eps4|eps, sig2, shift, c0|ndiv2, c2, c4, -}
    f.rec = take((uint32_t)kRecBytes * kFastBatch);
    f.red = take(8u * 2 * kFastWarps);
    f.cnt = take(8u * 2 * PMC_MAX_MOVES);
    f.rcs = take(8u * PMC_MAX_SPECIES);
    f.total = o;
    return f;
}

__device__ __forceinline__ void lds_f32x4(uint32_t a, float &v0, float &v1, float &v2, float &v3) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <int DIM>
__device__ __forceinline__ uint32_t dist2_u32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t a0, uint32_t a1, uint32_t a2) {
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

// fp32 pair potential, parameters p0..p6 = {eps4|eps, sig2, shift, c0|ndiv2, c2, c4}
template <int MODEL>
__device__ __forceinline__ float pair_potential_f32(float r2, float eps, float sig2, float shift, float c0, float c2, float c4) {
    const float x = sig2 * __frcp_rn(r2);
    if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
        const float x3 = x * x * x;
        return fmaf(eps, fmaf(x3, x3, -x3), -shift);
    } else if constexpr (MODEL == PMC_MODEL_SMOOTHLJ) {
        const float x3 = x * x * x;
        return eps * (fmaf(x3, x3, -x3) + c0 + r2 * fmaf(r2, c4, c2));
    } else {
        const int n = (int)c0;
        float v;
        if ((float)n == c0) {
            v = 1.0f;
            float b = x;
            for (int k = n; k > 0; k >>= 1) {
                if (k & 1) v *= b;
                b *= b;
            }
        } else {
            v = powf(x, c0);
        }
        return fmaf(eps, v, -shift);
    }
}

template <int DIM, int MODEL, int NPAD>
__global__ void __launch_bounds__(kFastThreads, 8) k_chain_sweep_mixed(const __grid_constant__ ChainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kFastCand = NPAD / kFastThreads;
    constexpr int Npad = NPAD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const int N = A.N, gNpad = A.Npad, ns = A.ns;
    const MixedLayout F = mixed_layout(DIM, Npad);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    constexpr uint32_t nb4 = 4u * (uint32_t)NPAD;

    const double L = A.box[c * 3];
    const double fscale = 4294967296.0 / L;
    const float r2scale = (float)(L * L * 0x1p-32);  // fixed-point r^2 units -> length^2
    double *gx = A.x + (size_t)c * DIM * gNpad;
    {
        uint32_t *su = (uint32_t *)(smem_raw + F.u);
        for (int a = 0; a < DIM; a++)
            for (int k = tid; k < Npad; k += kFastThreads) su[a * Npad + k] = k < gNpad ? to_fixed32(gx[a * gNpad + k], fscale) : 0u;
        uint8_t *ssp = smem_raw + F.sp;
        const uint8_t *gsp = A.sp + (size_t)c * gNpad;
        for (int k = tid; k < Npad; k += kFastThreads) ssp[k] = k < gNpad ? gsp[k] : 0;
        float *scp = (float *)(smem_raw + F.cp);
        for (int k = tid; k < ns * ns; k += kFastThreads) {
            const double *p = A.par + k * PMC_NPAR;
            scp[8 * k + 0] = (float)p[PMC_P_RCUT2];
            scp[8 * k + 1] = (float)p[PMC_P_EPS];
            scp[8 * k + 2] = (float)p[PMC_P_SIG2];
            scp[8 * k + 3] = (float)p[PMC_P_SHIFT];
            scp[8 * k + 4] = (float)p[5];  // C0 | ndiv2
            scp[8 * k + 5] = (float)p[6];  // C2 / sigma^2
            scp[8 * k + 6] = (float)p[7];  // C4 / sigma^4
            scp[8 * k + 7] = 0.0f;
        }
        unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
        if (tid < 2 * PMC_MAX_MOVES) scnt[tid] = 0ull;
        if (tid < PMC_MAX_SPECIES) {
            double rc2 = 0.0;
            for (int b = 0; b < ns; b++) rc2 = fmax(rc2, A.par[((tid < ns ? tid : 0) * ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            ((double *)(smem_raw + F.rcs))[tid] = sqrt(rc2);
        }
    }
    uint32_t myq[kFastCand];  // packed 8-bit prefilter coordinates of this thread's candidates (see pack8)
#pragma unroll
    for (int k = 0; k < kFastCand; k++) {
        const int j = k * kFastThreads + tid;
        uint32_t u[3] = {0u, 0u, 0u};
#pragma unroll
        for (int a = 0; a < DIM; a++) u[a] = j < gNpad ? to_fixed32(gx[a * gNpad + j], fscale) : 0u;
        myq[k] = pack8(u[0], u[1], u[2]);
    }
    const double Tk = A.temp[c];
    double E = A.energy[c];
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t gchain = (uint32_t)(A.chain_offset + c);
    int32_t *gimg = A.img + (size_t)c * DIM * gNpad;
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * kFastCand * 32);
    uint32_t slot = 0;

    for (long long tb = 0; tb < A.n_trials; tb += kFastBatch) {
        const int nb = (int)min((long long)kFastBatch, A.n_trials - tb);
        __syncthreads();
        if (tid < nb) {
            const long long q = tb + tid;
            pmc_trial tr;
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
                float z0, z1, z2, z3;
                box_muller(b.v[0], b.v[1], z0, z1);
                box_muller(b.v[2], b.v[3], z2, z3);
                const float sg = A.mv_sigma[m];
                tr.u = uniform53(a.v[2], a.v[3]);
                tr.move = m;
                tr.kind = PMC_MOVE_DISPLACEMENT;
                tr.i = (int)bounded(a.v[1], (uint32_t)N);
                tr.j = -1;
                tr.delta[0] = (double)(sg * z0);
                tr.delta[1] = (double)(sg * z1);
                tr.delta[2] = (DIM == 3) ? (double)(sg * z2) : 0.0;
                if (A.trace) A.trace[(size_t)c * A.n_trials + q] = tr;
            }
            unsigned char *rec = smem_raw + F.rec + (size_t)kRecBytes * tid;
            double *rd = (double *)rec;
            int *ri = (int *)(rec + 32);
            uint32_t *rt = (uint32_t *)(rec + 64);
            rd[3] = A.exact_exp ? tr.u : -Tk * log(tr.u);
            ri[0] = (int)__double2ll_rn(tr.delta[0] * fscale);
            ri[1] = (int)__double2ll_rn(tr.delta[1] * fscale);
            ri[2] = (int)__double2ll_rn(tr.delta[2] * fscale);
            ri[3] = tr.i;
            ri[4] = tr.move;
            const double hd = 0.5 * sqrt(tr.delta[0] * tr.delta[0] + tr.delta[1] * tr.delta[1] + tr.delta[2] * tr.delta[2]);
            const double *rcs = (const double *)(smem_raw + F.rcs);
#pragma unroll
            for (int s = 0; s < PMC_MAX_SPECIES; s++) rt[s] = neg_thr8((rcs[s] + hd) * fscale * 0x1p-24);
        }
        __syncthreads();

        for (int b = 0; b < nb; b++) {
            const uint32_t ra = sb + F.rec + (uint32_t)kRecBytes * (uint32_t)b;
            const double thr = lds_f64(ra + 24);
            int di0, di1, di2, i;
            lds_s32x4(ra + 32, di0, di1, di2, i);
            const uint32_t ua = sb + F.u + 4u * (uint32_t)i;
            const uint32_t uo0 = lds_u32(ua), uo1 = lds_u32(ua + nb4), uo2 = (DIM == 3) ? lds_u32(ua + 2 * nb4) : 0u;
            const uint32_t si = lds_u8(sb + F.sp + (uint32_t)i);
            const uint32_t un0 = uo0 + (uint32_t)di0, un1 = uo1 + (uint32_t)di1, un2 = uo2 + (uint32_t)di2;
            const uint32_t um0 = uo0 + (uint32_t)(di0 >> 1), um1 = uo1 + (uint32_t)(di1 >> 1), um2 = uo2 + (uint32_t)(di2 >> 1);
            const int fthr = (int)lds_u32(ra + 64 + 4u * si);
            const uint32_t umq = pack8(um0, um1, um2);
            uint32_t m8 = 0;
#pragma unroll
            for (int k = 0; k < kFastCand; k++) {  // survivor: bit kFastCand-1-k
                const uint32_t t = __vabsdiffu4(umq, myq[k]);
                m8 = __funnelshift_l((uint32_t)__dp4a((int)t, (int)t, fthr), m8, 1);
            }
            const int mine = __popc(m8);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                incl += (lane >= o) ? t : 0;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
            for (int k = 0; k < kFastCand; k++) {
                if (m8 & (1u << (kFastCand - 1 - k))) {
                    sts_u16(wp, (uint32_t)(k * kFastThreads + tid));
                    wp += 2;
                }
            }
            __syncwarp();
            float partf = 0.0f;
            const uint32_t prow = si * (uint32_t)ns;
            for (int q = lane; q < total; q += 32) {
                const uint32_t j = lds_u16(qa + 2u * (uint32_t)q);
                if (j < (uint32_t)N && j != (uint32_t)i) {
                    const uint32_t ja = sb + F.u + 4u * j;
                    const uint32_t a0 = lds_u32(ja), a1 = lds_u32(ja + nb4), a2 = (DIM == 3) ? lds_u32(ja + 2 * nb4) : 0u;
                    const float r2o = __uint2float_rn(dist2_u32<DIM>(uo0, uo1, uo2, a0, a1, a2)) * r2scale;
                    const float r2n = __uint2float_rn(dist2_u32<DIM>(un0, un1, un2, a0, a1, a2)) * r2scale;
                    const uint32_t sj = lds_u8(sb + F.sp + j);
                    float rc2, eps, sig2, shift, c0, c2, c4, pad_;
                    const uint32_t pa = sb + F.cp + 32u * (prow + sj);
                    lds_f32x4(pa, rc2, eps, sig2, shift);
                    lds_f32x4(pa + 16, c0, c2, c4, pad_);
                    const float eo = pair_potential_f32<MODEL>(r2o, eps, sig2, shift, c0, c2, c4);
                    const float en = pair_potential_f32<MODEL>(r2n, eps, sig2, shift, c0, c2, c4);
                    partf += (r2n <= rc2 ? en : 0.0f) - (r2o <= rc2 ? eo : 0.0f);
                }
            }
            __syncwarp();
            double part = warp_sum((double)partf);
            const uint32_t rda = sb + F.red + 32u * slot;
            if (lane == 0) sts_f64(rda + 8u * (uint32_t)warp, part);
            __syncthreads();
            double s0, s1, s2, s3;
            lds_f64x2(rda, s0, s1);
            lds_f64x2(rda + 16, s2, s3);
            const double dE = ((s0 + s1) + s2) + s3;
            slot ^= 1u;
            const bool acc = A.exact_exp ? accept_exact(dE, Tk, thr) : (dE < thr);
            if (acc) {
                sts_u32(ua, un0);
                sts_u32(ua + nb4, un1);
                if constexpr (DIM == 3) sts_u32(ua + 2 * nb4, un2);
                E += dE;
                if (tid == (i & (kFastThreads - 1))) {
                    const int ki = i >> 7;
                    const uint32_t f = pack8(un0, un1, un2);
#pragma unroll
                    for (int k = 0; k < kFastCand; k++) myq[k] = k == ki ? f : myq[k];
                }
                if (tid == 32) {  // image counters: the fixed-point add wrapped around the box
                    const int w0 = (di0 > 0 && un0 < uo0) - (di0 < 0 && un0 > uo0);
                    const int w1 = (di1 > 0 && un1 < uo1) - (di1 < 0 && un1 > uo1);
                    const int w2 = (di2 > 0 && un2 < uo2) - (di2 < 0 && un2 > uo2);
                    if (w0) atomicAdd(&gimg[i], w0);
                    if (w1) atomicAdd(&gimg[gNpad + i], w1);
                    if (DIM == 3 && w2) atomicAdd(&gimg[2 * gNpad + i], w2);
                }
            }
            if (tid == 64) {  // bookkeeping is spread over warps, as in k_chain_sweep_fast
                unsigned long long *scnt = (unsigned long long *)(smem_raw + F.cnt);
                const int m = (int)lds_u32(ra + 48);
                scnt[m] += 1ull;
                scnt[PMC_MAX_MOVES + m] += acc ? 1ull : 0ull;
                if (A.acc_out) A.acc_out[(size_t)c * A.n_trials + tb + b] = acc ? 1 : 0;
                if (A.dE_out) A.dE_out[(size_t)c * A.n_trials + tb + b] = dE;
            }
        }
    }
    __syncthreads();
    {
        // back to float64 at the centre of the fixed-point cell: re-quantising it gives the same integer again
        const uint32_t *su = (const uint32_t *)(smem_raw + F.u);
        const double inv = L * 0x1p-32;
        for (int a = 0; a < DIM; a++)
            for (int k = tid; k < gNpad; k += kFastThreads) gx[a * gNpad + k] = ((double)su[a * Npad + k] + 0.5) * inv;
        const unsigned long long *scnt = (const unsigned long long *)(smem_raw + F.cnt);
        if (tid == 0) A.energy[c] = E;
        if (tid < A.n_moves) {
            A.calls[(size_t)c * PMC_MAX_MOVES + tid] += scnt[tid];
            A.accepted[(size_t)c * PMC_MAX_MOVES + tid] += scnt[PMC_MAX_MOVES + tid];
        }
    }
}

}  // namespace fast
}  // namespace pmc
