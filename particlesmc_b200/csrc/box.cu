// box.cu -- PMC_MODE_BOX: a single large periodic box updated with checkerboard parallel sweeps.
//
// Reference counterparts
//   cell list build      src/neighbours.jl:236-270 (LinkedList ctor + build_neighbour_list!)  -> K1:
//                        bin -> prefix sum -> scatter -> per-cell canonical order (descending particle
//                        id, the order head-insertion produces, neighbours.jl:257-268) + gather to SoA
//   local / total energy src/atoms.jl:40-58, :81-88                                          -> K2
//   Metropolis trial     src/moves.jl:57-90 + src/utils.jl:8-10                               -> K5
// K5 has NO reference counterpart as an algorithm: the reference updates one particle at a time over
// the whole box.  Here the box is cut into cells of side >= rcut_max with an even cell count per axis;
// cells of one colour (2^d colours) do not interact, so each is advanced independently by one CTA for
// n_cell trials while its 3^d-cell neighbourhood is frozen.  Moves leaving the cell are rejected and the
// grid origin is shifted by a fresh random vector every sweep (Anderson et al., J. Comput. Phys. 254
// (2013) 27), which keeps detailed balance per sub-sweep and restores ergodicity.  Parity with the
// reference is therefore statistical for trajectories and exact (1e-12) for energies.
#include "box.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "rng.cuh"

namespace pmc {

namespace {

thread_local std::string g_box_err;

int bfail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_box_err = buf;
    return code;
}

#define BCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return bfail(PMC_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int PMC_MAX_PEERS = 7;  // other GPUs of one 8-GPU box

struct Geom {
    int nc[3];
    int ncell;
    double L[3];
    double cs[3];     // cell side
    double shift[3];  // grid origin of this sweep, in [0, cs)
};

struct BoxArgs {
    Geom g;
    int N, ns, cap;
    // canonical state (particle order)
    double *x;     // [dim][N] wrapped
    int32_t *img;  // [dim][N]
    uint8_t *sp;   // [N]
    // cell-sorted state of the current grid
    double *xs;       // [dim][N]
    uint8_t *sps;     // [N]
    int32_t *ids;     // [N] particle id of each sorted slot
    int32_t *start;   // [ncell+1]
    const double *par;
    double T;
    float sigma;
    unsigned long long seed;
    uint32_t sweep;
    // per-cell outputs
    double *cellE;            // [ncell] sum of accepted dE (sweep) or sum of local energies (energy)
    uint32_t *cell_acc;       // [ncell]
    double *eloc;             // [N] (energy kernel)
    int *overflow;
    // multi-GPU (replicated state): this rank sweeps active cells [cta_offset, cta_offset + gridDim.x) of the colour
    // and pushes every accepted move into the peers' replicas with plain stores over NVLink peer memory
    int cta_offset;
    int n_peers;
    double *peer_x[PMC_MAX_PEERS];
    double *peer_xs[PMC_MAX_PEERS];
    int32_t *peer_img[PMC_MAX_PEERS];
    double *peer_cellE[PMC_MAX_PEERS];
    uint32_t *peer_cacc[PMC_MAX_PEERS];
};

// Cell coordinate and in-cell coordinate of a wrapped position under grid origin s.
__device__ __forceinline__ int cell_of(double x, double s, double L, double cs, int n) {
    double y = x - s;
    if (y < 0.0) y += L;
    int c = (int)(y / cs);
    return c >= n ? n - 1 : c;
}
__device__ __forceinline__ double in_frame(double x, double s, double L, double cs, int c) {
    double y = x - s;
    if (y < 0.0) y += L;
    return y - (double)c * cs;
}

template <int DIM>
__device__ __forceinline__ int lin_cell(const int (&c)[3], const int (&nc)[3]) {  // last axis fastest (neighbours.jl:79-88)
    int l = c[0];
    l = l * nc[1] + c[1];
    if constexpr (DIM == 3) l = l * nc[2] + c[2];
    return l;
}

// ---- K0: ingest / egress -------------------------------------------------------------------------
__global__ void k_box_ingest(const double *__restrict__ raw, const long long *__restrict__ rsp, int N, int dim, int ns,
                             Geom g, double *x, int32_t *img, uint8_t *sp, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int a = 0; a < dim; a++) {
        const double L = g.L[a], v = raw[(size_t)i * dim + a];
        const double n = floor(v / L);
        double w = v - n * L;
        int im = (int)n;
        if (w >= L) { w -= L; im += 1; }
        if (w < 0.0) { w += L; im -= 1; }
        if (!(w >= 0.0 && w <= L)) atomicExch(bad, 1);
        x[(size_t)a * N + i] = w;
        img[(size_t)a * N + i] = im;
    }
    const long long lab = rsp[i];
    if (lab < 1 || lab > ns) atomicExch(bad, 2);
    sp[i] = (uint8_t)(lab - 1);
}

__global__ void k_box_egress(const double *__restrict__ x, const int32_t *__restrict__ img, const uint8_t *__restrict__ sp,
                             int N, int dim, Geom g, double *raw, long long *rsp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int a = 0; a < dim; a++) raw[(size_t)i * dim + a] = x[(size_t)a * N + i] + (double)img[(size_t)a * N + i] * g.L[a];
    rsp[i] = (long long)sp[i] + 1;
}

// ---- K1: cell list by sorting ----------------------------------------------------------------------
template <int DIM>
__global__ void k_box_count(const double *__restrict__ x, int N, Geom g, int32_t *cid, int32_t *count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < DIM; a++) c[a] = cell_of(x[(size_t)a * N + i], g.shift[a], g.L[a], g.cs[a], g.nc[a]);
    const int l = lin_cell<DIM>(c, g.nc);
    cid[i] = l;
    atomicAdd(&count[l], 1);
}

// exclusive prefix sum of count[0..n) into start[0..n] in two small launches:
//   k_box_scan_local : every CTA scans its 1024 coalesced entries, writes the CTA-local exclusive prefix and its total
//   k_box_scan_apply : every CTA adds the sum of the totals of the CTAs before it (<= a few dozen values)
__global__ void k_box_scan_local(const int32_t *__restrict__ count, int32_t *start, int32_t *blocksum, int n) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = blockIdx.x * blockDim.x + tid;
    const int v = k < n ? count[k] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    if (k < n) start[k] = (warp ? s_warp[warp - 1] : 0) + incl - v;
    if (tid == 0) blocksum[blockIdx.x] = s_warp[31];
}

__global__ void k_box_scan_apply(int32_t *start, int32_t *cursor, const int32_t *__restrict__ blocksum, int n) {
    __shared__ int s_off, s_tot;
    if (threadIdx.x < 32) {
        int off = 0, tot = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) {
            const int v = blocksum[b];
            tot += v;
            if (b < (int)blockIdx.x) off += v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        }
        if (threadIdx.x == 0) {
            s_off = off;
            s_tot = tot;
        }
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        start[k] += s_off;
        cursor[k] = 0;
    }
    if (k == 0) start[n] = s_tot;
}

__global__ void k_box_scatter(const int32_t *__restrict__ cid, const int32_t *__restrict__ start, int32_t *cursor, int N,
                              int32_t *ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = cid[i];
    ids[start[c] + atomicAdd(&cursor[c], 1)] = i;
}

// canonical order inside each cell (descending particle id) + gather into the sorted SoA: one warp per cell,
// rank sort through shuffles (cells hold ~20 particles); cells with more than 32 particles fall back to a serial
// insertion sort by lane 0
template <int DIM>
__global__ void k_box_finalize(const int32_t *__restrict__ start, int32_t *ids, int ncell, int N,
                               const double *__restrict__ x, const uint8_t *__restrict__ sp, double *xs, uint8_t *sps) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncell) return;
    const int b = start[c], e = start[c + 1], cnt = e - b;
    if (cnt <= 32) {
        const int mine = lane < cnt ? ids[b + lane] : -1;
        int rank = 0;
        for (int k = 0; k < cnt; k++) {
            const int v = __shfl_sync(0xffffffffu, mine, k);
            rank += (v > mine) ? 1 : 0;
        }
        __syncwarp();
        if (lane < cnt) {
            const int p = b + rank;
            ids[p] = mine;
#pragma unroll
            for (int a = 0; a < DIM; a++) xs[(size_t)a * N + p] = x[(size_t)a * N + mine];
            sps[p] = sp[mine];
        }
    } else {
        if (lane == 0) {
            for (int p = b + 1; p < e; p++) {  // insertion sort, descending
                const int v = ids[p];
                int q = p - 1;
                while (q >= b && ids[q] < v) {
                    ids[q + 1] = ids[q];
                    q--;
                }
                ids[q + 1] = v;
            }
        }
        __syncwarp();
        for (int p = b + lane; p < e; p += 32) {
            const int i = ids[p];
#pragma unroll
            for (int a = 0; a < DIM; a++) xs[(size_t)a * N + p] = x[(size_t)a * N + i];
            sps[p] = sp[i];
        }
    }
}

// ---- stencil loader shared by K2 and K5 ---------------------------------------------------------------
// Gathers the particles of the 3^d cells around `cc` into shared memory in the frame of the central
// cell (its particles in [0, cs)^d, neighbours shifted by whole cells), so no per-pair minimum image is
// needed.  The central cell comes first: candidates [0, ncen) are the movable particles.
template <int DIM>
struct Stencil {
    static constexpr int NST = DIM == 3 ? 27 : 9;
    int cell[NST];
    int off[NST + 1];  // candidate offsets of the stencil cells in the gathered list
    int cnt[NST];      // particles in each stencil cell
    int base[NST];     // first sorted slot of each stencil cell
    int cwrap[NST][3];
    int o[NST][3];
};

template <int DIM>
__device__ int load_stencil(const BoxArgs &A, const int (&cc)[3], Stencil<DIM> *st, double *sr, uint8_t *ssp) {
    constexpr int NST = Stencil<DIM>::NST;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (tid < NST) {
        // slot 0 = central cell; the others in first-axis-fastest order (Iterators.product, neighbours.jl:101)
        int k = tid == 0 ? (NST / 2) : (tid <= NST / 2 ? tid - 1 : tid);
        int c[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const int oa = k % 3 - 1;
            k /= 3;
            int v = cc[a] + oa;
            if (v < 0) v += A.g.nc[a];
            if (v >= A.g.nc[a]) v -= A.g.nc[a];
            c[a] = v;
            st->o[tid][a] = oa;
            st->cwrap[tid][a] = v;
        }
        const int l = lin_cell<DIM>(c, A.g.nc);
        st->cell[tid] = l;
        st->off[tid + 1] = A.start[l + 1] - A.start[l];
    }
    __syncthreads();
    if (tid == 0) {
        st->off[0] = 0;
        for (int k = 0; k < NST; k++) st->off[k + 1] += st->off[k];
    }
    __syncthreads();
    const int ncand = st->off[NST];
    if (ncand > A.cap) {
        if (tid == 0) atomicExch(A.overflow, 1);
        return -1;
    }
    for (int s = warp; s < NST; s += nwarp) {
        const int b = A.start[st->cell[s]], n = st->off[s + 1] - st->off[s], dst = st->off[s];
        for (int p = lane; p < n; p += 32) {
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                const double r = in_frame(A.xs[(size_t)a * A.N + b + p], A.g.shift[a], A.g.L[a], A.g.cs[a], st->cwrap[s][a]);
                sr[a * A.cap + dst + p] = r + (double)st->o[s][a] * A.g.cs[a];
            }
            ssp[dst + p] = A.sps[b + p];
        }
    }
    __syncthreads();
    return ncand;
}

// Flat variant used by the fast sweep kernel: the candidate index space of the whole stencil is spread over ALL
// threads (thread t handles candidates t, t + NT, ...), so the global loads of one CTA are independent and in
// flight together instead of one dependent start[] -> xs[] chain per stencil cell.  Same candidate order as
// load_stencil.  (A variant that also dropped candidates farther than rc from the central cell -- 24 % of the
// stencil -- was measured slower: the second pass over the stencil costs more than the smaller scan saves.)
template <int DIM, int NT>
__device__ int load_stencil_flat(const BoxArgs &A, const int (&cc)[3], Stencil<DIM> *st, double *sr, uint8_t *ssp) {
    constexpr int NST = Stencil<DIM>::NST;
    const int tid = threadIdx.x;
    if (tid < NST) {
        int k = tid == 0 ? (NST / 2) : (tid <= NST / 2 ? tid - 1 : tid);
        int c[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const int oa = k % 3 - 1;
            k /= 3;
            int v = cc[a] + oa;
            if (v < 0) v += A.g.nc[a];
            if (v >= A.g.nc[a]) v -= A.g.nc[a];
            c[a] = v;
            st->o[tid][a] = oa;
            st->cwrap[tid][a] = v;
        }
        const int l = lin_cell<DIM>(c, A.g.nc);
        const int b = A.start[l];
        st->cell[tid] = l;
        st->base[tid] = b;
        st->cnt[tid] = A.start[l + 1] - b;
    }
    __syncthreads();
    if (tid == 0) {
        st->off[0] = 0;
        for (int k = 0; k < NST; k++) st->off[k + 1] = st->off[k] + st->cnt[k];
    }
    __syncthreads();
    const int nall = st->off[NST];
    if (nall > A.cap) {
        if (tid == 0) atomicExch(A.overflow, 1);
        return -1;
    }
    for (int t = tid; t < nall; t += NT) {
        int lo = 0, hi = NST;  // stencil cell of flat index t: binary search over the 3^d offsets
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (t >= st->off[mid]) lo = mid; else hi = mid;
        }
        const int slot = st->base[lo] + (t - st->off[lo]);
#pragma unroll
        for (int a = 0; a < DIM; a++)
            sr[a * A.cap + t] = in_frame(A.xs[(size_t)a * A.N + slot], A.g.shift[a], A.g.L[a], A.g.cs[a], st->cwrap[lo][a]) +
                                (double)st->o[lo][a] * A.g.cs[a];
        ssp[t] = A.sps[slot];
    }
    __syncthreads();
    return nall;
}

template <int DIM>
__device__ __forceinline__ double d2_frame(const double *__restrict__ sr, int cap, int j, const double (&xi)[3]) {
    double d = xi[0] - sr[j];
    double r2 = d * d;
    d = xi[1] - sr[cap + j];
    r2 = fma(d, d, r2);
    if constexpr (DIM == 3) {
        d = xi[2] - sr[2 * cap + j];
        r2 = fma(d, d, r2);
    }
    return r2;
}

constexpr int kBoxThreads = 128;
constexpr int kBoxWarps = kBoxThreads / 32;

// ---- K2: local energies of all particles, per-cell sums --------------------------------------------
template <int DIM, int MODEL>
__global__ void __launch_bounds__(kBoxThreads) k_box_energy(const __grid_constant__ BoxArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Stencil<DIM> st;
    __shared__ double s_par[PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR];
    __shared__ double s_w[kBoxWarps];
    double *sr = (double *)smem_raw;
    uint8_t *ssp = (uint8_t *)(sr + DIM * A.cap);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < A.ns * A.ns * PMC_NPAR; k += kBoxThreads) s_par[k] = A.par[k];
    int cc[3] = {0, 0, 0};
    {
        int l = blockIdx.x;
        if constexpr (DIM == 3) { cc[2] = l % A.g.nc[2]; l /= A.g.nc[2]; }
        cc[1] = l % A.g.nc[1];
        cc[0] = l / A.g.nc[1];
    }
    const int ncand = load_stencil<DIM>(A, cc, &st, sr, ssp);
    if (ncand < 0) return;
    const int ncen = st.off[1], b = A.start[st.cell[0]];
    double wsum = 0.0;
    for (int k = warp; k < ncen; k += kBoxWarps) {
        const double xi[3] = {sr[k], sr[A.cap + k], DIM == 3 ? sr[2 * A.cap + k] : 0.0};
        const double *prow = s_par + ssp[k] * A.ns * PMC_NPAR;
        double e = 0.0;
        for (int j = lane; j < ncand; j += 32) {
            if (j == k) continue;
            const double *p = prow + ssp[j] * PMC_NPAR;
            const double r2 = d2_frame<DIM>(sr, A.cap, j, xi);
            if (r2 <= p[PMC_P_RCUT2]) e += pair_potential<MODEL>(p, r2);
        }
        e = warp_sum(e);
        if (lane == 0) A.eloc[A.ids[b + k]] = e;
        wsum += e;
    }
    if (lane == 0) s_w[warp] = wsum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kBoxWarps; w++) t += s_w[w];
        A.cellE[blockIdx.x] = t;
    }
}

// ---- K5: checkerboard sweep of one colour --------------------------------------------------------------
template <int DIM, int MODEL>
__global__ void __launch_bounds__(kBoxThreads) k_box_sweep(const __grid_constant__ BoxArgs A, int colour) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Stencil<DIM> st;
    __shared__ double s_par[PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR];
    __shared__ double s_red[2][kBoxWarps];
    __shared__ double s_delta[kBoxThreads][3];
    __shared__ double s_thr[kBoxThreads];
    __shared__ int s_k[kBoxThreads];
    double *sr = (double *)smem_raw;
    uint8_t *ssp = (uint8_t *)(sr + DIM * A.cap);
    uint8_t *moved = ssp + A.cap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < A.ns * A.ns * PMC_NPAR; k += kBoxThreads) s_par[k] = A.par[k];
    // active cell of this CTA: coordinates 2*h + colour bit
    int cc[3] = {0, 0, 0};
    {
        int l = blockIdx.x + A.cta_offset;
        if constexpr (DIM == 3) { cc[2] = 2 * (l % (A.g.nc[2] / 2)) + ((colour >> 2) & 1); l /= (A.g.nc[2] / 2); }
        cc[1] = 2 * (l % (A.g.nc[1] / 2)) + ((colour >> 1) & 1);
        cc[0] = 2 * (l / (A.g.nc[1] / 2)) + (colour & 1);
    }
    const int ncand = load_stencil<DIM>(A, cc, &st, sr, ssp);
    if (ncand < 0) return;
    const int cell = st.cell[0], ncen = st.off[1], b = A.start[cell];
    for (int k = tid; k < ncen; k += kBoxThreads) moved[k] = 0;
    const double cs[3] = {A.g.cs[0], A.g.cs[1], A.g.cs[2]};
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    double Esum = 0.0;
    uint32_t nacc = 0;
    int last_k = -1, slot = 0;
    double last_x[3] = {0.0, 0.0, 0.0};

    for (int tb = 0; tb < ncen; tb += kBoxThreads) {  // n_cell trials in this cell (one sweep = N trials)
        const int nb = min(kBoxThreads, ncen - tb);
        __syncthreads();
        if (tid < nb) {
            const uint32_t q = (uint32_t)(tb + tid);
            const Philox4 a = philox4x32_10(q, (uint32_t)cell, A.sweep, 0u, k0, k1);
            const Philox4 bb = philox4x32_10(q, (uint32_t)cell, A.sweep, 1u, k0, k1);
            float z0, z1, z2, z3;
            box_muller(bb.v[0], bb.v[1], z0, z1);
            box_muller(bb.v[2], bb.v[3], z2, z3);
            s_k[tid] = (int)bounded(a.v[1], (uint32_t)ncen);
            s_delta[tid][0] = (double)(A.sigma * z0);
            s_delta[tid][1] = (double)(A.sigma * z1);
            s_delta[tid][2] = (double)(A.sigma * z2);
            s_thr[tid] = -A.T * log(uniform53(a.v[2], a.v[3]));
        }
        __syncthreads();
        for (int t = 0; t < nb; t++) {
            const int k = s_k[t];
            double xo[3] = {0.0, 0.0, 0.0}, xn[3] = {0.0, 0.0, 0.0};
            bool inside = true;
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                xo[a] = (k == last_k) ? last_x[a] : sr[a * A.cap + k];
                xn[a] = xo[a] + s_delta[t][a];
                inside &= (xn[a] >= 0.0) && (xn[a] < cs[a]);
            }
            if (!inside) continue;  // leaves the cell: rejected (uniform across the CTA)
            const double *prow = s_par + ssp[k] * A.ns * PMC_NPAR;
            double part = 0.0;
            for (int j = tid; j < ncand; j += kBoxThreads) {
                if (j == k) continue;
                const double *p = prow + ssp[j] * PMC_NPAR;
                const double rc2 = p[PMC_P_RCUT2];
                const double r2o = d2_frame<DIM>(sr, A.cap, j, xo);
                const double r2n = d2_frame<DIM>(sr, A.cap, j, xn);
                if (r2o <= rc2) part -= pair_potential<MODEL>(p, r2o);
                if (r2n <= rc2) part += pair_potential<MODEL>(p, r2n);
            }
            part = warp_sum(part);
            if (lane == 0) s_red[slot][warp] = part;
            __syncthreads();
            double dE = s_red[slot][0];
#pragma unroll
            for (int w = 1; w < kBoxWarps; w++) dE += s_red[slot][w];
            slot ^= 1;
            if (dE < s_thr[t]) {
                if (tid == k % kBoxThreads) {
#pragma unroll
                    for (int a = 0; a < DIM; a++) sr[a * A.cap + k] = xn[a];
                    moved[k] = 1;
                }
                last_k = k;
#pragma unroll
                for (int a = 0; a < DIM; a++) last_x[a] = xn[a];
                Esum += dE;
                nacc++;
            }
        }
    }
    __syncthreads();
    // write moved particles back: canonical arrays (+ image counters) and the sorted copy
    for (int k = tid; k < ncen; k += kBoxThreads) {
        if (!moved[k]) continue;
        const int i = A.ids[b + k];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const double xold = A.xs[(size_t)a * A.N + b + k];
            const double r0 = in_frame(xold, A.g.shift[a], A.g.L[a], A.g.cs[a], cc[a]);
            int w;
            const double xnew = wrap1(xold + (sr[a * A.cap + k] - r0), A.g.L[a], w);
            const int im = A.img[(size_t)a * A.N + i] + w;
            A.xs[(size_t)a * A.N + b + k] = xnew;
            A.x[(size_t)a * A.N + i] = xnew;
            if (w) A.img[(size_t)a * A.N + i] = im;
            for (int p = 0; p < A.n_peers; p++) {
                A.peer_xs[p][(size_t)a * A.N + b + k] = xnew;
                A.peer_x[p][(size_t)a * A.N + i] = xnew;
                if (w) A.peer_img[p][(size_t)a * A.N + i] = im;
            }
        }
    }
    if (tid == 0) {
        A.cellE[cell] = Esum;
        A.cell_acc[cell] = nacc;
        for (int p = 0; p < A.n_peers; p++) {
            A.peer_cellE[p][cell] = Esum;
            A.peer_cacc[p][cell] = nacc;
        }
    }
    if (A.n_peers) __threadfence_system();
}

// ---- K5 (fast): checkerboard sweep with the integer prefilter ---------------------------------------------
// Same trials and the same fp64 pair terms as k_box_sweep, issued the way chains_fast.cuh does it: 128 threads
// per active cell, every thread keeps the packed 8-bit frame coordinates of its KC candidates in registers, one
// sphere test (midpoint of old/new, radius rc + |delta|/2; VABSDIFF4 + IDP.4A + funnel shift) per candidate, survivors
// compacted with one warp prefix sum, fp64 only for survivors (no minimum image in the cell frame), explicit
// 32-bit shared addressing.  Needs cubic cells (one fixed-point scale); otherwise k_box_sweep is used.
// The one-warp-per-trial speculative scheme of chains_spec.cuh was tried here too and measured 17 % SLOWER
// (5.3e8 vs 6.3e8 moves/s at N = 2^20): a cell holds only ~19 trials, so the rounds of four never fill up and
// every warp has to stream the whole stencil.
constexpr int kBfThreads = 128;
constexpr int kBfWarps = kBfThreads / 32;
// register candidates per thread offered (capacity = threads x KC): 512 / 640 / 768 / 1024 candidates
constexpr int kBfKc0 = 512 / kBfThreads, kBfKc1 = 640 / kBfThreads, kBfKc2 = 768 / kBfThreads, kBfKc3 = 1024 / kBfThreads;
constexpr int kBfBatch = 32;
constexpr int kBfRec = 80;

struct BfLayout {
    uint32_t r, sp, mv, q, cp, rec, red, par, rcs, total;
};
__host__ __device__ inline BfLayout bf_layout(int dim, int cap) {
    BfLayout f;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        uint32_t p = o;
        o += (bytes + 15u) & ~15u;
        return p;
    };
    f.r = take(8u * dim * cap);
    f.sp = take(cap);
    f.mv = take(cap);
    f.q = take(2u * cap);
    f.cp = take(32u * PMC_MAX_SPECIES * PMC_MAX_SPECIES);
    f.rec = take((uint32_t)kBfRec * kBfBatch);
    f.red = take(8u * 2 * kBfWarps);
    f.par = take(8u * PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR);
    f.rcs = take(8u * PMC_MAX_SPECIES);
    f.total = o;
    return f;
}

__device__ __forceinline__ double bf_lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void bf_lds_f64x2(uint32_t a, double &v0, double &v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a) : "memory");
}
__device__ __forceinline__ void bf_lds_s32x4(uint32_t a, int &v0, int &v1, int &v2, int &v3) {
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t bf_lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t bf_lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t bf_lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void bf_sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void bf_sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void bf_sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// frame coordinate r in [-cs, 2cs) -> fixed point in [0, 3 * 2^30); its top byte (pack8) counts units of cs / 64
// in [0, 192).  The moved particle and its trial position lie in the centre cell [64, 128), so no byte difference
// reaches 128 and the signed-byte reading of VABSDIFF4 (common.cuh) never aliases in the frame.
__device__ __forceinline__ uint32_t bf_fixed(double r, double cs, double scale) { return (uint32_t)__double2ull_rd((r + cs) * scale); }

template <int DIM, int MODEL, int KC>
__global__ void __launch_bounds__(kBfThreads, kBfThreads == 64 ? 12 : 8) k_box_sweep_fast(const __grid_constant__ BoxArgs A, int colour) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Stencil<DIM> st;
    constexpr int CAP = kBfThreads * KC;
    const BfLayout F = bf_layout(DIM, CAP);
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    constexpr uint32_t cap8 = 8u * CAP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *sr = (double *)(smem_raw + F.r);
    uint8_t *ssp = smem_raw + F.sp;
    {
        double *spar = (double *)(smem_raw + F.par), *scp = (double *)(smem_raw + F.cp);
        for (int k = tid; k < A.ns * A.ns * PMC_NPAR; k += kBfThreads) spar[k] = A.par[k];
        for (int k = tid; k < A.ns * A.ns; k += kBfThreads) {
            scp[4 * k + 0] = A.par[k * PMC_NPAR + PMC_P_RCUT2];
            scp[4 * k + 1] = A.par[k * PMC_NPAR + PMC_P_EPS];
            scp[4 * k + 2] = A.par[k * PMC_NPAR + PMC_P_SIG2];
            scp[4 * k + 3] = A.par[k * PMC_NPAR + PMC_P_SHIFT];
        }
        if (tid < PMC_MAX_SPECIES) {
            double rc2 = 0.0;
            for (int b = 0; b < A.ns; b++) rc2 = fmax(rc2, A.par[((tid < A.ns ? tid : 0) * A.ns + b) * PMC_NPAR + PMC_P_RCUT2]);
            ((double *)(smem_raw + F.rcs))[tid] = sqrt(rc2);
        }
    }
    int cc[3] = {0, 0, 0};
    {
        int l = blockIdx.x + A.cta_offset;
        if constexpr (DIM == 3) { cc[2] = 2 * (l % (A.g.nc[2] / 2)) + ((colour >> 2) & 1); l /= (A.g.nc[2] / 2); }
        cc[1] = 2 * (l % (A.g.nc[1] / 2)) + ((colour >> 1) & 1);
        cc[0] = 2 * (l / (A.g.nc[1] / 2)) + (colour & 1);
    }
    const int ncand = load_stencil_flat<DIM, kBfThreads>(A, cc, &st, sr, ssp);  // A.cap == CAP on this path
    if (ncand < 0) return;
    const int cell = st.cell[0], ncen = st.off[1], bstart = A.start[cell];
    for (int k = tid; k < ncen; k += kBfThreads) smem_raw[F.mv + k] = 0;
    const double cs = A.g.cs[0];
    const double fscale = 1073741824.0 / cs;  // 2^30 / cell side
    uint32_t myq[KC];                          // packed 8-bit prefilter coordinates of this thread's candidates
#pragma unroll
    for (int k = 0; k < KC; k++) {
        const int j = k * kBfThreads + tid;
        uint32_t u[3] = {0u, 0u, 0u};
#pragma unroll
        for (int a = 0; a < DIM; a++) u[a] = j < ncand ? bf_fixed(sr[a * CAP + j], cs, fscale) : 0u;
        myq[k] = pack8(u[0], u[1], u[2]);
    }
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * KC * 32);
    double Esum = 0.0;
    uint32_t nacc = 0, slot = 0;

    for (int tb = 0; tb < ncen; tb += kBfBatch) {
        const int nb = min(kBfBatch, ncen - tb);
        __syncthreads();
        if (tid < nb) {
            const uint32_t q = (uint32_t)(tb + tid);
            const Philox4 a = philox4x32_10(q, (uint32_t)cell, A.sweep, 0u, k0, k1);
            const Philox4 bb = philox4x32_10(q, (uint32_t)cell, A.sweep, 1u, k0, k1);
            float z0, z1, z2, z3;
            box_muller(bb.v[0], bb.v[1], z0, z1);
            box_muller(bb.v[2], bb.v[3], z2, z3);
            unsigned char *rec = smem_raw + F.rec + (size_t)kBfRec * tid;
            double *rd = (double *)rec;
            int *ri = (int *)(rec + 32);
            uint32_t *rt = (uint32_t *)(rec + 64);
            const double dx = (double)(A.sigma * z0), dy = (double)(A.sigma * z1), dz = DIM == 3 ? (double)(A.sigma * z2) : 0.0;
            rd[0] = dx;
            rd[1] = dy;
            rd[2] = dz;
            rd[3] = -A.T * log(uniform53(a.v[2], a.v[3]));
            ri[0] = (int)__double2ll_rn(dx * fscale);
            ri[1] = (int)__double2ll_rn(dy * fscale);
            ri[2] = (int)__double2ll_rn(dz * fscale);
            ri[3] = (int)bounded(a.v[1], (uint32_t)ncen);
            const double hd = 0.5 * sqrt(dx * dx + dy * dy + dz * dz);
            const double *rcs = (const double *)(smem_raw + F.rcs);
#pragma unroll
            for (int s = 0; s < PMC_MAX_SPECIES; s++) rt[s] = neg_thr8((rcs[s] + hd) * fscale * 0x1p-24);
        }
        __syncthreads();
        for (int t = 0; t < nb; t++) {
            const uint32_t ra = sb + F.rec + (uint32_t)kBfRec * (uint32_t)t;
            double d0, d1, d2, thr;
            int di0, di1, di2, k;
            bf_lds_f64x2(ra, d0, d1);
            bf_lds_f64x2(ra + 16, d2, thr);
            bf_lds_s32x4(ra + 32, di0, di1, di2, k);
            const uint32_t xa = sb + F.r + 8u * (uint32_t)k;
            double xo[3], xn[3];
            xo[0] = bf_lds_f64(xa);
            xo[1] = bf_lds_f64(xa + cap8);
            xo[2] = DIM == 3 ? bf_lds_f64(xa + 2 * cap8) : 0.0;
            xn[0] = xo[0] + d0;
            xn[1] = xo[1] + d1;
            xn[2] = xo[2] + d2;
            bool inside = xn[0] >= 0.0 && xn[0] < cs && xn[1] >= 0.0 && xn[1] < cs;
            if constexpr (DIM == 3) inside = inside && xn[2] >= 0.0 && xn[2] < cs;
            if (!inside) continue;  // leaves the cell: rejected (uniform across the CTA)
            const uint32_t si = bf_lds_u8(sb + F.sp + (uint32_t)k);
            const uint32_t um0 = bf_fixed(xo[0], cs, fscale) + (uint32_t)(di0 >> 1);
            const uint32_t um1 = bf_fixed(xo[1], cs, fscale) + (uint32_t)(di1 >> 1);
            const uint32_t um2 = DIM == 3 ? bf_fixed(xo[2], cs, fscale) + (uint32_t)(di2 >> 1) : 0u;
            const int fthr = (int)bf_lds_u32(ra + 64 + 4u * si);
            const uint32_t umq = pack8(um0, um1, um2);
            uint32_t m = 0;
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {  // survivor: bit KC-1-kk
                const uint32_t v = __vabsdiffu4(umq, myq[kk]);
                m = __funnelshift_l((uint32_t)__dp4a((int)v, (int)v, fthr), m, 1);
            }
            const int mine = __popc(m);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                incl += (lane >= o) ? v : 0;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                if (m & (1u << (KC - 1 - kk))) {
                    bf_sts_u16(wp, (uint32_t)(kk * kBfThreads + tid));
                    wp += 2;
                }
            }
            __syncwarp();
            double part = 0.0;
            const uint32_t prow = si * (uint32_t)A.ns;
            for (int q = lane; q < total; q += 32) {
                const uint32_t j = bf_lds_u16(qa + 2u * (uint32_t)q);
                if (j < (uint32_t)ncand && j != (uint32_t)k) {
                    const uint32_t ja = sb + F.r + 8u * j;
                    const double x0 = bf_lds_f64(ja), x1 = bf_lds_f64(ja + cap8);
                    double a_ = xo[0] - x0, b_ = xn[0] - x0;
                    double r2o = a_ * a_, r2n = b_ * b_;
                    a_ = xo[1] - x1;
                    b_ = xn[1] - x1;
                    r2o = fma(a_, a_, r2o);
                    r2n = fma(b_, b_, r2n);
                    if constexpr (DIM == 3) {
                        const double x2 = bf_lds_f64(ja + 2 * cap8);
                        a_ = xo[2] - x2;
                        b_ = xn[2] - x2;
                        r2o = fma(a_, a_, r2o);
                        r2n = fma(b_, b_, r2n);
                    }
                    const uint32_t sj = bf_lds_u8(sb + F.sp + j);
                    if constexpr (MODEL == PMC_MODEL_LJ || MODEL == PMC_MODEL_KG) {
                        double rc2, eps4, sig2, shift;
                        const uint32_t pa = sb + F.cp + 32u * (prow + sj);
                        bf_lds_f64x2(pa, rc2, eps4);
                        bf_lds_f64x2(pa + 16, sig2, shift);
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
            part = warp_sum(part);
            const uint32_t rda = sb + F.red + 8u * kBfWarps * slot;
            if (lane == 0) bf_sts_f64(rda + 8u * (uint32_t)warp, part);
            __syncthreads();
            double s0, s1, s2 = 0.0, s3 = 0.0;
            bf_lds_f64x2(rda, s0, s1);
            if constexpr (kBfWarps == 4) bf_lds_f64x2(rda + 16, s2, s3);
            const double dE = kBfWarps == 4 ? ((s0 + s1) + s2) + s3 : s0 + s1;
            slot ^= 1u;
            if (dE < thr) {
                bf_sts_f64(xa, xn[0]);
                bf_sts_f64(xa + cap8, xn[1]);
                if constexpr (DIM == 3) bf_sts_f64(xa + 2 * cap8, xn[2]);
                bf_sts_u8(sb + F.mv + (uint32_t)k, 1u);
                Esum += dE;
                nacc++;
                if (tid == (k & (kBfThreads - 1))) {  // owner refreshes its register copy
                    const int ki = k / kBfThreads;
                    const uint32_t f = pack8(bf_fixed(xn[0], cs, fscale), bf_fixed(xn[1], cs, fscale), DIM == 3 ? bf_fixed(xn[2], cs, fscale) : 0u);
#pragma unroll
                    for (int kk = 0; kk < KC; kk++) myq[kk] = kk == ki ? f : myq[kk];
                }
            }
        }
    }
    __syncthreads();
    // write moved particles back: canonical arrays (+ image counters) and the sorted copy
    for (int k = tid; k < ncen; k += kBfThreads) {
        if (!smem_raw[F.mv + k]) continue;
        const int i = A.ids[bstart + k];
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const double xold = A.xs[(size_t)a * A.N + bstart + k];
            const double r0 = in_frame(xold, A.g.shift[a], A.g.L[a], A.g.cs[a], cc[a]);
            int w;
            const double xnew = wrap1(xold + (sr[a * CAP + k] - r0), A.g.L[a], w);
            const int im = A.img[(size_t)a * A.N + i] + w;
            A.xs[(size_t)a * A.N + bstart + k] = xnew;
            A.x[(size_t)a * A.N + i] = xnew;
            if (w) A.img[(size_t)a * A.N + i] = im;
            for (int p = 0; p < A.n_peers; p++) {
                A.peer_xs[p][(size_t)a * A.N + bstart + k] = xnew;
                A.peer_x[p][(size_t)a * A.N + i] = xnew;
                if (w) A.peer_img[p][(size_t)a * A.N + i] = im;
            }
        }
    }
    if (tid == 0) {
        A.cellE[cell] = Esum;
        A.cell_acc[cell] = nacc;
        for (int p = 0; p < A.n_peers; p++) {
            A.peer_cellE[p][cell] = Esum;
            A.peer_cacc[p][cell] = nacc;
        }
    }
    if (A.n_peers) __threadfence_system();
}

// (Fusing this barrier into the colour kernels -- last CTA signals, every CTA of the next colour waits -- was tried and
// measured 6 % slower on 2 GPUs: the system-scope fence + counter per CTA costs more than the launch it saves.)
// Inter-GPU barrier between colour phases: every rank stores `epoch` into its slot of every peer's flag array
// (NVLink peer store, after a system fence so the sweep kernel's pushes are visible first), then spins on its own
// flag array until all ranks have arrived.  Bounded spin: a peer that never arrives raises an error flag
// instead of hanging the GPU.
__global__ void k_peer_barrier(volatile uint32_t *local_flags, uint32_t *const *peer_flags, int rank, int world, uint32_t epoch,
                               int *error) {
    const int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();
        volatile uint32_t *dst = peer_flags[t] + rank;
        *dst = epoch;
        __threadfence_system();
        unsigned long long spins = 0;
        while ((int32_t)(local_flags[t] - epoch) < 0) {
            if (++spins > 50000000ull) {
                atomicExch(error, 3);
                break;
            }
        }
    }
}

// ---- pair-distance histogram through the cell list (raw counts of g(r)); rmax <= cell side ------------------
template <int DIM>
__global__ void __launch_bounds__(kBoxThreads) k_box_pair_histogram(const __grid_constant__ BoxArgs A, int sa, int sb, double rmax,
                                                                    int nbins, unsigned long long *hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Stencil<DIM> st;
    double *sr = (double *)smem_raw;
    uint8_t *ssp = (uint8_t *)(sr + DIM * A.cap);
    int *sid = (int *)(ssp + ((A.cap + 3) & ~3));
    unsigned int *sh = (unsigned int *)(sid + A.cap);
    const int tid = threadIdx.x;
    for (int k = tid; k < nbins; k += kBoxThreads) sh[k] = 0u;
    int cc[3] = {0, 0, 0};
    {
        int l = blockIdx.x;
        if constexpr (DIM == 3) { cc[2] = l % A.g.nc[2]; l /= A.g.nc[2]; }
        cc[1] = l % A.g.nc[1];
        cc[0] = l / A.g.nc[1];
    }
    const int ncand = load_stencil<DIM>(A, cc, &st, sr, ssp);
    if (ncand < 0) return;
    // particle ids of the candidates (to count every unordered pair once: only id_i < id_j)
    for (int s = 0; s < Stencil<DIM>::NST; s++) {
        const int b = A.start[st.cell[s]], n = st.off[s + 1] - st.off[s];
        for (int p = tid; p < n; p += kBoxThreads) sid[st.off[s] + p] = A.ids[b + p];
    }
    __syncthreads();
    const int ncen = st.off[1];
    const double inv_dr = (double)nbins / rmax, rmax2 = rmax * rmax;
    for (int k = 0; k < ncen; k++) {
        const double xi[3] = {sr[k], sr[A.cap + k], DIM == 3 ? sr[2 * A.cap + k] : 0.0};
        const int si = ssp[k], idi = sid[k];
        for (int j = tid; j < ncand; j += kBoxThreads) {
            if (sid[j] <= idi) continue;
            const int sj = ssp[j];
            const bool match = (sa < 0 && sb < 0) || (sa < 0 && (si == sb || sj == sb)) || (sb < 0 && (si == sa || sj == sa)) ||
                               (si == sa && sj == sb) || (si == sb && sj == sa);
            if (!match) continue;
            const double r2 = d2_frame<DIM>(sr, A.cap, j, xi);
            if (r2 < rmax2) {
                int bin = (int)(sqrt(r2) * inv_dr);
                bin = bin < nbins ? bin : nbins - 1;
                atomicAdd(&sh[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < nbins; k += kBoxThreads)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// deterministic reduction of per-cell values: out[0] (+)= scale * sum(cellE), acc[0] += sum(cell_acc)
__global__ void k_box_reduce(const double *__restrict__ cellE, const uint32_t *__restrict__ cell_acc, int n, double scale,
                             int accumulate, double *outE, unsigned long long *out_acc) {
    __shared__ double s_e[32];
    __shared__ unsigned long long s_a[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    double e = 0.0;
    unsigned long long a = 0;
    for (int k = tid; k < n; k += blockDim.x) {
        e += cellE[k];
        if (cell_acc) a += cell_acc[k];
    }
    e = warp_sum(e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
        s_e[warp] = e;
        s_a[warp] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double te = 0.0;
        unsigned long long ta = 0;
        for (int w = 0; w < nwarp; w++) {
            te += s_e[w];
            ta += s_a[w];
        }
        outE[0] = (accumulate ? outE[0] : 0.0) + scale * te;
        if (out_acc) out_acc[0] += ta;
    }
}

}  // namespace

// =================================================================================================
struct BoxState {
    pmc_config cfg{};
    cudaStream_t stream = nullptr;
    Geom g{};
    int N = 0, dim = 0, ns = 0, cap = 0;
    double rcut_max = 0.0, T = 1.0, sigma = 0.05;
    unsigned long long seed = 0;
    uint32_t sweep = 0;
    bool geom_ready = false, model_ready = false;
    double *x = nullptr, *xs = nullptr, *par = nullptr, *cellE = nullptr, *eloc = nullptr, *energy = nullptr, *etmp = nullptr;
    int32_t *img = nullptr, *ids = nullptr, *start = nullptr, *cursor = nullptr, *count = nullptr, *cid = nullptr;
    uint8_t *sp = nullptr, *sps = nullptr;
    uint32_t *cell_acc = nullptr;
    unsigned long long *acc_total = nullptr;
    int *flags = nullptr;  // [0] bad input, [1] overflow
    double *raw = nullptr;
    long long *rsp = nullptr;
    int64_t calls = 0;
    int64_t launches = 0;
    size_t smem = 0;
    int fast_kc = 0;        // > 0: k_box_sweep_fast<.., KC> is used for the sweeps
    // multi-GPU replicas (box_peer_attach)
    int rank = 0, world = 1;
    uint32_t *bar_flags = nullptr;      // [8] local barrier flags
    uint32_t **d_peer_flags = nullptr;  // device array [world] of flag arrays (self included)
    uint32_t epoch = 0;
    unsigned char *shared_block = nullptr;  // ONE allocation [x | xs | img | cellE | cell_acc | flags]: one IPC handle
    size_t shared_bytes = 0;
    unsigned char *peer_block[PMC_MAX_PEERS + 1] = {};  // opened IPC mapping of every rank's block (self = local)
    size_t fast_smem = 0;
    int ncell_alloc = 0;
    int32_t *cid_blocksum = nullptr;  // [ceil(ncell/1024)] partial sums of the cell-count scan
};

namespace {

template <typename T>
cudaError_t balloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, sizeof(T) * (n ? n : 1));
    return e;
}

template <typename F>
int bdispatch(int dim, int model, F &&f) {
#define PMC_BCASE(D, MDL) \
    if (dim == D && model == MDL) return f(std::integral_constant<int, D>{}, std::integral_constant<int, MDL>{});
    PMC_BCASE(3, PMC_MODEL_LJ)
    PMC_BCASE(2, PMC_MODEL_LJ)
    PMC_BCASE(3, PMC_MODEL_SOFT)
    PMC_BCASE(2, PMC_MODEL_SOFT)
    PMC_BCASE(3, PMC_MODEL_SMOOTHLJ)
    PMC_BCASE(2, PMC_MODEL_SMOOTHLJ)
    PMC_BCASE(3, PMC_MODEL_KG)
    PMC_BCASE(2, PMC_MODEL_KG)
#undef PMC_BCASE
    return bfail(PMC_ERR_INVALID, "unsupported dim/model combination");
}

// Everything a peer GPU writes into lives in ONE allocation: one IPC handle, fixed offsets on every rank.
struct BlockLayout {
    size_t x, xs, img, cellE, cacc, flags, total;
};
BlockLayout block_layout(int N, int dim, int ncell) {
    BlockLayout l;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t p = o;
        o += (bytes + 255) & ~(size_t)255;
        return p;
    };
    l.x = take(sizeof(double) * dim * (size_t)N);
    l.xs = take(sizeof(double) * dim * (size_t)N);
    l.img = take(sizeof(int32_t) * dim * (size_t)N);
    l.cellE = take(sizeof(double) * (size_t)ncell);
    l.cacc = take(sizeof(uint32_t) * (size_t)ncell);
    l.flags = take(256);
    l.total = o;
    return l;
}

void fill_args(BoxState *b, BoxArgs &a) {
    a.g = b->g;
    a.N = b->N;
    a.ns = b->ns;
    a.cap = b->cap;
    a.x = b->x;
    a.img = b->img;
    a.sp = b->sp;
    a.xs = b->xs;
    a.sps = b->sps;
    a.ids = b->ids;
    a.start = b->start;
    a.par = b->par;
    a.T = b->T;
    a.sigma = (float)b->sigma;
    a.seed = b->seed;
    a.sweep = b->sweep;
    a.cellE = b->cellE;
    a.cell_acc = b->cell_acc;
    a.eloc = b->eloc;
    a.overflow = b->flags + 1;
    a.cta_offset = 0;
    a.n_peers = 0;
    if (b->world > 1) {
        const BlockLayout bl = block_layout(b->N, b->dim, b->g.ncell);
        for (int r = 0; r < b->world; r++) {
            if (r == b->rank) continue;
            const int p = a.n_peers++;
            unsigned char *blk = b->peer_block[r];
            a.peer_x[p] = (double *)(blk + bl.x);
            a.peer_xs[p] = (double *)(blk + bl.xs);
            a.peer_img[p] = (int32_t *)(blk + bl.img);
            a.peer_cellE[p] = (double *)(blk + bl.cellE);
            a.peer_cacc[p] = (uint32_t *)(blk + bl.cacc);
        }
    }
}

// K1: (re)build the cell-sorted arrays for grid origin g.shift
int build_cells(BoxState *b) {
    const int N = b->N, nb = (N + 255) / 256;
    BCU(cudaMemsetAsync(b->count, 0, sizeof(int32_t) * b->g.ncell, b->stream));
    if (b->dim == 3)
        k_box_count<3><<<nb, 256, 0, b->stream>>>(b->x, N, b->g, b->cid, b->count);
    else
        k_box_count<2><<<nb, 256, 0, b->stream>>>(b->x, N, b->g, b->cid, b->count);
    {
        const int nsb = (b->g.ncell + 1023) / 1024;
        k_box_scan_local<<<nsb, 1024, 0, b->stream>>>(b->count, b->start, b->cid_blocksum, b->g.ncell);
        k_box_scan_apply<<<nsb, 1024, 0, b->stream>>>(b->start, b->cursor, b->cid_blocksum, b->g.ncell);
    }
    k_box_scatter<<<nb, 256, 0, b->stream>>>(b->cid, b->start, b->cursor, N, b->ids);
    const int nbc = (b->g.ncell + 7) / 8;  // one warp per cell, 8 warps per CTA
    if (b->dim == 3)
        k_box_finalize<3><<<nbc, 256, 0, b->stream>>>(b->start, b->ids, b->g.ncell, N, b->x, b->sp, b->xs, b->sps);
    else
        k_box_finalize<2><<<nbc, 256, 0, b->stream>>>(b->start, b->ids, b->g.ncell, N, b->x, b->sp, b->xs, b->sps);
    BCU(cudaGetLastError());
    b->launches += 5;
    return PMC_OK;
}

int setup_geometry(BoxState *b, const double *box3) {
    if (!b->model_ready) return bfail(PMC_ERR_STATE, "pmc_set_model must precede pmc_upload in PMC_MODE_BOX");
    b->g.ncell = 1;
    for (int a = 0; a < 3; a++) {
        b->g.nc[a] = 1;
        b->g.L[a] = 1.0;
        b->g.cs[a] = 1.0;
        b->g.shift[a] = 0.0;
    }
    double occ = (double)b->N;
    for (int a = 0; a < b->dim; a++) {
        // cells of side >= rcut_max (src/neighbours.jl:236-238), count rounded down to even for the colouring
        int n = (int)std::floor(box3[a] / b->rcut_max);
        n -= n % 2;
        if (n < 2)
            return bfail(PMC_ERR_UNSUPPORTED, "box side %g holds fewer than 2 cells of side >= rcut %g: use PMC_MODE_CHAINS",
                         box3[a], b->rcut_max);
        b->g.nc[a] = n;
        b->g.L[a] = box3[a];
        b->g.cs[a] = box3[a] / (double)n;
        b->g.ncell *= n;
    }
    occ /= (double)b->g.ncell;
    const int nst = b->dim == 3 ? 27 : 9;
    int cap = (int)(occ * nst * 1.5) + 96;
    cap = (cap + 31) / 32 * 32;
    b->cap = cap;
    b->smem = sizeof(double) * (size_t)b->dim * cap + 2 * (size_t)cap + 16;
    if (b->smem > 200 * 1024) return bfail(PMC_ERR_UNSUPPORTED, "stencil of %d candidates does not fit shared memory", cap);
    // fast sweep kernel: cubic cells, stencil fits kBfThreads x KC register candidates
    b->fast_kc = 0;
    bool cubic = true;
    for (int a = 1; a < b->dim; a++) cubic = cubic && b->g.cs[a] == b->g.cs[0];
    if (cubic && b->cfg.prefilter >= 0) {
        // register-candidate budget: 1.45 x the mean stencil occupancy covers a simple-cubic lattice start, where a
        // 3-cell span holds 8 or 9 lattice planes (overflow is detected and reported, never silent)
        const int need = (int)(occ * nst * 1.45) + 16;
        constexpr int kcs[4] = {kBfKc0, kBfKc1, kBfKc2, kBfKc3};
        for (int q = 3; q >= 0; q--)
            if (need <= kBfThreads * kcs[q]) b->fast_kc = kcs[q];
    }
    if (b->fast_kc) {
        b->cap = kBfThreads * b->fast_kc;
        b->fast_smem = bf_layout(b->dim, b->cap).total;
        b->smem = sizeof(double) * (size_t)b->dim * b->cap + 2 * (size_t)b->cap + 16;
    }
    int rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
        if (b->fast_kc) {
            BCU(cudaFuncSetAttribute(k_box_sweep_fast<decltype(D)::value, decltype(MDL)::value, kBfKc0>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf_layout(b->dim, kBfThreads * kBfKc0).total));
            BCU(cudaFuncSetAttribute(k_box_sweep_fast<decltype(D)::value, decltype(MDL)::value, kBfKc1>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf_layout(b->dim, kBfThreads * kBfKc1).total));
            BCU(cudaFuncSetAttribute(k_box_sweep_fast<decltype(D)::value, decltype(MDL)::value, kBfKc2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf_layout(b->dim, kBfThreads * kBfKc2).total));
            BCU(cudaFuncSetAttribute(k_box_sweep_fast<decltype(D)::value, decltype(MDL)::value, kBfKc3>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf_layout(b->dim, kBfThreads * kBfKc3).total));
        }
        BCU(cudaFuncSetAttribute(k_box_sweep<decltype(D)::value, decltype(MDL)::value>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
        BCU(cudaFuncSetAttribute(k_box_energy<decltype(D)::value, decltype(MDL)::value>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
        return (int)PMC_OK;
    });
    if (rc) return rc;
    if (b->ncell_alloc != b->g.ncell) {  // buffers survive re-uploads of the same geometry (peers map them)
        if (b->world > 1) return bfail(PMC_ERR_STATE, "the cell grid changed after peers were attached");
        for (void *p : {(void *)b->count, (void *)b->cursor, (void *)b->start, (void *)b->shared_block})
            if (p) cudaFree(p);
        b->count = b->cursor = b->start = nullptr;
        b->shared_block = nullptr;
        BCU(balloc(&b->count, b->g.ncell));
        BCU(balloc(&b->cursor, b->g.ncell));
        BCU(balloc(&b->start, b->g.ncell + 1));
        if (b->cid_blocksum) cudaFree(b->cid_blocksum);
        BCU(balloc(&b->cid_blocksum, (b->g.ncell + 1023) / 1024));
        // everything a peer GPU writes into lives in ONE allocation, so one IPC handle and fixed offsets suffice
        const BlockLayout bl = block_layout(b->N, b->dim, b->g.ncell);
        BCU(balloc(&b->shared_block, bl.total));
        b->shared_bytes = bl.total;
        b->x = (double *)(b->shared_block + bl.x);
        b->xs = (double *)(b->shared_block + bl.xs);
        b->img = (int32_t *)(b->shared_block + bl.img);
        b->cellE = (double *)(b->shared_block + bl.cellE);
        b->cell_acc = (uint32_t *)(b->shared_block + bl.cacc);
        b->bar_flags = (uint32_t *)(b->shared_block + bl.flags);
        BCU(cudaDeviceSynchronize());  // zero-fills ran on the legacy default stream
        b->ncell_alloc = b->g.ncell;
    }
    b->geom_ready = true;
    return PMC_OK;
}

int check_overflow(BoxState *b) {
    int fl[2] = {0, 0};
    BCU(cudaMemcpyAsync(fl, b->flags, sizeof fl, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    if (fl[1]) return bfail(PMC_ERR_UNSUPPORTED, "a 3^d-cell neighbourhood holds more than %d particles (density too inhomogeneous)", b->cap);
    if (fl[0] == 3) return bfail(PMC_ERR_CUDA, "inter-GPU barrier timed out: a peer rank did not arrive");
    return PMC_OK;
}

// local energies (grid origin 0) -> eloc in particle order, etmp[0] = sum/2
int compute_energy(BoxState *b) {
    for (int a = 0; a < 3; a++) b->g.shift[a] = 0.0;
    int rc = build_cells(b);
    if (rc) return rc;
    BoxArgs A;
    fill_args(b, A);
    rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
        k_box_energy<decltype(D)::value, decltype(MDL)::value><<<b->g.ncell, kBoxThreads, b->smem, b->stream>>>(A);
        BCU(cudaGetLastError());
        return (int)PMC_OK;
    });
    if (rc) return rc;
    k_box_reduce<<<1, 1024, 0, b->stream>>>(b->cellE, nullptr, b->g.ncell, 0.5, 0, b->etmp, nullptr);
    BCU(cudaGetLastError());
    b->launches += 2;
    return check_overflow(b);
}

// one inter-GPU barrier on the context's stream
int peer_barrier(BoxState *b) {
    b->epoch++;
    k_peer_barrier<<<1, 32, 0, b->stream>>>(b->bar_flags, b->d_peer_flags, b->rank, b->world, b->epoch, b->flags);
    BCU(cudaGetLastError());
    return PMC_OK;
}

}  // namespace

const char *box_error() { return g_box_err.c_str(); }

int box_check(BoxState *b) { return check_overflow(b); }

int box_pair_histogram(BoxState *b, int sa, int sb, double rmax, int nbins, unsigned long long *d_hist) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    for (int a = 0; a < b->dim; a++)
        if (rmax > b->g.cs[a]) return bfail(PMC_ERR_INVALID, "rmax %g exceeds the cell side %g of the device cell list", rmax, b->g.cs[a]);
    for (int a = 0; a < 3; a++) b->g.shift[a] = 0.0;
    int rc = build_cells(b);
    if (rc) return rc;
    BoxArgs A;
    fill_args(b, A);
    const size_t smem = b->smem + sizeof(int) * (size_t)b->cap + sizeof(unsigned int) * (size_t)nbins + 16;
    if (b->dim == 3) {
        BCU(cudaFuncSetAttribute(k_box_pair_histogram<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_pair_histogram<3><<<b->g.ncell, kBoxThreads, smem, b->stream>>>(A, sa, sb, rmax, nbins, d_hist);
    } else {
        BCU(cudaFuncSetAttribute(k_box_pair_histogram<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_pair_histogram<2><<<b->g.ncell, kBoxThreads, smem, b->stream>>>(A, sa, sb, rmax, nbins, d_hist);
    }
    BCU(cudaGetLastError());
    b->launches++;
    return check_overflow(b);
}

int box_peer_export(BoxState *b, unsigned char *handle64) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "pmc_upload must precede pmc_box_peer_export");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    BCU(cudaIpcGetMemHandle(&h, b->shared_block));
    memcpy(handle64, &h, 64);
    return PMC_OK;
}

int box_peer_attach(BoxState *b, int rank, int world, const unsigned char *handles) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "pmc_upload must precede pmc_box_peer_attach");
    if (world < 1 || world > PMC_MAX_PEERS + 1 || rank < 0 || rank >= world)
        return bfail(PMC_ERR_INVALID, "rank %d / world %d out of range (max %d ranks)", rank, world, PMC_MAX_PEERS + 1);
    if (b->world > 1) return bfail(PMC_ERR_STATE, "peers are already attached");
    const BlockLayout bl = block_layout(b->N, b->dim, b->g.ncell);
    std::vector<uint32_t *> flags(world);
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            b->peer_block[r] = b->shared_block;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)64 * r, 64);
            void *ptr = nullptr;
            BCU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            b->peer_block[r] = (unsigned char *)ptr;
        }
        flags[r] = (uint32_t *)(b->peer_block[r] + bl.flags);
    }
    if (b->d_peer_flags) cudaFree(b->d_peer_flags);
    BCU(cudaMalloc((void **)&b->d_peer_flags, sizeof(uint32_t *) * world));
    BCU(cudaMemcpy(b->d_peer_flags, flags.data(), sizeof(uint32_t *) * world, cudaMemcpyHostToDevice));
    b->rank = rank;
    b->world = world;
    b->epoch = 0;
    return PMC_OK;
}

int box_create(BoxState **out, const pmc_config &cfg) {
    if (cfg.n_chains != 1) return bfail(PMC_ERR_INVALID, "PMC_MODE_BOX holds exactly one system (n_chains = %d)", cfg.n_chains);
    if (cfg.molecules) return bfail(PMC_ERR_UNSUPPORTED, "Molecules are not supported in PMC_MODE_BOX");
    BoxState *b = new BoxState();
    b->cfg = cfg;
    b->N = cfg.n_particles;
    b->dim = cfg.dim;
    b->ns = cfg.n_species;
    const size_t N = b->N, d = b->dim;
    cudaError_t e = balloc(&b->ids, N);
    if (e == cudaSuccess) e = balloc(&b->cid, N);
    if (e == cudaSuccess) e = balloc(&b->sp, N);
    if (e == cudaSuccess) e = balloc(&b->sps, N);
    if (e == cudaSuccess) e = balloc(&b->eloc, N);
    if (e == cudaSuccess) e = balloc(&b->par, (size_t)PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR);
    if (e == cudaSuccess) e = balloc(&b->energy, 1);
    if (e == cudaSuccess) e = balloc(&b->etmp, 1);
    if (e == cudaSuccess) e = balloc(&b->acc_total, 1);
    if (e == cudaSuccess) e = balloc(&b->flags, 2);
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->raw, sizeof(double) * d * N);
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->rsp, sizeof(long long) * N);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // zero-fills ran on the legacy default stream
    if (e != cudaSuccess) {
        box_destroy(b);
        return bfail(PMC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    *out = b;
    return PMC_OK;
}

void box_destroy(BoxState *b) {
    if (!b) return;
    for (int r = 0; r < b->world; r++)
        if (r != b->rank && b->peer_block[r]) cudaIpcCloseMemHandle(b->peer_block[r]);
    void *bufs[] = {b->shared_block, b->ids, b->cid, b->sp, b->sps, b->eloc, b->par, b->energy, b->etmp, b->acc_total,
                    b->flags, b->raw, b->rsp, b->count, b->cursor, b->start, b->d_peer_flags, b->cid_blocksum};
    for (void *p : bufs)
        if (p) cudaFree(p);
    delete b;
}

void box_set_stream(BoxState *b, cudaStream_t st) { b->stream = st; }
void box_set_sigma(BoxState *b, double sigma) { b->sigma = sigma; }
void box_seed(BoxState *b, uint64_t seed) {
    b->seed = seed;
    b->sweep = 0;
}
int64_t box_take_launches(BoxState *b) {
    const int64_t n = b->launches;
    b->launches = 0;
    return n;
}

int box_set_model(BoxState *b, const double *params) {
    b->rcut_max = 0.0;
    for (int k = 0; k < b->ns * b->ns; k++) b->rcut_max = std::fmax(b->rcut_max, params[(size_t)k * PMC_NPAR + PMC_P_RCUT]);
    BCU(cudaMemcpyAsync(b->par, params, sizeof(double) * b->ns * b->ns * PMC_NPAR, cudaMemcpyHostToDevice, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    b->model_ready = true;
    return PMC_OK;
}

int box_upload(BoxState *b, const double *pos, const int64_t *species, const double *box3, double temperature) {
    int rc = setup_geometry(b, box3);
    if (rc) return rc;
    b->T = temperature;
    const size_t N = b->N, d = b->dim;
    BCU(cudaMemcpyAsync(b->raw, pos, sizeof(double) * d * N, cudaMemcpyHostToDevice, b->stream));
    BCU(cudaMemcpyAsync(b->rsp, species, sizeof(long long) * N, cudaMemcpyHostToDevice, b->stream));
    BCU(cudaMemsetAsync(b->flags, 0, 2 * sizeof(int), b->stream));
    k_box_ingest<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->raw, b->rsp, b->N, b->dim, b->ns, b->g, b->x, b->img, b->sp,
                                                            b->flags);
    BCU(cudaGetLastError());
    b->launches++;
    int fl[2];
    BCU(cudaMemcpyAsync(fl, b->flags, sizeof fl, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    if (fl[0] == 1) return bfail(PMC_ERR_INVALID, "positions contain NaN or Inf");
    if (fl[0] == 2) return bfail(PMC_ERR_INVALID, "species labels must lie in 1..%d", b->ns);
    b->calls = 0;
    BCU(cudaMemsetAsync(b->acc_total, 0, sizeof(unsigned long long), b->stream));
    return PMC_OK;
}

int box_total_energy(BoxState *b, double *e_out) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    int rc = compute_energy(b);
    if (rc) return rc;
    BCU(cudaMemcpyAsync(e_out, b->etmp, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    return PMC_OK;
}

int box_init_energy(BoxState *b, double *e_out) {
    int rc = box_total_energy(b, e_out);
    if (rc) return rc;
    BCU(cudaMemcpyAsync(b->energy, b->etmp, sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    return PMC_OK;
}

int box_local_energy(BoxState *b, double *eloc_out) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    int rc = compute_energy(b);
    if (rc) return rc;
    BCU(cudaMemcpyAsync(eloc_out, b->eloc, sizeof(double) * b->N, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    return PMC_OK;
}

int box_energy(BoxState *b, double *e_out) {
    BCU(cudaMemcpyAsync(e_out, b->energy, sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    return PMC_OK;
}

// One sweep = N trials: fresh random grid origin, rebuild the cell list, then the 2^d colours in a
// random order.  n_trials is rounded up to whole sweeps.
int box_run(BoxState *b, int64_t n_trials) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    const int64_t sweeps = (n_trials + b->N - 1) / b->N;
    const int ncol = 1 << b->dim;
    const uint32_t k0 = (uint32_t)b->seed, k1 = (uint32_t)(b->seed >> 32);
    int nactive = 1;
    for (int a = 0; a < b->dim; a++) nactive *= b->g.nc[a] / 2;
    for (int64_t s = 0; s < sweeps; s++) {
        const Philox4 r = philox4x32_10(b->sweep, 0u, 0u, 2u, k0, k1);
        const Philox4 r2 = philox4x32_10(b->sweep, 1u, 0u, 2u, k0, k1);
        for (int a = 0; a < b->dim; a++) b->g.shift[a] = b->g.cs[a] * ((double)r.v[a] * 0x1p-32);
        int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        for (int k = ncol - 1; k > 0; k--) {  // Fisher-Yates
            const int j = (int)(((uint64_t)r2.v[k % 4] >> (8 * (k / 4))) % (uint64_t)(k + 1));
            const int t = order[k];
            order[k] = order[j];
            order[j] = t;
        }
        int rc = build_cells(b);
        if (rc) return rc;
        BoxArgs A;
        fill_args(b, A);
        // multi-GPU: this rank sweeps an even share of the colour's active cells (same-colour cells never interact,
        // so any split is valid); peers must have finished their rebuild before anyone pushes into their arrays
        const int lo = (int)((int64_t)nactive * b->rank / b->world), hi = (int)((int64_t)nactive * (b->rank + 1) / b->world);
        A.cta_offset = lo;
        if (b->world > 1) {
            rc = peer_barrier(b);
            if (rc) return rc;
        }
        rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
            constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
            for (int k = 0; k < ncol; k++) {
                if (hi > lo) {
                    if (b->fast_kc == kBfKc0)
                        k_box_sweep_fast<d, mdl, kBfKc0><<<hi - lo, kBfThreads, b->fast_smem, b->stream>>>(A, order[k]);
                    else if (b->fast_kc == kBfKc1)
                        k_box_sweep_fast<d, mdl, kBfKc1><<<hi - lo, kBfThreads, b->fast_smem, b->stream>>>(A, order[k]);
                    else if (b->fast_kc == kBfKc2)
                        k_box_sweep_fast<d, mdl, kBfKc2><<<hi - lo, kBfThreads, b->fast_smem, b->stream>>>(A, order[k]);
                    else if (b->fast_kc == kBfKc3)
                        k_box_sweep_fast<d, mdl, kBfKc3><<<hi - lo, kBfThreads, b->fast_smem, b->stream>>>(A, order[k]);
                    else
                        k_box_sweep<d, mdl><<<hi - lo, kBoxThreads, b->smem, b->stream>>>(A, order[k]);
                }
                BCU(cudaGetLastError());
                if (b->world > 1) {
                    const int brc = peer_barrier(b);
                    if (brc) return brc;
                }
            }
            return (int)PMC_OK;
        });
        if (rc) return rc;
        k_box_reduce<<<1, 1024, 0, b->stream>>>(b->cellE, b->cell_acc, b->g.ncell, 1.0, 1, b->energy, b->acc_total);
        BCU(cudaGetLastError());
        b->launches += ncol + 1 + (b->world > 1 ? ncol + 1 : 0);
        b->sweep++;
        b->calls += b->N;
    }
    return PMC_OK;
}

int box_download(BoxState *b, double *pos, int64_t *species) {
    k_box_egress<<<(b->N + 255) / 256, 256, 0, b->stream>>>(b->x, b->img, b->sp, b->N, b->dim, b->g, b->raw, b->rsp);
    BCU(cudaGetLastError());
    b->launches++;
    BCU(cudaMemcpyAsync(pos, b->raw, sizeof(double) * (size_t)b->dim * b->N, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaMemcpyAsync(species, b->rsp, sizeof(long long) * (size_t)b->N, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    return check_overflow(b);
}

int box_counters(BoxState *b, int64_t *calls, int64_t *accepted) {
    unsigned long long a = 0;
    BCU(cudaMemcpyAsync(&a, b->acc_total, sizeof a, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    calls[0] = b->calls;
    accepted[0] = (int64_t)a;
    return PMC_OK;
}

}  // namespace pmc
