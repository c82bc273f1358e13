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

// exclusive prefix sum of count[0..n) into start[0..n], single CTA, chunked with a running carry
__global__ void k_box_scan(const int32_t *__restrict__ count, int32_t *start, int32_t *cursor, int n) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int k = base + tid;
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
            int w = lane < nwarp ? s_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
        if (k < n) {
            start[k] = excl;
            cursor[k] = 0;
        }
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = carry + s_warp[nwarp - 1];
        __syncthreads();
    }
    if (tid == 0) start[n] = s_carry;
}

__global__ void k_box_scatter(const int32_t *__restrict__ cid, const int32_t *__restrict__ start, int32_t *cursor, int N,
                              int32_t *ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = cid[i];
    ids[start[c] + atomicAdd(&cursor[c], 1)] = i;
}

// canonical order inside each cell (descending particle id) + gather into the sorted SoA
template <int DIM>
__global__ void k_box_finalize(const int32_t *__restrict__ start, int32_t *ids, int ncell, int N,
                               const double *__restrict__ x, const uint8_t *__restrict__ sp, double *xs, uint8_t *sps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int b = start[c], e = start[c + 1];
    for (int p = b + 1; p < e; p++) {  // insertion sort, descending
        const int v = ids[p];
        int q = p - 1;
        while (q >= b && ids[q] < v) {
            ids[q + 1] = ids[q];
            q--;
        }
        ids[q + 1] = v;
    }
    for (int p = b; p < e; p++) {
        const int i = ids[p];
#pragma unroll
        for (int a = 0; a < DIM; a++) xs[(size_t)a * N + p] = x[(size_t)a * N + i];
        sps[p] = sp[i];
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
    int off[NST + 1];
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
        int l = blockIdx.x;
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
            A.xs[(size_t)a * A.N + b + k] = xnew;
            A.x[(size_t)a * A.N + i] = xnew;
            if (w) A.img[(size_t)a * A.N + i] += w;
        }
    }
    if (tid == 0) {
        A.cellE[cell] = Esum;
        A.cell_acc[cell] = nacc;
    }
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
}

// K1: (re)build the cell-sorted arrays for grid origin g.shift
int build_cells(BoxState *b) {
    const int N = b->N, nb = (N + 255) / 256;
    BCU(cudaMemsetAsync(b->count, 0, sizeof(int32_t) * b->g.ncell, b->stream));
    if (b->dim == 3)
        k_box_count<3><<<nb, 256, 0, b->stream>>>(b->x, N, b->g, b->cid, b->count);
    else
        k_box_count<2><<<nb, 256, 0, b->stream>>>(b->x, N, b->g, b->cid, b->count);
    k_box_scan<<<1, 1024, 0, b->stream>>>(b->count, b->start, b->cursor, b->g.ncell);
    k_box_scatter<<<nb, 256, 0, b->stream>>>(b->cid, b->start, b->cursor, N, b->ids);
    const int nbc = (b->g.ncell + 127) / 128;
    if (b->dim == 3)
        k_box_finalize<3><<<nbc, 128, 0, b->stream>>>(b->start, b->ids, b->g.ncell, N, b->x, b->sp, b->xs, b->sps);
    else
        k_box_finalize<2><<<nbc, 128, 0, b->stream>>>(b->start, b->ids, b->g.ncell, N, b->x, b->sp, b->xs, b->sps);
    BCU(cudaGetLastError());
    b->launches += 4;
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
    int rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
        BCU(cudaFuncSetAttribute(k_box_sweep<decltype(D)::value, decltype(MDL)::value>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
        BCU(cudaFuncSetAttribute(k_box_energy<decltype(D)::value, decltype(MDL)::value>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
        return (int)PMC_OK;
    });
    if (rc) return rc;
    for (void *p : {(void *)b->count, (void *)b->cursor, (void *)b->start, (void *)b->cellE, (void *)b->cell_acc})
        if (p) cudaFree(p);
    b->count = b->cursor = b->start = nullptr;
    b->cellE = nullptr;
    b->cell_acc = nullptr;
    BCU(balloc(&b->count, b->g.ncell));
    BCU(balloc(&b->cursor, b->g.ncell));
    BCU(balloc(&b->start, b->g.ncell + 1));
    BCU(balloc(&b->cellE, b->g.ncell));
    BCU(balloc(&b->cell_acc, b->g.ncell));
    BCU(cudaDeviceSynchronize());  // zero-fills ran on the legacy default stream
    b->geom_ready = true;
    return PMC_OK;
}

int check_overflow(BoxState *b) {
    int fl[2] = {0, 0};
    BCU(cudaMemcpyAsync(fl, b->flags, sizeof fl, cudaMemcpyDeviceToHost, b->stream));
    BCU(cudaStreamSynchronize(b->stream));
    if (fl[1]) return bfail(PMC_ERR_UNSUPPORTED, "a 3^d-cell neighbourhood holds more than %d particles (density too inhomogeneous)", b->cap);
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

}  // namespace

const char *box_error() { return g_box_err.c_str(); }

int box_create(BoxState **out, const pmc_config &cfg) {
    if (cfg.n_chains != 1) return bfail(PMC_ERR_INVALID, "PMC_MODE_BOX holds exactly one system (n_chains = %d)", cfg.n_chains);
    if (cfg.molecules) return bfail(PMC_ERR_UNSUPPORTED, "Molecules are not supported in PMC_MODE_BOX");
    BoxState *b = new BoxState();
    b->cfg = cfg;
    b->N = cfg.n_particles;
    b->dim = cfg.dim;
    b->ns = cfg.n_species;
    const size_t N = b->N, d = b->dim;
    cudaError_t e = balloc(&b->x, d * N);
    if (e == cudaSuccess) e = balloc(&b->xs, d * N);
    if (e == cudaSuccess) e = balloc(&b->img, d * N);
    if (e == cudaSuccess) e = balloc(&b->ids, N);
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
    void *bufs[] = {b->x, b->xs, b->img, b->ids, b->cid, b->sp, b->sps, b->eloc, b->par, b->energy, b->etmp, b->acc_total,
                    b->flags, b->raw, b->rsp, b->count, b->cursor, b->start, b->cellE, b->cell_acc};
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
        rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
            for (int k = 0; k < ncol; k++)
                k_box_sweep<decltype(D)::value, decltype(MDL)::value><<<nactive, kBoxThreads, b->smem, b->stream>>>(A, order[k]);
            BCU(cudaGetLastError());
            return (int)PMC_OK;
        });
        if (rc) return rc;
        k_box_reduce<<<1, 1024, 0, b->stream>>>(b->cellE, b->cell_acc, b->g.ncell, 1.0, 1, b->energy, b->acc_total);
        BCU(cudaGetLastError());
        b->launches += ncol + 1;
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
