// box.cu -- PMC_MODE_BOX: a single large periodic box updated with checkerboard parallel sweeps.
//
// Reference counterparts
//   cell list build      src/neighbours.jl:236-270 (LinkedList ctor + build_neighbour_list!)  -> K1:
//                        bin -> prefix sum per x-plane of cells -> scatter -> per-cell canonical order
//                        (descending particle id, the order head-insertion produces, neighbours.jl:257-268)
//                        + gather of in-cell coordinates (fp64 and packed bytes) into the sorted arrays
//   local / total energy src/atoms.jl:40-58, :81-88                                          -> K2
//   Metropolis trial     src/moves.jl:57-90 + src/utils.jl:8-10                               -> K5
// K5 has NO reference counterpart as an algorithm: the reference updates one particle at a time over
// the whole box.  Here the box is cut into cells of side >= rcut_max with an even cell count per axis;
// cells of one colour (2^d colours) do not interact, so each is advanced independently by one CTA for
// n_cell trials while its 3^d-cell neighbourhood is frozen.  Moves leaving the cell are rejected and the
// grid origin is shifted by a fresh random vector every sweep (Anderson et al., J. Comput. Phys. 254
// (2013) 27), which keeps detailed balance per sub-sweep and restores ergodicity.  Parity with the
// reference is therefore statistical for trajectories and exact (1e-12) for energies.
// One sweep = 4 rebuild kernels + ONE persistent sweep kernel (all colours; a cell waits for the completion stamps of
// its own neighbour cells, no barrier between the phases) + one deterministic reduction.  Over several GPUs the cell
// grid is cut into x-slabs: a rank keeps cell lists for its slab plus a halo plane, sweeps its own cells and stores
// accepted moves straight into its peers' memory (NVLink, CUDA IPC); flags chain the sweeps, nothing else does.
#include "box.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
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

// Flag block of a rank (inside the IPC-shared allocation): peer r announces in MY block that it finished the sweep /
// the cell rebuild of stamp s.  Plain 32-bit stores over NVLink, each behind a system-scope fence.
constexpr int kFlagSwept = 0, kFlagRebuilt = 8, kFlagWords = 64;

struct BoxArgs {
    Geom g;
    int N, ns, cap;
    // canonical state (particle order)
    double *x;     // [dim][N] wrapped
    int32_t *img;  // [dim][N]
    uint8_t *sp;   // [N]
    // cell-sorted state of the current grid.  Slots are reserved per x-PLANE of cells (plane_cap each), so the slot of
    // (cell, k) is the same on every rank that keeps that plane without any rank having to bin the whole box.
    int plane_cap;
    int plane_lo, plane_n;  // this rank keeps cell lists for planes plane_lo .. plane_lo + plane_n - 1 (mod nc[0])
    double *rs;             // [dim][nslot] coordinates INSIDE the particle's own cell, [0, cs)
    uint32_t *qs;           // [nslot] the same as packed bytes in units of cs / 64 (prefilter, common.cuh)
    uint8_t *sps;           // [nslot]
    int32_t *ids;           // [nslot] particle id of each sorted slot
    int32_t *start;         // [ncell] first slot of a cell
    int32_t *count;         // [ncell]
    const double *par;
    double T;
    float sigma;
    unsigned long long seed;
    uint32_t sweep;
    // per-cell outputs
    double *cellE;            // [ncell] sum of accepted dE (sweep) or sum of local energies (energy)
    uint32_t *cell_acc;       // [ncell]
    double *eloc;             // [N] (energy kernel)
    int *overflow;  // 1: a stencil exceeds `cap`, 2: a plane exceeds plane_cap
    int *error;     // 3: a wait on another GPU timed out
    unsigned long long *stats;  // pmc_work_counters (nullptr: off)
    // dataflow of one sweep: persistent CTAs pull (phase, cell) units from work[0]; a cell starts when the neighbour
    // cells of EARLIER phases carry this sweep's stamp in done[] -- no barrier between the colour phases
    uint32_t stamp;
    uint32_t *done;   // [ncell]
    int *work;        // [0] next unit, [1] finished units
    int cell_lo, cell_n;  // this rank's share of every colour: active cells [cell_lo, cell_lo + cell_n)
    int order[8];         // colour processed in phase k
    int phase_of[8];      // inverse
    // multi-GPU: every rank holds the canonical state; cell lists only for its planes.  Accepted moves are stored
    // straight into the peers' arrays over NVLink peer memory (canonical state: every peer; sorted copy and done[]:
    // the peers that keep the plane)
    int rank, n_peers;
    volatile uint32_t *flags;  // my flag block
    int peer_rank[PMC_MAX_PEERS];
    int peer_plane_lo[PMC_MAX_PEERS], peer_plane_n[PMC_MAX_PEERS];
    double *peer_x[PMC_MAX_PEERS];
    double *peer_rs[PMC_MAX_PEERS];
    uint32_t *peer_qs[PMC_MAX_PEERS];
    int32_t *peer_img[PMC_MAX_PEERS];
    double *peer_cellE[PMC_MAX_PEERS];
    uint32_t *peer_cacc[PMC_MAX_PEERS];
    uint32_t *peer_done[PMC_MAX_PEERS];
    uint32_t *peer_flags[PMC_MAX_PEERS];
};

// Bounded wait on a 32-bit word another agent (CTA or GPU) will set: never hangs the device -- a peer that does not
// arrive within ~20 s raises the error flag and the kernel runs on (the host reports PMC_ERR_CUDA at the next sync).
__device__ __forceinline__ bool wait_for(const volatile uint32_t *p, uint32_t want, int *error, bool at_least) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; spins++) {
        // acquire load at system scope: pairs with the fence + flag store of the CTA (or GPU) that produced the data, so
        // that this thread's later loads -- and, through the barrier that follows every wait, its CTA's -- see that data
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (at_least ? (int32_t)(v - want) >= 0 : v == want) return true;
        if (*(volatile int *)error == 3) return false;
        if ((spins & 1023u) == 1023u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            if (t - t0 > 20000000000ull) {
                atomicExch(error, 3);
                return false;
            }
        }
        if (spins > 64) __nanosleep(200);
    }
}

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
// byte of an in-cell coordinate: units of cs / 64, [0, 63]
__device__ __forceinline__ uint32_t cell_byte(double r, double inv_unit) {
    int q = (int)floor(r * inv_unit);
    return (uint32_t)(q < 0 ? 0 : (q > 63 ? 63 : q));
}

template <int DIM>
__device__ __forceinline__ int lin_cell(const int (&c)[3], const int (&nc)[3]) {  // last axis fastest (neighbours.jl:79-88)
    int l = c[0];
    l = l * nc[1] + c[1];
    if constexpr (DIM == 3) l = l * nc[2] + c[2];
    return l;
}
__device__ __forceinline__ bool plane_kept(int px, int lo, int n, int nx) {
    int d = px - lo;
    if (d < 0) d += nx;
    return d < n;
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
// Replaces build_neighbour_list! (src/neighbours.jl:251-270).  A rank bins only the particles whose x-plane of cells
// it keeps (all planes on one GPU); slots are reserved per plane (plane_cap), so the layout of a plane is the same on
// every rank that keeps it.  With peers, the first kernel waits until every rank has finished the previous sweep (their
// pushes into the canonical state have landed), the last one announces the finished rebuild.
template <int DIM>
__global__ void k_box_count(const __grid_constant__ BoxArgs A, int32_t *cid, uint32_t wait_stamp) {
    if (A.n_peers && wait_stamp) {
        if (threadIdx.x < A.n_peers) wait_for(A.flags + kFlagSwept + A.peer_rank[threadIdx.x], wait_stamp, A.error, true);
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.N) return;
    int c[3] = {0, 0, 0};
    c[0] = cell_of(__ldcg(A.x + i), A.g.shift[0], A.g.L[0], A.g.cs[0], A.g.nc[0]);
    if (!plane_kept(c[0], A.plane_lo, A.plane_n, A.g.nc[0])) {
        cid[i] = -1;
        return;
    }
#pragma unroll
    for (int a = 1; a < DIM; a++) c[a] = cell_of(__ldcg(A.x + (size_t)a * A.N + i), A.g.shift[a], A.g.L[a], A.g.cs[a], A.g.nc[a]);
    const int l = lin_cell<DIM>(c, A.g.nc);
    cid[i] = l;
    atomicAdd(&A.count[l], 1);
}

// one CTA per kept plane: exclusive prefix sum of the plane's cell counts -> start[], cursor[] = 0
__global__ void __launch_bounds__(1024) k_box_scan_plane(const __grid_constant__ BoxArgs A, int32_t *cursor) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int px = A.plane_lo + blockIdx.x;
    if (px >= A.g.nc[0]) px -= A.g.nc[0];
    const int cpp = A.g.ncell / A.g.nc[0];
    const int c0 = px * cpp;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < cpp; base += 1024) {
        const int k = base + tid;
        const int v = k < cpp ? A.count[c0 + k] : 0;
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
        const int carry = s_carry;
        if (k < cpp) {
            A.start[c0 + k] = px * A.plane_cap + carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
            cursor[c0 + k] = 0;
        }
        __syncthreads();
        if (tid == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0 && s_carry > A.plane_cap) atomicExch(A.overflow, 2);
}

__global__ void k_box_scatter(const int32_t *__restrict__ cid, const int32_t *__restrict__ start, int32_t *cursor, int N,
                              int32_t *ids, const int *overflow) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || *overflow) return;
    const int c = cid[i];
    if (c >= 0) ids[start[c] + atomicAdd(&cursor[c], 1)] = i;
}

// canonical order inside each cell (descending particle id = the order head insertion produces, neighbours.jl:257-268)
// + gather into the sorted arrays as IN-CELL coordinates (fp64 and packed bytes): one warp per cell, rank sort
// through shuffles (cells hold ~20 particles); cells with more than 32 particles fall back to a serial insertion sort
template <int DIM>
__global__ void k_box_finalize(const __grid_constant__ BoxArgs A, int *ticket, uint32_t announce_stamp) {
    const int lane = threadIdx.x & 31;
    const int cpp = A.g.ncell / A.g.nc[0];
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // cell within the kept planes
    if (k < A.plane_n * cpp && !*A.overflow) {
        int px = A.plane_lo + k / cpp;
        if (px >= A.g.nc[0]) px -= A.g.nc[0];
        const int c = px * cpp + k % cpp;
        int cc[3] = {0, 0, 0};
        {
            int l = c;
            if constexpr (DIM == 3) { cc[2] = l % A.g.nc[2]; l /= A.g.nc[2]; }
            cc[1] = l % A.g.nc[1];
            cc[0] = l / A.g.nc[1];
        }
        const int b = A.start[c], cnt = A.count[c], e = b + cnt;
        const size_t ns_ = (size_t)A.g.nc[0] * A.plane_cap;
        auto put = [&](int p, int i) {
            uint32_t q = 0;
#pragma unroll
            for (int a = 0; a < DIM; a++) {
                const double r = in_frame(__ldcg(A.x + (size_t)a * A.N + i), A.g.shift[a], A.g.L[a], A.g.cs[a], cc[a]);
                A.rs[(size_t)a * ns_ + p] = r;
                q |= cell_byte(r, 64.0 / A.g.cs[a]) << (8 * a);
            }
            A.qs[p] = q;
            A.sps[p] = A.sp[i];
        };
        if (cnt <= 32) {
            const int mine = lane < cnt ? A.ids[b + lane] : -1;
            int rank = 0;
            for (int j = 0; j < cnt; j++) {
                const int v = __shfl_sync(0xffffffffu, mine, j);
                rank += (v > mine) ? 1 : 0;
            }
            __syncwarp();
            if (lane < cnt) {
                A.ids[b + rank] = mine;
                put(b + rank, mine);
            }
        } else {
            if (lane == 0) {
                for (int p = b + 1; p < e; p++) {  // insertion sort, descending
                    const int v = A.ids[p];
                    int q = p - 1;
                    while (q >= b && A.ids[q] < v) {
                        A.ids[q + 1] = A.ids[q];
                        q--;
                    }
                    A.ids[q + 1] = v;
                }
            }
            __syncwarp();
            for (int p = b + lane; p < e; p += 32) put(p, A.ids[p]);
        }
    }
    if (A.n_peers && announce_stamp) {  // the last CTA tells every peer that my arrays may be written into again
        __threadfence();
        __syncthreads();
        __shared__ int s_last;
        if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last && (int)threadIdx.x < A.n_peers) {
            __threadfence_system();
            *(volatile uint32_t *)(A.peer_flags[threadIdx.x] + kFlagRebuilt + A.rank) = announce_stamp;
        }
        if (s_last && threadIdx.x == 0) *ticket = 0;
    }
}

// ---- stencil loader ---------------------------------------------------------------------------------------
// Gathers the particles of the 3^d cells around `cc` into shared memory in the frame of the central cell (its
// particles in [0, cs)^d, neighbours shifted by whole cells), so no per-pair minimum image is needed.  The central
// cell comes first: candidates [0, ncen) are the movable particles.  The sorted arrays already hold in-cell
// coordinates, so a candidate costs its loads and one add per axis.
template <int DIM>
struct Stencil {
    static constexpr int NST = DIM == 3 ? 27 : 9;
    int cell[NST];
    int off[NST + 1];  // candidate offsets of the stencil cells in the gathered list
    int cnt[NST];      // particles in each stencil cell
    int base[NST];     // first sorted slot of each stencil cell
    double sh[NST][3];   // frame shift of the cell: o * cs
    uint32_t qsh[NST];   // the same in packed bytes: (o + 1) * 64 per axis
};

// The candidate index space of the whole stencil is spread over ALL threads (thread t handles candidates t, t + NT,
// ...), so the global loads of one CTA are independent and in flight together.  `sink(t, q)` receives the packed
// byte coordinates of candidate t in the stencil frame (fast kernel: into the thread's registers).  Loads bypass L1:
// inside a persistent kernel the sorted arrays change between the cells a CTA visits (own CTAs, other SMs, peer GPUs).
template <int DIM, int NT, typename Sink>
__device__ int load_stencil(const BoxArgs &A, const int (&cc)[3], Stencil<DIM> *st, double *sr, uint8_t *ssp, uint8_t *cellof, int cap,
                            Sink &&sink) {
    constexpr int NST = Stencil<DIM>::NST;
    const int tid = threadIdx.x;
    if (tid < NST) {
        // slot 0 = central cell; the others in first-axis-fastest order (Iterators.product, neighbours.jl:101)
        int k = tid == 0 ? (NST / 2) : (tid <= NST / 2 ? tid - 1 : tid);
        int c[3] = {0, 0, 0};
        uint32_t qsh = 0;
#pragma unroll
        for (int a = 0; a < DIM; a++) {
            const int oa = k % 3 - 1;
            k /= 3;
            int v = cc[a] + oa;
            if (v < 0) v += A.g.nc[a];
            if (v >= A.g.nc[a]) v -= A.g.nc[a];
            c[a] = v;
            st->sh[tid][a] = (double)oa * A.g.cs[a];
            qsh |= (uint32_t)((oa + 1) * 64) << (8 * a);
        }
        const int l = lin_cell<DIM>(c, A.g.nc);
        st->cell[tid] = l;
        st->qsh[tid] = qsh;
        st->base[tid] = __ldcg(A.start + l);
        st->cnt[tid] = __ldcg(A.count + l);
    }
    __syncthreads();
    if (tid == 0) {
        st->off[0] = 0;
        for (int k = 0; k < NST; k++) st->off[k + 1] = st->off[k] + st->cnt[k];
    }
    __syncthreads();
    const int nall = st->off[NST];
    if (nall > cap) {
        if (tid == 0) atomicExch(A.overflow, 1);
        return -1;
    }
    // stencil cell of every flat index, written once per cell (`cellof`: `cap` bytes of scratch) -- a binary search
    // over the 3^d offsets per candidate is five dependent shared-memory loads in front of every global load
    for (int s_ = tid >> 5; s_ < NST; s_ += NT >> 5)
        for (int p_ = tid & 31; p_ < st->cnt[s_]; p_ += 32) cellof[st->off[s_] + p_] = (uint8_t)s_;
    __syncthreads();
    const size_t ns_ = (size_t)A.g.nc[0] * A.plane_cap;
    // thread (warp, lane) owns the candidates t = 32 * (lane / 8) + 8 * warp + lane % 8 (mod NT): runs of eight neighbouring
    // candidates go to different warps, so the particles of one stencil cell -- inside a trial's filter sphere together or
    // not at all; the centre cell always -- are shared out (with whole 32-runs per warp, warp 0 holds the centre cell alone
    // and often needs a second pass over its survivors while the others wait at the block barrier).  Measured: +2.7 % where
    // the sweep is bound by one cell's latency (N = 131072, a GPU's share of the box at 8 GPUs), -0.6 % at N = 2^20 (a
    // warp's survivors sit on half of the banks); runs of 16: no effect either way; a stride-4 interleave: 13 % slower (a
    // quarter of the banks).
    for (int t = 32 * ((tid & 31) >> 3) + 8 * (tid >> 5) + (tid & 7); t < nall; t += NT) {
        static_assert(NT == 128, "candidate interleave assumes four warps");
        const int lo = cellof[t];
        const int slot = st->base[lo] + (t - st->off[lo]);
#pragma unroll
        for (int a = 0; a < DIM; a++) sr[a * cap + t] = __ldcg(A.rs + (size_t)a * ns_ + slot) + st->sh[lo][a];
        ssp[t] = __ldcg(A.sps + slot);
        sink(t, __ldcg(A.qs + slot) + st->qsh[lo]);
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
    const int ncand = load_stencil<DIM, kBoxThreads>(A, cc, &st, sr, ssp, ssp + A.cap, A.cap, [](int, uint32_t) {});
    if (ncand < 0) return;
    const int ncen = st.off[1], b = st.base[0];
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

// ---- K5: checkerboard sweep -------------------------------------------------------------------------------
// ONE persistent kernel per sweep.  CTAs pull (phase, cell) units from a global counter in phase order; a cell of
// phase k is started once the neighbour cells that belong to EARLIER phases carry this sweep's stamp in done[] (one
// warp polls its <= 26 flags).  There is no barrier between the colour phases, neither on the GPU nor between GPUs:
// a cell only ever waits for its own neighbourhood, the tail of one colour overlaps the head of the next, and with
// several GPUs only the cells next to a rank boundary wait for a flag that arrives over NVLink.  Units are handed
// out in phase order, so whatever a cell waits for has been handed out before it: no deadlock, whatever the grid.
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
// (floor of the exact product in one DFMA.RM against 2^52, low mantissa word -- DMUL + F2I.U64.FLOOR costs more issue
// slots and sits on the dependent path of every trial)
__device__ __forceinline__ uint32_t bf_fixed(double r, double cs, double scale) {
    return (uint32_t)__double2loint(__fma_rd(r + cs, scale, 4503599627370496.0));
}

// KC > 0: the trials are issued the way chains_fast.cuh does it -- 128 threads per active cell, every thread keeps the
// packed 8-bit frame coordinates of its KC candidates in registers, one sphere test (midpoint of old/new, radius rc +
// |delta|/2; VABSDIFF4 + IDP.4A + funnel shift) per candidate, survivors compacted with one warp prefix sum, fp64 only
// for survivors (no minimum image in the cell frame).  Needs cubic cells (one fixed-point scale).  KC == 0: every
// candidate in fp64 (non-cubic cells, prefilter = -1).
// The one-warp-per-trial speculative scheme of chains_spec.cuh was tried here too and measured 17 % SLOWER (5.3e8 vs
// 6.3e8 moves/s at N = 2^20): the trials of a cell nearly always conflict (filter sphere > cell).
template <int DIM, int MODEL, int KC>
__global__ void __launch_bounds__(kBfThreads, 8) k_box_sweep_all(const __grid_constant__ BoxArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Stencil<DIM> st;
    __shared__ int s_unit;
    __shared__ unsigned int s_stat[2];  // pmc_work_counters: fp64-evaluated candidates, trials evaluated
    constexpr bool FAST = KC > 0;
    constexpr int NST = Stencil<DIM>::NST;
    const int CAP = FAST ? kBfThreads * KC : A.cap;
    const BfLayout F = bf_layout(DIM, CAP);
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem_raw);
    // keep the shared-window base in a register: left alone, the compiler re-derives it (S2R SR_CgaCtaId + LEA, tens of
    // cycles of latency) three times per trial, once right in front of the survivor loads
    asm volatile("" : "+r"(sb));
    const uint32_t cap8 = 8u * (uint32_t)CAP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ptid = 32 * (lane >> 3) + 8 * warp + (lane & 7);  // this thread owns candidates ptid, ptid + 128, ... (load_stencil)
    double *sr = (double *)(smem_raw + F.r);
    uint8_t *ssp = smem_raw + F.sp;
    const double *spar = (const double *)(smem_raw + F.par);
    {
        double *wpar = (double *)(smem_raw + F.par), *scp = (double *)(smem_raw + F.cp);
        for (int k = tid; k < A.ns * A.ns * PMC_NPAR; k += kBfThreads) wpar[k] = A.par[k];
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
    const double cs0 = A.g.cs[0];
    const double fscale = 1073741824.0 / cs0;  // 2^30 / cell side (FAST: cubic cells)
    const uint32_t k0 = (uint32_t)A.seed, k1 = (uint32_t)(A.seed >> 32);
    const uint32_t qa = sb + F.q + (uint32_t)warp * (2u * (FAST ? KC : 1) * 32);
    const size_t ns_ = (size_t)A.g.nc[0] * A.plane_cap;
    const int n_units = (1 << DIM) * A.cell_n;
    if (threadIdx.x < 2) s_stat[threadIdx.x] = 0u;
    uint32_t rebuilt_ok = 0;  // bit p: peer p has finished this sweep's rebuild (its arrays may be written into)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(A.work, 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= n_units) break;
        const int phase = unit / A.cell_n, colour = A.order[phase];
        int cc[3] = {0, 0, 0};
        {
            // within a phase the cells are taken from both ends of this rank's slab towards its middle: the cells other
            // ranks wait for go first, and the cells that wait for other ranks find their flags long set
            const int idx = unit % A.cell_n;
            int l = (idx & 1) ? A.cell_lo + A.cell_n - 1 - (idx >> 1) : A.cell_lo + (idx >> 1);
            if constexpr (DIM == 3) { cc[2] = 2 * (l % (A.g.nc[2] / 2)) + ((colour >> 2) & 1); l /= (A.g.nc[2] / 2); }
            cc[1] = 2 * (l % (A.g.nc[1] / 2)) + ((colour >> 1) & 1);
            cc[0] = 2 * (l / (A.g.nc[1] / 2)) + (colour & 1);
        }
        // ---- wait for the neighbour cells of earlier phases (they may be another CTA's, or another GPU's) ----
        if (phase > 0) {
            if (tid >= 1 && tid < NST) {
                int k = tid <= NST / 2 ? tid - 1 : tid, c[3] = {0, 0, 0}, col = 0;
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    int v = cc[a] + k % 3 - 1;
                    k /= 3;
                    if (v < 0) v += A.g.nc[a];
                    if (v >= A.g.nc[a]) v -= A.g.nc[a];
                    c[a] = v;
                    col |= (v & 1) << a;
                }
                if (A.phase_of[col] < phase) wait_for(A.done + lin_cell<DIM>(c, A.g.nc), A.stamp, A.error, false);
            }
            __syncthreads();
        }
        uint32_t myq[FAST ? KC : 1];  // packed 8-bit prefilter coordinates of this thread's candidates
#pragma unroll
        for (int k = 0; k < (FAST ? KC : 1); k++) myq[k] = 0x80000000u;  // empty slot: fourth byte 128 never passes the sphere test
        const int ncand = load_stencil<DIM, kBfThreads>(A, cc, &st, sr, ssp, smem_raw + F.mv, CAP, [&](int t, uint32_t q) {
            if constexpr (FAST) {
#pragma unroll
                for (int k = 0; k < KC; k++)
                    if (t / kBfThreads == k) myq[k] = q;
            }
        });
        const int cell = st.cell[0];
        double Esum = 0.0;
        uint32_t nacc = 0;
        if (ncand >= 0) {
            const int ncen = st.off[1];
            for (int k = tid; k < ncen; k += kBfThreads) smem_raw[F.mv + k] = 0;
            uint32_t slot = 0;
            for (int tb = 0; tb < ncen; tb += kBfBatch) {  // n_cell trials in this cell (one sweep = N trials)
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
                    bool inside = xn[0] >= 0.0 && xn[0] < A.g.cs[0] && xn[1] >= 0.0 && xn[1] < A.g.cs[1];
                    if constexpr (DIM == 3) inside = inside && xn[2] >= 0.0 && xn[2] < A.g.cs[2];
                    if (!inside) continue;  // leaves the cell: rejected (uniform across the CTA)
                    const uint32_t si = bf_lds_u8(sb + F.sp + (uint32_t)k);
                    const uint32_t prow = si * (uint32_t)A.ns;
                    double part = 0.0;
                    auto pair_terms = [&](uint32_t j) {
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
                            const double *p = spar + (prow + sj) * PMC_NPAR;
                            const double rc2 = p[PMC_P_RCUT2];
                            if (r2o <= rc2) part -= pair_potential<MODEL>(p, r2o);
                            if (r2n <= rc2) part += pair_potential<MODEL>(p, r2n);
                        }
                    };
                    if constexpr (FAST) {
                        const uint32_t um0 = bf_fixed(xo[0], cs0, fscale) + (uint32_t)(di0 >> 1);
                        const uint32_t um1 = bf_fixed(xo[1], cs0, fscale) + (uint32_t)(di1 >> 1);
                        const uint32_t um2 = DIM == 3 ? bf_fixed(xo[2], cs0, fscale) + (uint32_t)(di2 >> 1) : 0u;
                        const int fthr = (int)bf_lds_u32(ra + 64 + 4u * si);
                        const uint32_t umq = pack8(um0, um1, um2);
                        uint32_t m = 0;
#pragma unroll
                        for (int kk = 0; kk < KC; kk++) {  // survivor: bit KC-1-kk
                            const uint32_t v = __vabsdiffu4(umq, myq[kk]);
                            m = __funnelshift_l((uint32_t)__dp4a((int)v, (int)v, fthr), m, 1);
                        }
                        // exclusive prefix of the per-lane survivor counts (<= KC each) from one ballot per count bit: the
                        // ballots are independent, where a shuffle scan is five dependent steps on the path of every trial
                        if (ptid == (k & (kBfThreads - 1))) m &= ~(1u << (KC - 1 - k / kBfThreads));  // the moved particle itself
                        const int mine = __popc(m);
                        constexpr int CB = KC < 4 ? 2 : (KC < 8 ? 3 : 4);  // bits of a count <= KC
                        const uint32_t lt = (1u << lane) - 1u;
                        int before = 0, total = 0;
#pragma unroll
                        for (int b = 0; b < CB; b++) {
                            const uint32_t bal = __ballot_sync(0xffffffffu, (mine >> b) & 1);
                            before += __popc(bal & lt) << b;
                            total += __popc(bal) << b;
                        }
                        const int incl = before + mine;
                        if (A.stats && lane == 0) atomicAdd(&s_stat[0], (unsigned int)total);
                        if (A.stats && tid == 0) atomicAdd(&s_stat[1], 1u);
                        uint32_t wp = qa + 2u * (uint32_t)(incl - mine);
#pragma unroll
                        for (int kk = 0; kk < KC; kk++) {
                            if (m & (1u << (KC - 1 - kk))) {
                                bf_sts_u16(wp, (uint32_t)(kk * kBfThreads + ptid));
                                wp += 2;
                            }
                        }
                        __syncwarp();
                        // (empty slots and the moved particle left at the mask level: no test per survivor)
                        for (int q = lane; q < total; q += 32) pair_terms(bf_lds_u16(qa + 2u * (uint32_t)q));
                        __syncwarp();
                    } else {
                        if (A.stats && tid == 0) {
                            atomicAdd(&s_stat[0], (unsigned int)(ncand - 1));
                            atomicAdd(&s_stat[1], 1u);
                        }
                        for (int j = tid; j < ncand; j += kBfThreads)
                            if (j != k) pair_terms((uint32_t)j);
                    }
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
                        if constexpr (FAST) {
                            if (ptid == (k & (kBfThreads - 1))) {  // owner refreshes its register copy
                                const int ki = k / kBfThreads;
                                const uint32_t f = pack8(bf_fixed(xn[0], cs0, fscale), bf_fixed(xn[1], cs0, fscale), DIM == 3 ? bf_fixed(xn[2], cs0, fscale) : 0u);
#pragma unroll
                                for (int kk = 0; kk < KC; kk++) myq[kk] = kk == ki ? f : myq[kk];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- write moved particles back: canonical arrays (+ image counters) and the sorted copy, here and on the peers ----
            const int bstart = st.base[0];
            if (A.n_peers) {  // a peer's arrays may only be written once it has rebuilt its cell lists for this sweep
                if (tid < A.n_peers && !(rebuilt_ok & (1u << tid))) wait_for(A.flags + kFlagRebuilt + A.peer_rank[tid], A.stamp, A.error, true);
                __syncthreads();
                rebuilt_ok = 0xFFu;
            }
            uint32_t keeps = 0;  // peers that keep the plane of this cell (they need the sorted copy and done[])
            for (int p = 0; p < A.n_peers; p++)
                if (plane_kept(cc[0], A.peer_plane_lo[p], A.peer_plane_n[p], A.g.nc[0])) keeps |= 1u << p;
            for (int k = tid; k < ncen; k += kBfThreads) {
                if (!smem_raw[F.mv + k]) continue;
                const int i = A.ids[bstart + k];
                uint32_t q = 0;
#pragma unroll
                for (int a = 0; a < DIM; a++) {
                    const double rnew = sr[a * CAP + k];
                    const double r0 = __ldcg(A.rs + (size_t)a * ns_ + bstart + k);
                    int w;
                    const double xnew = wrap1(__ldcg(A.x + (size_t)a * A.N + i) + (rnew - r0), A.g.L[a], w);
                    const int im = __ldcg(A.img + (size_t)a * A.N + i) + w;
                    A.rs[(size_t)a * ns_ + bstart + k] = rnew;
                    A.x[(size_t)a * A.N + i] = xnew;
                    if (w) A.img[(size_t)a * A.N + i] = im;
                    q |= cell_byte(rnew, 64.0 / A.g.cs[a]) << (8 * a);
                    for (int p = 0; p < A.n_peers; p++) {
                        A.peer_x[p][(size_t)a * A.N + i] = xnew;
                        if (w) A.peer_img[p][(size_t)a * A.N + i] = im;
                        if (keeps & (1u << p)) A.peer_rs[p][(size_t)a * ns_ + bstart + k] = rnew;
                    }
                }
                A.qs[bstart + k] = q;
                for (int p = 0; p < A.n_peers; p++)
                    if (keeps & (1u << p)) A.peer_qs[p][bstart + k] = q;
            }
            if (tid == 0) {
                A.cellE[cell] = Esum;
                A.cell_acc[cell] = nacc;
                for (int p = 0; p < A.n_peers; p++) {
                    A.peer_cellE[p][cell] = Esum;
                    A.peer_cacc[p][cell] = nacc;
                }
            }
            // ---- publish: this cell is done for this sweep (system-wide only if a peer keeps its plane; the pushes into the
            // canonical state of the other peers are fenced once, when the CTA leaves) ----
            // (the barrier orders every thread's stores before thread 0's fence, which makes them visible -- system-wide if
            // a peer keeps the plane -- before the stamp: one fence per cell instead of one per thread)
            __syncthreads();
            if (tid == 0) {
                if (keeps) __threadfence_system(); else __threadfence();
                *(volatile uint32_t *)(A.done + cell) = A.stamp;
                for (int p = 0; p < A.n_peers; p++)
                    if (keeps & (1u << p)) *(volatile uint32_t *)(A.peer_done[p] + cell) = A.stamp;
            }
        } else if (tid == 0) {
            *(volatile uint32_t *)(A.done + cell) = A.stamp;  // overflow: reported by the host, nobody may hang on this cell
            for (int p = 0; p < A.n_peers; p++) *(volatile uint32_t *)(A.peer_done[p] + cell) = A.stamp;
        }
    }
    if (A.stats) {
        __syncthreads();
        if (tid < 2) atomicAdd(A.stats + tid, (unsigned long long)s_stat[tid]);
    }
    // the CTA that leaves last tells every peer that this rank's sweep (all its pushes) is complete
    if (A.n_peers) {
        __syncthreads();
        if (tid == 0) __threadfence_system();  // this CTA's pushes (ordered before by the barrier) are performed system-wide
        if (tid == 0 && atomicAdd(A.work + 1, 1) == (int)gridDim.x - 1) {
            __threadfence_system();
            for (int p = 0; p < A.n_peers; p++) *(volatile uint32_t *)(A.peer_flags[p] + kFlagSwept + A.rank) = A.stamp;
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
    const int ncand = load_stencil<DIM, kBoxThreads>(A, cc, &st, sr, ssp, (uint8_t *)sid, A.cap, [](int, uint32_t) {});
    if (ncand < 0) return;
    // particle ids of the candidates (to count every unordered pair once: only id_i < id_j)
    for (int s = 0; s < Stencil<DIM>::NST; s++) {
        const int b = st.base[s], n = st.cnt[s];
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

// Deterministic reduction of per-cell values: out[0] (+)= scale * sum(cellE), acc[0] += sum(cell_acc).  Every CTA sums a
// fixed chunk of 1024 cells in a fixed order; the CTA that finishes last adds the partial sums in chunk order, so the
// result does not depend on the number of GPUs or on timing.  With peers, waits first until every rank has finished the
// sweep (their per-cell values have landed here).
__global__ void __launch_bounds__(1024) k_box_reduce(const __grid_constant__ BoxArgs A, uint32_t wait_stamp, int with_acc, double scale,
                                                     int accumulate, double *partE, unsigned long long *partA, int *ticket, double *outE,
                                                     unsigned long long *out_acc) {
    __shared__ double s_e[32];
    __shared__ unsigned long long s_a[32];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (A.n_peers && wait_stamp) {
        if (tid < A.n_peers) wait_for(A.flags + kFlagSwept + A.peer_rank[tid], wait_stamp, A.error, true);
        __syncthreads();
    }
    const int k = blockIdx.x * 1024 + tid;
    double e = k < A.g.ncell ? __ldcg(A.cellE + k) : 0.0;
    unsigned long long a = (with_acc && k < A.g.ncell) ? __ldcg(A.cell_acc + k) : 0ull;
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
        for (int w = 0; w < 32; w++) {
            te += s_e[w];
            ta += s_a[w];
        }
        partE[blockIdx.x] = te;
        partA[blockIdx.x] = ta;
        __threadfence();
        s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        double te = 0.0;
        unsigned long long ta = 0;
        for (int b = 0; b < (int)gridDim.x; b++) {
            te += __ldcg(partE + b);
            ta += __ldcg(partA + b);
        }
        outE[0] = (accumulate ? outE[0] : 0.0) + scale * te;
        if (out_acc) out_acc[0] += ta;
        *ticket = 0;
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
    uint32_t sweep = 0;   // sweeps since pmc_seed: RNG counter
    uint32_t stamp = 0;   // sweeps since creation: never reset, tags done[] and the inter-GPU flags
    bool geom_ready = false, model_ready = false;
    // canonical state + everything a peer GPU writes into: ONE allocation (one IPC handle, fixed offsets on every rank)
    unsigned char *shared_block = nullptr;
    size_t shared_bytes = 0;
    double *x = nullptr, *rs = nullptr, *cellE = nullptr;
    int32_t *img = nullptr;
    uint32_t *qs = nullptr, *cell_acc = nullptr, *done = nullptr, *bar_flags = nullptr;
    // local only
    double *par = nullptr, *eloc = nullptr, *energy = nullptr, *etmp = nullptr, *partE = nullptr;
    unsigned long long *partA = nullptr;
    int32_t *ids = nullptr, *start = nullptr, *cursor = nullptr, *count = nullptr, *cid = nullptr;
    uint8_t *sp = nullptr, *sps = nullptr;
    unsigned long long *acc_total = nullptr;
    int *flags = nullptr;  // [0] bad input (1, 2) / peer timeout (3), [1] overflow (1 stencil, 2 plane)
    int *work = nullptr;   // [0] next unit, [1] finished units, [2] finalize ticket, [3] reduce ticket
    double *raw = nullptr;
    long long *rsp = nullptr;
    int64_t calls = 0;
    int64_t launches = 0;
    unsigned long long *stats = nullptr;  // device work counters (owned by the context), nullptr: off
    size_t smem = 0;        // energy / histogram kernels
    int fast_kc = 0;        // > 0: k_box_sweep_all<.., KC> with the prefilter
    size_t sweep_smem = 0;
    int sweep_grid = 0;     // persistent CTAs of the sweep kernel
    int plane_cap = 0, ncell_alloc = 0;
    // multi-GPU
    int rank = 0, world = 1;
    unsigned char *peer_block[PMC_MAX_PEERS + 1] = {};  // opened IPC mapping of every rank's block (self = local)
    // PMC_BOX_TIMING=1 in the environment: device time of the three parts of every sweep (CUDA events), printed at destroy
    bool timing = false;
    std::vector<cudaEvent_t> tev;
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
    size_t x, img, rs, qs, cellE, cacc, done, flags, total;
};
BlockLayout block_layout(int N, int dim, int ncell, size_t nslot) {
    BlockLayout l;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t p = o;
        o += (bytes + 255) & ~(size_t)255;
        return p;
    };
    l.x = take(sizeof(double) * dim * (size_t)N);
    l.img = take(sizeof(int32_t) * dim * (size_t)N);
    l.rs = take(sizeof(double) * dim * nslot);
    l.qs = take(sizeof(uint32_t) * nslot);
    l.cellE = take(sizeof(double) * (size_t)ncell);
    l.cacc = take(sizeof(uint32_t) * (size_t)ncell);
    l.done = take(sizeof(uint32_t) * (size_t)ncell);
    l.flags = take(sizeof(uint32_t) * kFlagWords);
    l.total = o;
    return l;
}

int active_cells(const BoxState *b) {
    int n = 1;
    for (int a = 0; a < b->dim; a++) n *= b->g.nc[a] / 2;
    return n;
}

// Rank r of `world` sweeps the active cells [lo, hi) of every colour (x-major order: a slab of the box, cut to
// balance) and keeps cell lists for the planes those cells and their stencils touch.
void rank_share(const BoxState *b, int r, int world, int &lo, int &hi, int &plane_lo, int &plane_n) {
    const int nactive = active_cells(b), nx = b->g.nc[0];
    lo = (int)((int64_t)nactive * r / world);
    hi = (int)((int64_t)nactive * (r + 1) / world);
    if (world == 1 || hi <= lo) {
        plane_lo = 0;
        plane_n = nx;
        return;
    }
    const int app = nactive / (nx / 2);  // active cells per x-plane of active cells
    const int ax_lo = lo / app, ax_hi = (hi - 1) / app;
    plane_lo = 2 * ax_lo - 1;          // active x-coordinate 2h or 2h+1, stencil +-1
    plane_n = 2 * (ax_hi - ax_lo) + 4;
    if (plane_n >= nx) {
        plane_lo = 0;
        plane_n = nx;
    } else if (plane_lo < 0) {
        plane_lo += nx;
    }
}

void fill_args(BoxState *b, BoxArgs &a) {
    memset(&a, 0, sizeof a);
    a.g = b->g;
    a.N = b->N;
    a.ns = b->ns;
    a.cap = b->cap;
    a.x = b->x;
    a.img = b->img;
    a.sp = b->sp;
    a.plane_cap = b->plane_cap;
    a.rs = b->rs;
    a.qs = b->qs;
    a.sps = b->sps;
    a.ids = b->ids;
    a.start = b->start;
    a.count = b->count;
    a.par = b->par;
    a.T = b->T;
    a.sigma = (float)b->sigma;
    a.seed = b->seed;
    a.sweep = b->sweep;
    a.cellE = b->cellE;
    a.cell_acc = b->cell_acc;
    a.eloc = b->eloc;
    a.error = b->flags;
    a.stats = b->stats;
    a.overflow = b->flags + 1;
    a.stamp = b->stamp;
    a.done = b->done;
    a.work = b->work;
    a.rank = b->rank;
    a.flags = b->bar_flags;
    int lo, hi;
    rank_share(b, b->rank, b->world, lo, hi, a.plane_lo, a.plane_n);
    a.cell_lo = lo;
    a.cell_n = hi - lo;
    a.n_peers = 0;
    if (b->world > 1) {
        const size_t nslot = (size_t)b->g.nc[0] * b->plane_cap;
        const BlockLayout bl = block_layout(b->N, b->dim, b->g.ncell, nslot);
        for (int r = 0; r < b->world; r++) {
            if (r == b->rank) continue;
            const int p = a.n_peers++;
            unsigned char *blk = b->peer_block[r];
            int plo, phi;
            rank_share(b, r, b->world, plo, phi, a.peer_plane_lo[p], a.peer_plane_n[p]);
            a.peer_rank[p] = r;
            a.peer_x[p] = (double *)(blk + bl.x);
            a.peer_img[p] = (int32_t *)(blk + bl.img);
            a.peer_rs[p] = (double *)(blk + bl.rs);
            a.peer_qs[p] = (uint32_t *)(blk + bl.qs);
            a.peer_cellE[p] = (double *)(blk + bl.cellE);
            a.peer_cacc[p] = (uint32_t *)(blk + bl.cacc);
            a.peer_done[p] = (uint32_t *)(blk + bl.done);
            a.peer_flags[p] = (uint32_t *)(blk + bl.flags);
        }
    }
}

// K1: (re)build the cell lists of this rank's planes for grid origin g.shift.  whole_box: every plane (energies and
// histograms look at all cells); wait_stamp / announce_stamp: the inter-GPU handshake around a sweep's rebuild.
int build_cells(BoxState *b, bool whole_box, uint32_t wait_stamp, uint32_t announce_stamp) {
    const int N = b->N, nb = (N + 255) / 256;
    BoxArgs A;
    fill_args(b, A);
    if (whole_box) {
        A.plane_lo = 0;
        A.plane_n = b->g.nc[0];
    }
    BCU(cudaMemsetAsync(b->count, 0, sizeof(int32_t) * b->g.ncell, b->stream));
    if (b->dim == 3)
        k_box_count<3><<<nb, 256, 0, b->stream>>>(A, b->cid, wait_stamp);
    else
        k_box_count<2><<<nb, 256, 0, b->stream>>>(A, b->cid, wait_stamp);
    k_box_scan_plane<<<A.plane_n, 1024, 0, b->stream>>>(A, b->cursor);
    k_box_scatter<<<nb, 256, 0, b->stream>>>(b->cid, b->start, b->cursor, N, b->ids, b->flags + 1);
    const int cells = A.plane_n * (b->g.ncell / b->g.nc[0]);
    const int nbc = (cells + 7) / 8;  // one warp per cell, 8 warps per CTA
    if (b->dim == 3)
        k_box_finalize<3><<<nbc, 256, 0, b->stream>>>(A, b->work + 2, announce_stamp);
    else
        k_box_finalize<2><<<nbc, 256, 0, b->stream>>>(A, b->work + 2, announce_stamp);
    BCU(cudaGetLastError());
    b->launches += 4;
    return PMC_OK;
}

int setup_geometry(BoxState *b, const double *box3) {
    if (!b->model_ready) return bfail(PMC_ERR_STATE, "pmc_set_model must precede pmc_upload in PMC_MODE_BOX");
    Geom g{};
    g.ncell = 1;
    for (int a = 0; a < 3; a++) {
        g.nc[a] = 1;
        g.L[a] = 1.0;
        g.cs[a] = 1.0;
        g.shift[a] = 0.0;
    }
    double occ = (double)b->N;
    for (int a = 0; a < b->dim; a++) {
        // cells of side >= rcut_max (src/neighbours.jl:236-238), count rounded down to even for the colouring
        int n = (int)std::floor(box3[a] / b->rcut_max);
        n -= n % 2;
        if (n < 2)
            return bfail(PMC_ERR_UNSUPPORTED, "box side %g holds fewer than 2 cells of side >= rcut %g: use PMC_MODE_CHAINS",
                         box3[a], b->rcut_max);
        g.nc[a] = n;
        g.L[a] = box3[a];
        g.cs[a] = box3[a] / (double)n;
        g.ncell *= n;
    }
    if (b->world > 1 && (g.ncell != b->g.ncell || memcmp(g.nc, b->g.nc, sizeof g.nc) != 0))
        return bfail(PMC_ERR_STATE, "the cell grid changed after peers were attached");
    b->g = g;
    occ /= (double)b->g.ncell;
    const int nst = b->dim == 3 ? 27 : 9;
    int cap = (int)(occ * nst * 1.5) + 96;
    cap = (cap + 31) / 32 * 32;
    b->cap = cap;
    b->smem = sizeof(double) * (size_t)b->dim * cap + 2 * (size_t)cap + 16;
    if (b->smem > 200 * 1024) return bfail(PMC_ERR_UNSUPPORTED, "stencil of %d candidates does not fit shared memory", cap);
    // prefilter kernel: cubic cells, stencil fits kBfThreads x KC register candidates
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
        b->smem = sizeof(double) * (size_t)b->dim * b->cap + 2 * (size_t)b->cap + 16;
    }
    b->sweep_smem = bf_layout(b->dim, b->cap).total;
    int sms = 0;
    BCU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->cfg.device));
    int rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
        constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
        int occb = 0;
        auto prep = [&](auto kernel) {
            BCU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->sweep_smem));
            BCU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occb, kernel, kBfThreads, b->sweep_smem));
            return (int)PMC_OK;
        };
        int r2;
        if (b->fast_kc == kBfKc0) r2 = prep(k_box_sweep_all<d, mdl, kBfKc0>);
        else if (b->fast_kc == kBfKc1) r2 = prep(k_box_sweep_all<d, mdl, kBfKc1>);
        else if (b->fast_kc == kBfKc2) r2 = prep(k_box_sweep_all<d, mdl, kBfKc2>);
        else if (b->fast_kc == kBfKc3) r2 = prep(k_box_sweep_all<d, mdl, kBfKc3>);
        else r2 = prep(k_box_sweep_all<d, mdl, 0>);
        if (r2) return r2;
        if (occb < 1) return bfail(PMC_ERR_UNSUPPORTED, "the sweep kernel does not fit an SM (%zu B of shared memory)", b->sweep_smem);
        b->sweep_grid = sms * occb;
        BCU(cudaFuncSetAttribute(k_box_energy<d, mdl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
        return (int)PMC_OK;
    });
    if (rc) return rc;
    if (b->ncell_alloc != b->g.ncell) {  // buffers survive re-uploads of the same geometry (peers map them)
        for (void *p : {(void *)b->count, (void *)b->cursor, (void *)b->start, (void *)b->shared_block, (void *)b->ids,
                        (void *)b->sps, (void *)b->partE, (void *)b->partA})
            if (p) cudaFree(p);
        b->count = b->cursor = b->start = b->ids = nullptr;
        b->shared_block = nullptr;
        b->sps = nullptr;
        b->partE = nullptr;
        b->partA = nullptr;
        // sorted slots are reserved per x-plane of cells: 1.5 x the mean plane occupancy (a lattice start puts 3 or 4
        // lattice planes into a plane of cells); overflow is detected by the scan and reported
        b->plane_cap = ((int)((double)b->N / b->g.nc[0] * 1.5) + 64 + 31) / 32 * 32;
        const size_t nslot = (size_t)b->g.nc[0] * b->plane_cap;
        BCU(balloc(&b->count, b->g.ncell));
        BCU(balloc(&b->cursor, b->g.ncell));
        BCU(balloc(&b->start, b->g.ncell + 1));
        BCU(balloc(&b->ids, nslot));
        BCU(balloc(&b->sps, nslot));
        BCU(balloc(&b->partE, (b->g.ncell + 1023) / 1024));
        BCU(balloc(&b->partA, (b->g.ncell + 1023) / 1024));
        const BlockLayout bl = block_layout(b->N, b->dim, b->g.ncell, nslot);
        BCU(balloc(&b->shared_block, bl.total));
        b->shared_bytes = bl.total;
        b->x = (double *)(b->shared_block + bl.x);
        b->img = (int32_t *)(b->shared_block + bl.img);
        b->rs = (double *)(b->shared_block + bl.rs);
        b->qs = (uint32_t *)(b->shared_block + bl.qs);
        b->cellE = (double *)(b->shared_block + bl.cellE);
        b->cell_acc = (uint32_t *)(b->shared_block + bl.cacc);
        b->done = (uint32_t *)(b->shared_block + bl.done);
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
    if (fl[1] == 2) return bfail(PMC_ERR_UNSUPPORTED, "an x-plane of cells holds more than %d particles (density too inhomogeneous)", b->plane_cap);
    if (fl[1]) return bfail(PMC_ERR_UNSUPPORTED, "a 3^d-cell neighbourhood holds more than %d particles (density too inhomogeneous)", b->cap);
    if (fl[0] == 3) return bfail(PMC_ERR_CUDA, "inter-GPU handshake timed out: a peer rank did not arrive");
    return PMC_OK;
}

int reduce_cells(BoxState *b, const BoxArgs &A, uint32_t wait_stamp, bool sweep) {
    const int nb = (b->g.ncell + 1023) / 1024;
    if (sweep)
        k_box_reduce<<<nb, 1024, 0, b->stream>>>(A, wait_stamp, 1, 1.0, 1, b->partE, b->partA, b->work + 3, b->energy, b->acc_total);
    else
        k_box_reduce<<<nb, 1024, 0, b->stream>>>(A, 0u, 0, 0.5, 0, b->partE, b->partA, b->work + 3, b->etmp, nullptr);
    BCU(cudaGetLastError());
    b->launches++;
    return PMC_OK;
}

// local energies (grid origin 0) -> eloc in particle order, etmp[0] = sum/2
int compute_energy(BoxState *b) {
    for (int a = 0; a < 3; a++) b->g.shift[a] = 0.0;
    int rc = build_cells(b, true, 0u, 0u);
    if (rc) return rc;
    BoxArgs A;
    fill_args(b, A);
    rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
        k_box_energy<decltype(D)::value, decltype(MDL)::value><<<b->g.ncell, kBoxThreads, b->smem, b->stream>>>(A);
        BCU(cudaGetLastError());
        return (int)PMC_OK;
    });
    if (rc) return rc;
    b->launches++;
    rc = reduce_cells(b, A, 0u, false);
    if (rc) return rc;
    return check_overflow(b);
}

}  // namespace

const char *box_error() { return g_box_err.c_str(); }

int box_check(BoxState *b) { return check_overflow(b); }

int box_pair_histogram(BoxState *b, int sa, int sb, double rmax, int nbins, unsigned long long *d_hist) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    for (int a = 0; a < b->dim; a++)
        if (rmax > b->g.cs[a]) return bfail(PMC_ERR_INVALID, "rmax %g exceeds the cell side %g of the device cell list", rmax, b->g.cs[a]);
    for (int a = 0; a < 3; a++) b->g.shift[a] = 0.0;
    int rc = build_cells(b, true, 0u, 0u);
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
    if (b->stamp != 0) return bfail(PMC_ERR_STATE, "attach the peers before the first sweep (all ranks count sweeps alike)");
    if (world > active_cells(b) / 2)
        return bfail(PMC_ERR_UNSUPPORTED, "%d ranks for %d active cells per colour: the box is too small to split", world, active_cells(b));
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
    }
    b->rank = rank;
    b->world = world;
    return PMC_OK;
}

int box_create(BoxState **out, const pmc_config &cfg) {
    if (cfg.n_chains != 1) return bfail(PMC_ERR_INVALID, "PMC_MODE_BOX holds exactly one system (n_chains = %d)", cfg.n_chains);
    if (cfg.molecules) return bfail(PMC_ERR_UNSUPPORTED, "Molecules are not supported in PMC_MODE_BOX");
    BoxState *b = new BoxState();
    b->cfg = cfg;
    b->timing = std::getenv("PMC_BOX_TIMING") != nullptr;
    b->N = cfg.n_particles;
    b->dim = cfg.dim;
    b->ns = cfg.n_species;
    const size_t N = b->N, d = b->dim;
    cudaError_t e = balloc(&b->cid, N);
    if (e == cudaSuccess) e = balloc(&b->sp, N);
    if (e == cudaSuccess) e = balloc(&b->eloc, N);
    if (e == cudaSuccess) e = balloc(&b->par, (size_t)PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR);
    if (e == cudaSuccess) e = balloc(&b->energy, 1);
    if (e == cudaSuccess) e = balloc(&b->etmp, 1);
    if (e == cudaSuccess) e = balloc(&b->acc_total, 1);
    if (e == cudaSuccess) e = balloc(&b->flags, 2);
    if (e == cudaSuccess) e = balloc(&b->work, 4);
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
    if (b->timing && b->tev.size() >= 4) {
        cudaDeviceSynchronize();
        double t[3] = {0, 0, 0};
        const size_t n = b->tev.size() / 4, skip = n / 2;  // second half of the run
        for (size_t s = skip; s < n; s++)
            for (int k = 0; k < 3; k++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, b->tev[4 * s + k], b->tev[4 * s + k + 1]);
                t[k] += ms;
            }
        fprintf(stderr, "[pmc box timing] rank %d/%d: %zu sweeps, per sweep: rebuild %.1f us, sweep kernel %.1f us, reduce (+ wait for peers) %.1f us\n",
                b->rank, b->world, n - skip, 1e3 * t[0] / (n - skip), 1e3 * t[1] / (n - skip), 1e3 * t[2] / (n - skip));
        for (auto e : b->tev) cudaEventDestroy(e);
    }
    for (int r = 0; r < b->world; r++)
        if (r != b->rank && b->peer_block[r]) cudaIpcCloseMemHandle(b->peer_block[r]);
    void *bufs[] = {b->shared_block, b->ids, b->cid, b->sp, b->sps, b->eloc, b->par, b->energy, b->etmp, b->acc_total,
                    b->flags, b->work, b->raw, b->rsp, b->count, b->cursor, b->start, b->partE, b->partA};
    for (void *p : bufs)
        if (p) cudaFree(p);
    delete b;
}

void box_set_stream(BoxState *b, cudaStream_t st) { b->stream = st; }
void box_set_sigma(BoxState *b, double sigma) { b->sigma = sigma; }
void box_set_stats(BoxState *b, unsigned long long *stats) { b->stats = stats; }
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

// One sweep = N trials: fresh random grid origin, rebuild this rank's cell lists, then ONE persistent kernel that runs
// the 2^d colours in a random order (dataflow between cells, no barriers), then the deterministic reduction of the per-
// cell energy changes.  n_trials is rounded up to whole sweeps.  Four small launches + one sweep kernel + one reduction
// per sweep; with peers the only inter-GPU waits are flag polls inside these kernels (rebuild: peers finished the previous
// sweep; first push into a peer: it finished its rebuild; cell: its neighbours; reduction: peers finished this sweep).
int box_run(BoxState *b, int64_t n_trials) {
    if (!b->geom_ready) return bfail(PMC_ERR_STATE, "nothing uploaded yet");
    const int64_t sweeps = (n_trials + b->N - 1) / b->N;
    const int ncol = 1 << b->dim;
    const uint32_t k0 = (uint32_t)b->seed, k1 = (uint32_t)(b->seed >> 32);
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
        const uint32_t prev = b->stamp;
        b->stamp++;
        auto mark = [&]() {
            if (!b->timing) return;
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, b->stream);
            b->tev.push_back(e);
        };
        mark();
        int rc = build_cells(b, false, prev, b->stamp);
        if (rc) return rc;
        BoxArgs A;
        fill_args(b, A);
        for (int k = 0; k < ncol; k++) {
            A.order[k] = order[k];
            A.phase_of[order[k]] = k;
        }
        BCU(cudaMemsetAsync(b->work, 0, 2 * sizeof(int), b->stream));
        mark();
        const int grid = std::min(b->sweep_grid, ncol * A.cell_n);
        if (grid > 0) {
            rc = bdispatch(b->dim, b->cfg.model_kind, [&](auto D, auto MDL) {
                constexpr int d = decltype(D)::value, mdl = decltype(MDL)::value;
                if (b->fast_kc == kBfKc0)
                    k_box_sweep_all<d, mdl, kBfKc0><<<grid, kBfThreads, b->sweep_smem, b->stream>>>(A);
                else if (b->fast_kc == kBfKc1)
                    k_box_sweep_all<d, mdl, kBfKc1><<<grid, kBfThreads, b->sweep_smem, b->stream>>>(A);
                else if (b->fast_kc == kBfKc2)
                    k_box_sweep_all<d, mdl, kBfKc2><<<grid, kBfThreads, b->sweep_smem, b->stream>>>(A);
                else if (b->fast_kc == kBfKc3)
                    k_box_sweep_all<d, mdl, kBfKc3><<<grid, kBfThreads, b->sweep_smem, b->stream>>>(A);
                else
                    k_box_sweep_all<d, mdl, 0><<<grid, kBfThreads, b->sweep_smem, b->stream>>>(A);
                BCU(cudaGetLastError());
                return (int)PMC_OK;
            });
            if (rc) return rc;
        }
        mark();
        rc = reduce_cells(b, A, b->stamp, true);
        if (rc) return rc;
        mark();
        b->launches += 1;
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
