// api.cu -- the C ABI of include/pmc_b200.h: context management, host<->device marshalling,
// kernel orchestration.  No torch, no Python: plain CUDA runtime behind extern "C".
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "box.cuh"
#include "chains.cuh"
#include "common.cuh"
#include "pmc_b200.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(PMC_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
cudaError_t dalloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, sizeof(T) * (n ? n : 1));
    return e;
}

// ---- layout kernels ------------------------------------------------------------------------------
// Ingest: caller layout (AoS float64 positions, int64 labels) -> device layout (wrapped SoA + image
// counters, uint8 species) and the SpeciesList of src/utils.jl:36-49 (ids ascending per species).
// One warp per chain for the (ordered) species lists, all threads for the transpose.
__global__ void k_ingest(const double *__restrict__ raw_pos, const long long *__restrict__ raw_sp,
                         const double *__restrict__ box, int first, int count, int N, int Npad, int dim, int ns,
                         double *x, int32_t *img, uint8_t *sp, uint16_t *spids, uint16_t *heads, int32_t *spoff,
                         int *bad) {
    const int lc = blockIdx.x;  // chain within this upload
    const int c = first + lc;
    const int tid = threadIdx.x, NT = blockDim.x;
    for (int k = tid; k < N * dim; k += NT) {
        const int i = k / dim, a = k % dim;
        const double L = box[c * 3 + a];
        const double v = raw_pos[((size_t)lc * N + i) * dim + a];
        const double n = floor(v / L);  // fold_back (src/utils.jl:12, src/IO/IO.jl:284)
        double w = v - n * L;
        int im = (int)n;
        if (w >= L) {
            w -= L;
            im += 1;
        }
        if (w < 0.0) {
            w += L;
            im -= 1;
        }
        if (!(w >= 0.0 && w <= L)) atomicExch(bad, 1);  // NaN / Inf coordinates
        x[((size_t)c * dim + a) * Npad + i] = w;
        img[((size_t)c * dim + a) * Npad + i] = im;
    }
    for (int i = tid; i < Npad; i += NT) {
        int s = 0;
        if (i < N) {
            const long long lab = raw_sp[(size_t)lc * N + i];
            if (lab < 1 || lab > ns) atomicExch(bad, 2);
            s = (int)(lab - 1);
        }
        sp[(size_t)c * Npad + i] = (uint8_t)s;
    }
    __syncthreads();
    if (tid < 32) {  // ordered compaction per species by one warp
        int off = 0;
        for (int s = 0; s < ns; s++) {
            if (tid == 0) spoff[c * (PMC_MAX_SPECIES + 1) + s] = off;
            int cnt = 0;
            for (int base = 0; base < N; base += 32) {
                const int i = base + tid;
                const bool f = (i < N) && sp[(size_t)c * Npad + i] == s;
                const unsigned m = __ballot_sync(0xffffffffu, f);
                if (f) {
                    const int pos = cnt + __popc(m & ((1u << tid) - 1u));
                    spids[(size_t)c * Npad + off + pos] = (uint16_t)i;
                    heads[(size_t)c * Npad + i] = (uint16_t)pos;
                }
                cnt += __popc(m);
            }
            off += cnt;
        }
        if (tid == 0)
            for (int s = ns; s <= PMC_MAX_SPECIES; s++) spoff[c * (PMC_MAX_SPECIES + 1) + s] = off;
    }
}

// Egress: device layout -> caller layout, positions unwrapped again (x + img * L).
__global__ void k_egress(const double *__restrict__ x, const int32_t *__restrict__ img,
                         const uint8_t *__restrict__ sp, const double *__restrict__ box, int first, int N, int Npad,
                         int dim, double *raw_pos, long long *raw_sp) {
    const int lc = blockIdx.x, c = first + lc;
    for (int k = threadIdx.x; k < N * dim; k += blockDim.x) {
        const int i = k / dim, a = k % dim;
        const size_t s = ((size_t)c * dim + a) * Npad + i;
        raw_pos[((size_t)lc * N + i) * dim + a] = x[s] + (double)img[s] * box[c * 3 + a];
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) raw_sp[(size_t)lc * N + i] = (long long)sp[(size_t)c * Npad + i] + 1;
}

// ---- observables of the species field --------------------------------------------------------------
// compute_chain_correlation (src/molecules.jl:224-243): monodisperse molecules of `len` sites; species 2 counts as
// -1, any other species as its label; for every site pair a < b the cross term sum_mol v_a v_b / Nmol, result =
// sum of the squared cross terms.  The per-pair sums are integers (exact); one CTA per chain, thread p owns pair p.
__global__ void k_chain_correlation(const uint8_t *__restrict__ sp, const int32_t *__restrict__ mol_start, int n_mol,
                                    int len, int Npad, double *out) {
    extern __shared__ double s_cross2[];
    const int c = blockIdx.x, npairs = len * (len - 1) / 2;
    const uint8_t *csp = sp + (size_t)c * Npad;
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        int a = 0, rem = p;  // pair index -> (a, b), pairs ordered (0,1), (0,2), ..., (1,2), ... as in the reference loop
        while (rem >= len - 1 - a) {
            rem -= len - 1 - a;
            a++;
        }
        const int b = a + 1 + rem;
        long long acc = 0;
        for (int m = 0; m < n_mol; m++) {
            const int s0 = mol_start[m];
            const int la = (int)csp[s0 + a] + 1, lb = (int)csp[s0 + b] + 1;  // 1-based labels
            acc += (long long)(la == 2 ? -1 : la) * (long long)(lb == 2 ? -1 : lb);
        }
        const double cross = (double)acc / (double)n_mol;
        s_cross2[p] = cross * cross;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int p = 0; p < npairs; p++) t += s_cross2[p];
        out[c] = t;
    }
}

// histogram of the running energies of all chains (per particle if per_n), nbins equal bins on [emin, emax)
__global__ void k_energy_histogram(const double *__restrict__ energy, int M, double inv_n, double emin, double inv_w, int nbins,
                                   unsigned long long *hist) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    const double t = (energy[c] * inv_n - emin) * inv_w;
    if (t >= 0.0 && t < (double)nbins) atomicAdd(&hist[(int)t], 1ull);
}

// ---- FMA burst micro-benchmark (roofline denominator) ---------------------------------------------
template <typename T>
__global__ void k_fma_burst(T *out, int iters) {
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999, c = (T)1e-3;
    for (int k = 0; k < iters; k++) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

struct pmc_ctx {
    pmc_config cfg{};
    int Npad = 0;
    int threads = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_run_events = false;
    // device state (chains mode)
    double *x = nullptr;
    int32_t *img = nullptr;
    uint8_t *sp = nullptr;
    uint16_t *spids = nullptr, *heads = nullptr;
    int32_t *spoff = nullptr;
    double *box = nullptr, *temp = nullptr, *energy = nullptr, *etot = nullptr, *eloc = nullptr, *par = nullptr;
    unsigned long long *calls = nullptr, *accepted = nullptr;
    uint16_t *bonds = nullptr;
    int32_t *mol_start = nullptr, *mol_len = nullptr;
    int n_mol = 0;
    int mol_uniform_len = 0;  // > 0: every molecule has this many sites (pmc_chain_correlation)
    int *bad = nullptr;
    int32_t *queue = nullptr;  // [1 + n_chains] work queue of the speculative kernel (chains_spec.cuh)
    unsigned long long *stats = nullptr;  // [4] pmc_work_counters
    bool stats_on = false;
    // staging
    double *raw_pos = nullptr;
    long long *raw_sp = nullptr;
    size_t raw_chains = 0;
    // host-side bookkeeping
    std::vector<pmc_move> pool;
    unsigned long long seed = 0, t0 = 0;
    bool model_set = false, uploaded = false, energy_set = false, bonds_set = false;
    int64_t launches = 0;
    size_t sweep_smem = 0, sweep_smem_filter = 0, energy_smem = 0, fast_smem = 0, spec_smem = 0;
    bool spec_swaps = false;
    bool sweep_swap_cfg = false;
    bool cubic = true;  // every uploaded chain has a cubic box (enables the fixed-point prefilter)
    double min_box = 1e300;  // shortest box length uploaded so far (proposal widths are checked against it)
    pmc::BoxState *boxst = nullptr;
};

namespace {

int ensure_staging(pmc_ctx *c, size_t chains) {
    if (c->raw_chains >= chains) return PMC_OK;
    if (c->raw_pos) cudaFree(c->raw_pos);
    if (c->raw_sp) cudaFree(c->raw_sp);
    c->raw_pos = nullptr;
    c->raw_sp = nullptr;
    c->raw_chains = 0;
    CU(cudaMalloc((void **)&c->raw_pos, sizeof(double) * chains * c->cfg.n_particles * c->cfg.dim));
    CU(cudaMalloc((void **)&c->raw_sp, sizeof(long long) * chains * c->cfg.n_particles));
    c->raw_chains = chains;
    return PMC_OK;
}

bool pool_has_swap(const pmc_ctx *c) {
    for (auto &m : c->pool)
        if (m.kind == PMC_MOVE_SWAP || m.kind == PMC_MOVE_FLIP) return true;
    return false;
}

int configure_sweep(pmc_ctx *c, bool any_swap) {
    const bool mol = c->cfg.molecules != 0;
    size_t s = pmc::chain_sweep_smem_bytes(c->cfg.dim, c->Npad, c->cfg.n_species, c->threads, mol, any_swap, false);
    size_t sf = pmc::chain_sweep_smem_bytes(c->cfg.dim, c->Npad, c->cfg.n_species, c->threads, mol, any_swap, true);
    size_t e = pmc::chain_energy_smem_bytes(c->cfg.dim, c->Npad, c->cfg.n_species, mol);
    int dev_max = 0;
    CU(cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->cfg.device));
    if (s > (size_t)dev_max || e > (size_t)dev_max)
        return fail(PMC_ERR_UNSUPPORTED,
                    "chain state needs %zu B of shared memory per CTA (limit %d): use PMC_MODE_BOX for N=%d", s > e ? s : e,
                    dev_max, c->cfg.n_particles);
    if (sf > (size_t)dev_max) sf = s;  // prefilter does not fit: use_filter() falls back to the direct kernel
    if (s != c->sweep_smem || sf != c->sweep_smem_filter || e != c->energy_smem) {
        CU(pmc::configure_chain_kernels(c->cfg.dim, c->cfg.model_kind, mol, s, sf, e));
        c->sweep_smem = s;
        c->sweep_smem_filter = sf;
        c->energy_smem = e;
    }
    return PMC_OK;
}

void fill_chain_args(pmc_ctx *c, pmc::ChainArgs &a, int64_t n_trials, bool any_swap) {
    memset(&a, 0, sizeof a);
    a.x = c->x;
    a.img = c->img;
    a.sp = c->sp;
    a.spids = c->spids;
    a.heads = c->heads;
    a.spoff = c->spoff;
    a.box = c->box;
    a.temp = c->temp;
    a.energy = c->energy;
    a.calls = c->calls;
    a.accepted = c->accepted;
    a.par = c->par;
    a.bonds = c->bonds;
    a.mol_start = c->mol_start;
    a.mol_len = c->mol_len;
    a.n_mol = c->n_mol;
    a.n_moves = (int)c->pool.size();
    double tot = 0.0, cum = 0.0;
    for (auto &m : c->pool) tot += m.probability;
    for (size_t k = 0; k < c->pool.size(); k++) {
        cum += c->pool[k].probability / tot;
        a.mv_kind[k] = c->pool[k].kind;
        a.mv_a[k] = c->pool[k].species_a - 1;
        a.mv_b[k] = c->pool[k].species_b - 1;
        a.mv_cum[k] = cum;
        a.mv_sigma[k] = (float)c->pool[k].sigma;
    }
    a.any_swap = any_swap ? 1 : 0;
    a.seed = c->seed;
    a.t0 = c->t0;
    a.chain_offset = c->cfg.chain_offset;
    a.N = c->cfg.n_particles;
    a.Npad = c->Npad;
    a.ns = c->cfg.n_species;
    a.n_trials = n_trials;
    a.stats = c->stats_on ? c->stats : nullptr;
}

int check_ready(pmc_ctx *c, bool need_moves) {
    if (!c) return fail(PMC_ERR_INVALID, "null context");
    if (!c->model_set) return fail(PMC_ERR_STATE, "pmc_set_model has not been called");
    if (!c->uploaded) return fail(PMC_ERR_STATE, "pmc_upload has not been called");
    if (c->cfg.molecules && !c->bonds_set) return fail(PMC_ERR_STATE, "pmc_set_bonds has not been called");
    if (need_moves) {
        if (c->pool.empty()) return fail(PMC_ERR_STATE, "pmc_set_moves has not been called");
        if (!c->energy_set) return fail(PMC_ERR_STATE, "pmc_init_energy has not been called");
    }
    return PMC_OK;
}

int run_energy(pmc_ctx *c) {
    int rc = configure_sweep(c, c->sweep_swap_cfg);
    if (rc) return rc;
    pmc::EnergyArgs e{};
    e.x = c->x;
    e.sp = c->sp;
    e.box = c->box;
    e.par = c->par;
    e.bonds = c->bonds;
    e.eloc = c->eloc;
    e.etot = c->etot;
    e.N = c->cfg.n_particles;
    e.Npad = c->Npad;
    e.ns = c->cfg.n_species;
    if (c->cubic && c->cfg.prefilter >= 0 && !c->cfg.molecules && c->Npad <= 1024)
        CU(pmc::launch_chain_energy_fast(c->cfg.dim, c->cfg.model_kind, c->cfg.n_chains, e, c->stream));
    else
        CU(pmc::launch_chain_energy(c->cfg.dim, c->cfg.model_kind, c->cfg.molecules != 0, c->cfg.n_chains, c->energy_smem,
                                    e, c->stream));
    c->launches++;
    return PMC_OK;
}

int sweep(pmc_ctx *c, int64_t n_trials, const pmc_trial *d_replay, pmc_trial *d_trace, uint8_t *d_acc, double *d_dE,
          bool exact_exp, bool replay_has_swap = false) {
    const bool any_swap = pool_has_swap(c) || replay_has_swap;
    c->sweep_swap_cfg = any_swap;
    int rc = configure_sweep(c, any_swap);
    if (rc) return rc;
    pmc::ChainArgs a;
    fill_chain_args(c, a, n_trials, any_swap);
    a.replay = d_replay;
    a.trace = d_trace;
    a.acc_out = d_acc;
    a.dE_out = d_dE;
    a.exact_exp = exact_exp ? 1 : 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    const bool filter = c->cubic && c->cfg.prefilter >= 0 && c->sweep_smem_filter > c->sweep_smem;
    bool flips = false;  // MoleculeFlip: the speculative kernel (Molecules) or the general one, never the one-trial-at-a-time kernel
    for (auto &m : c->pool) flips = flips || m.kind == PMC_MOVE_FLIP;
    const bool fastk = filter && !flips && !c->cfg.molecules && pmc::chain_fast_supported(c->cfg.dim, c->Npad, c->threads);
    const bool mol = c->cfg.molecules != 0, mixed = c->cfg.precision == PMC_MIXED;
    const bool speck = c->cubic && (!flips || mol) && c->cfg.prefilter == 0 &&
                       pmc::chain_spec_supported(c->cfg.dim, c->cfg.model_kind, c->Npad, c->threads, mol, mixed, any_swap);
    if (speck) {
        const size_t ss = pmc::chain_spec_smem_bytes(c->cfg.dim, c->Npad, c->cfg.model_kind, mixed, mol, any_swap);
        if (ss != c->spec_smem || any_swap != c->spec_swaps) {
            CU(pmc::configure_chain_spec(c->cfg.dim, c->cfg.model_kind, c->Npad, ss, mixed, mol, any_swap));
            c->spec_smem = ss;
            c->spec_swaps = any_swap;
        }
        CU(pmc::launch_chain_sweep_spec(c->cfg.dim, c->cfg.model_kind, c->cfg.n_chains, ss, a, c->stream, mixed, mol, any_swap, c->queue));
    } else if (mixed) {
        if (!fastk || any_swap)
            return fail(PMC_ERR_UNSUPPORTED, "PMC_MIXED needs a cubic box, a Displacement-only pool and %d threads per CTA", 128);
        CU(pmc::launch_chain_sweep_mixed(c->cfg.dim, c->cfg.model_kind, c->cfg.n_chains,
                                         pmc::chain_mixed_smem_bytes(c->cfg.dim, c->Npad), a, c->stream));
    } else if (fastk) {
        const size_t fs = pmc::chain_fast_smem_bytes(c->cfg.dim, c->Npad, c->cfg.model_kind, any_swap);
        if (fs != c->fast_smem) {
            CU(pmc::configure_chain_fast(c->cfg.dim, c->cfg.model_kind, c->Npad, any_swap, fs));
            c->fast_smem = fs;
        }
        CU(pmc::launch_chain_sweep_fast(c->cfg.dim, c->cfg.model_kind, c->cfg.n_chains, fs, a, c->stream));
    } else {
        CU(pmc::launch_chain_sweep(c->cfg.dim, c->cfg.model_kind, c->cfg.molecules != 0, filter, c->cfg.n_chains,
                                   c->threads, filter ? c->sweep_smem_filter : c->sweep_smem, a, c->stream));
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    c->have_run_events = true;
    c->launches++;
    if (!d_replay) c->t0 += (unsigned long long)n_trials;
    return PMC_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int pmc_abi_version(void) { return PMC_ABI_VERSION; }
const char *pmc_last_error(void) { return g_err.c_str(); }

int pmc_create(const pmc_config *cfg, pmc_ctx **out) {
    if (!cfg || !out) return fail(PMC_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dim != 2 && cfg->dim != 3) return fail(PMC_ERR_INVALID, "dim must be 2 or 3 (got %d)", cfg->dim);
    if (cfg->n_species < 1 || cfg->n_species > PMC_MAX_SPECIES)
        return fail(PMC_ERR_INVALID, "n_species must be in 1..%d (got %d)", PMC_MAX_SPECIES, cfg->n_species);
    if (cfg->model_kind < PMC_MODEL_LJ || cfg->model_kind > PMC_MODEL_KG)
        return fail(PMC_ERR_INVALID, "unknown model kind %d", cfg->model_kind);
    if (cfg->molecules && cfg->model_kind != PMC_MODEL_KG)
        return fail(PMC_ERR_INVALID, "Molecules require the GeneralKG model (bond_potential)");
    if (cfg->n_chains < 1 || cfg->n_particles < 1) return fail(PMC_ERR_INVALID, "n_chains and n_particles must be >= 1");
    if (cfg->precision != PMC_FP64 && cfg->precision != PMC_MIXED) return fail(PMC_ERR_INVALID, "unknown precision %d", cfg->precision);
    if (cfg->precision == PMC_MIXED && (cfg->mode != PMC_MODE_CHAINS || cfg->molecules || cfg->n_particles > 1024))
        return fail(PMC_ERR_UNSUPPORTED, "PMC_MIXED is implemented for PMC_MODE_CHAINS, Atoms, N <= 1024 (Displacement pools, cubic boxes)");
    if (cfg->mode != PMC_MODE_CHAINS && cfg->mode != PMC_MODE_BOX) return fail(PMC_ERR_INVALID, "unknown mode %d", cfg->mode);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PMC_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PMC_ERR_INVALID, "device %d out of range", cfg->device);
    CU(cudaSetDevice(cfg->device));
    pmc_ctx *c = new pmc_ctx();
    c->cfg = *cfg;
    c->Npad = (cfg->n_particles + 31) / 32 * 32;
    // default CTA size: 128 threads while a chain fits the register-resident kernels (several CTAs per SM), 256 for
    // large chains whose shared-memory footprint allows only one or two CTAs per SM
    c->threads = cfg->threads > 0 ? cfg->threads : (c->Npad > (cfg->dim == 2 ? 2048 : 1024) ? 256 : 128);
    if (c->threads % 32 != 0 || c->threads > 256) {
        delete c;
        return fail(PMC_ERR_INVALID, "threads must be a multiple of 32, at most 256");
    }
    if (cfg->mode == PMC_MODE_CHAINS && c->Npad > 65535) {
        delete c;
        return fail(PMC_ERR_UNSUPPORTED, "PMC_MODE_CHAINS supports at most 65535 particles per chain");
    }
    const size_t M = cfg->n_chains, Np = c->Npad, d = cfg->dim;
    cudaError_t a = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    c->stream = c->own_stream;
    if (a == cudaSuccess) a = cudaEventCreate(&c->ev0);
    if (a == cudaSuccess) a = cudaEventCreate(&c->ev1);
    if (cfg->mode == PMC_MODE_CHAINS) {
        if (a == cudaSuccess) a = dalloc(&c->x, M * d * Np);
        if (a == cudaSuccess) a = dalloc(&c->img, M * d * Np);
        if (a == cudaSuccess) a = dalloc(&c->sp, M * Np);
        if (a == cudaSuccess) a = dalloc(&c->spids, M * Np);
        if (a == cudaSuccess) a = dalloc(&c->heads, M * Np);
        if (a == cudaSuccess) a = dalloc(&c->spoff, M * (PMC_MAX_SPECIES + 1));
        if (a == cudaSuccess) a = dalloc(&c->eloc, M * Np);
        if (a == cudaSuccess) a = dalloc(&c->bonds, Np * PMC_MAX_BONDS);
    }
    if (a == cudaSuccess) a = dalloc(&c->box, M * 3);
    if (a == cudaSuccess) a = dalloc(&c->temp, M);
    if (a == cudaSuccess) a = dalloc(&c->energy, M);
    if (a == cudaSuccess) a = dalloc(&c->etot, M);
    if (a == cudaSuccess) a = dalloc(&c->calls, M * PMC_MAX_MOVES);
    if (a == cudaSuccess) a = dalloc(&c->accepted, M * PMC_MAX_MOVES);
    if (a == cudaSuccess) a = dalloc(&c->par, (size_t)PMC_MAX_SPECIES * PMC_MAX_SPECIES * PMC_NPAR);
    if (a == cudaSuccess) a = dalloc(&c->bad, 1);
    if (a == cudaSuccess) a = dalloc(&c->stats, 4);
    if (a == cudaSuccess && cfg->mode == PMC_MODE_CHAINS) a = dalloc(&c->queue, M + 1);
    // the zero-fills above ran on the legacy default stream, which the context's non-blocking stream does not
    // wait for: drain them before any kernel of this context can touch the buffers
    if (a == cudaSuccess) a = cudaDeviceSynchronize();
    if (a != cudaSuccess) {
        pmc_destroy(c);
        return fail(PMC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(a));
    }
    if (cfg->mode == PMC_MODE_BOX) {
        int rc = pmc::box_create(&c->boxst, *cfg);
        if (rc) {
            std::string msg = pmc::box_error();
            pmc_destroy(c);
            return fail(rc, "%s", msg.c_str());
        }
    }
    *out = c;
    return PMC_OK;
}

void pmc_destroy(pmc_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->boxst) pmc::box_destroy(c->boxst);
    void *bufs[] = {c->x, c->img, c->sp, c->spids, c->heads, c->spoff, c->box, c->temp, c->energy, c->etot, c->eloc,
                    c->par, c->calls, c->accepted, c->bonds, c->bad, c->queue, c->stats, c->raw_pos, c->raw_sp, c->mol_start, c->mol_len};
    for (void *p : bufs)
        if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int pmc_set_stream(pmc_ctx *c, void *cuda_stream) {
    if (!c) return fail(PMC_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    if (c->boxst) pmc::box_set_stream(c->boxst, c->stream);
    return PMC_OK;
}

int pmc_set_model(pmc_ctx *c, const double *params) {
    if (!c || !params) return fail(PMC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    const int ns = c->cfg.n_species;
    for (int k = 0; k < ns * ns; k++) {
        const double *p = params + (size_t)k * PMC_NPAR;
        if (!(p[PMC_P_RCUT] > 0.0) || !std::isfinite(p[PMC_P_RCUT2]))
            return fail(PMC_ERR_INVALID, "species pair %d has a non-positive or non-finite cutoff", k);
    }
    CU(cudaMemcpyAsync(c->par, params, sizeof(double) * ns * ns * PMC_NPAR, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->boxst) {
        int rc = pmc::box_set_model(c->boxst, params);
        if (rc) return fail(rc, "%s", pmc::box_error());
    }
    c->model_set = true;
    return PMC_OK;
}

int pmc_set_molecules(pmc_ctx *c, int32_t n_mol, const int32_t *start, const int32_t *length) {
    if (!c || !start || !length) return fail(PMC_ERR_INVALID, "null argument");
    if (n_mol < 1) return fail(PMC_ERR_INVALID, "n_molecules must be >= 1");
    CU(cudaSetDevice(c->cfg.device));
    for (int m = 0; m < n_mol; m++)
        if (start[m] < 0 || length[m] < 1 || start[m] + length[m] > c->cfg.n_particles)
            return fail(PMC_ERR_INVALID, "molecule %d: sites [%d, %d) outside 0..%d", m, start[m], start[m] + length[m], c->cfg.n_particles);
    if (c->mol_start) cudaFree(c->mol_start);
    if (c->mol_len) cudaFree(c->mol_len);
    c->mol_start = c->mol_len = nullptr;
    CU(cudaMalloc((void **)&c->mol_start, sizeof(int32_t) * n_mol));
    CU(cudaMalloc((void **)&c->mol_len, sizeof(int32_t) * n_mol));
    CU(cudaMemcpyAsync(c->mol_start, start, sizeof(int32_t) * n_mol, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->mol_len, length, sizeof(int32_t) * n_mol, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->n_mol = n_mol;
    c->mol_uniform_len = length[0];
    for (int m = 1; m < n_mol; m++)
        if (length[m] != length[0]) c->mol_uniform_len = 0;
    return PMC_OK;
}

int pmc_set_bonds(pmc_ctx *c, const int32_t *off, const int32_t *idx) {
    if (!c || !off || !idx) return fail(PMC_ERR_INVALID, "null argument");
    if (!c->cfg.molecules) return fail(PMC_ERR_INVALID, "context was not created with molecules = 1");
    if (c->cfg.mode != PMC_MODE_CHAINS) return fail(PMC_ERR_UNSUPPORTED, "bonds are only supported in PMC_MODE_CHAINS");
    CU(cudaSetDevice(c->cfg.device));
    const int N = c->cfg.n_particles;
    std::vector<uint16_t> b((size_t)c->Npad * PMC_MAX_BONDS, (uint16_t)0xFFFF);
    for (int i = 0; i < N; i++) {
        const int n = off[i + 1] - off[i];
        if (n < 0 || n > PMC_MAX_BONDS) return fail(PMC_ERR_UNSUPPORTED, "site %d has %d bonds (max %d)", i, n, PMC_MAX_BONDS);
        for (int k = 0; k < n; k++) {
            const int j = idx[off[i] + k];
            if (j < 0 || j >= N || j == i) return fail(PMC_ERR_INVALID, "bond %d-%d out of range", i, j);
            b[(size_t)i * PMC_MAX_BONDS + k] = (uint16_t)j;
        }
    }
    CU(cudaMemcpyAsync(c->bonds, b.data(), sizeof(uint16_t) * b.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->bonds_set = true;
    return PMC_OK;
}

int pmc_upload(pmc_ctx *c, int32_t first, int32_t count, const double *position, const int64_t *species,
               const double *box, const double *temperature) {
    if (!c || !position || !species || !box || !temperature) return fail(PMC_ERR_INVALID, "null argument");
    if (first < 0 || count < 1 || first + count > c->cfg.n_chains)
        return fail(PMC_ERR_INVALID, "chain range [%d, %d) outside 0..%d", first, first + count, c->cfg.n_chains);
    CU(cudaSetDevice(c->cfg.device));
    const int N = c->cfg.n_particles, d = c->cfg.dim;
    std::vector<double> b3((size_t)count * 3, 1.0);
    for (int k = 0; k < count; k++) {
        for (int a = 0; a < d; a++) {
            const double L = box[(size_t)k * d + a];
            if (!(L > 0.0) || !std::isfinite(L)) return fail(PMC_ERR_INVALID, "chain %d: box length must be positive", first + k);
            b3[(size_t)k * 3 + a] = L;
            if (L != box[(size_t)k * d]) c->cubic = false;
            if (L < c->min_box) c->min_box = L;
        }
        if (!(temperature[k] > 0.0)) return fail(PMC_ERR_INVALID, "chain %d: temperature must be positive", first + k);
    }
    if (c->cfg.mode == PMC_MODE_BOX) {
        int rc = pmc::box_upload(c->boxst, position, species, b3.data(), temperature[0]);
        if (rc) return fail(rc, "%s", pmc::box_error());
        c->uploaded = true;
        c->energy_set = false;
        return PMC_OK;
    }
    int rc = ensure_staging(c, (size_t)count);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->raw_pos, position, sizeof(double) * (size_t)count * N * d, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->raw_sp, species, sizeof(long long) * (size_t)count * N, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->box + (size_t)first * 3, b3.data(), sizeof(double) * 3 * count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->temp + first, temperature, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->bad, 0, sizeof(int), c->stream));
    k_ingest<<<count, 256, 0, c->stream>>>(c->raw_pos, c->raw_sp, c->box, first, count, N, c->Npad, d,
                                           c->cfg.n_species, c->x, c->img, c->sp, c->spids, c->heads, c->spoff, c->bad);
    CU(cudaGetLastError());
    c->launches++;
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, c->bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // b3 / caller buffers may go away after return
    if (bad == 1) return fail(PMC_ERR_INVALID, "positions contain NaN or Inf");
    if (bad == 2) return fail(PMC_ERR_INVALID, "species labels must lie in 1..%d", c->cfg.n_species);
    c->uploaded = true;
    c->energy_set = false;
    return PMC_OK;
}

int pmc_init_energy(pmc_ctx *c) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    CU(cudaSetDevice(c->cfg.device));
    const int M = c->cfg.n_chains;
    std::vector<double> e(M);
    if (c->cfg.mode == PMC_MODE_BOX) {
        rc = pmc::box_init_energy(c->boxst, e.data());
        if (rc) return fail(rc, "%s", pmc::box_error());
        c->launches += pmc::box_take_launches(c->boxst);
    } else {
        rc = run_energy(c);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->energy, c->etot, sizeof(double) * M, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(e.data(), c->etot, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    for (int k = 0; k < M; k++)
        if (!std::isfinite(e[k]))
            return fail(PMC_ERR_NONFINITE, "Initial configuration has infinite or NaN energy. (chain %d)", k);
    c->energy_set = true;
    return PMC_OK;
}

int pmc_set_moves(pmc_ctx *c, const pmc_move *pool, int32_t n) {
    if (!c || !pool) return fail(PMC_ERR_INVALID, "null argument");
    if (n < 1 || n > PMC_MAX_MOVES) return fail(PMC_ERR_INVALID, "n_moves must be in 1..%d", PMC_MAX_MOVES);
    double tot = 0.0;
    for (int k = 0; k < n; k++) {
        const pmc_move &m = pool[k];
        if (m.kind != PMC_MOVE_DISPLACEMENT && m.kind != PMC_MOVE_SWAP && m.kind != PMC_MOVE_FLIP)
            return fail(PMC_ERR_INVALID, "move %d: unknown kind %d", k, m.kind);
        if (m.kind == PMC_MOVE_FLIP) {
            if (c->n_mol < 1) return fail(PMC_ERR_STATE, "move %d: MoleculeFlip needs pmc_set_molecules first", k);
            if (c->cfg.mode == PMC_MODE_BOX) return fail(PMC_ERR_UNSUPPORTED, "MoleculeFlip is not available in PMC_MODE_BOX");
        }
        if (!(m.probability >= 0.0)) return fail(PMC_ERR_INVALID, "move %d: negative probability", k);
        if (m.kind == PMC_MOVE_DISPLACEMENT && !(m.sigma > 0.0)) return fail(PMC_ERR_INVALID, "move %d: sigma must be positive", k);
        if (m.kind == PMC_MOVE_SWAP) {
            if (m.species_a < 1 || m.species_a > c->cfg.n_species || m.species_b < 1 || m.species_b > c->cfg.n_species ||
                m.species_a == m.species_b)
                return fail(PMC_ERR_INVALID, "move %d: swap species (%d, %d) invalid", k, m.species_a, m.species_b);
            if (c->cfg.mode == PMC_MODE_BOX) return fail(PMC_ERR_UNSUPPORTED, "DiscreteSwap is not available in PMC_MODE_BOX");
        }
        tot += m.probability;
    }
    if (!(tot > 0.0)) return fail(PMC_ERR_INVALID, "move probabilities sum to zero");
    if (c->cfg.mode == PMC_MODE_BOX && n != 1)
        return fail(PMC_ERR_UNSUPPORTED, "PMC_MODE_BOX sweeps with ONE Displacement move (got a pool of %d)", n);
    c->pool.assign(pool, pool + n);
    if (c->boxst) pmc::box_set_sigma(c->boxst, pool[0].sigma);
    return PMC_OK;
}

int pmc_seed(pmc_ctx *c, uint64_t seed) {
    if (!c) return fail(PMC_ERR_INVALID, "null context");
    c->seed = seed;
    c->t0 = 0;
    if (c->boxst) pmc::box_seed(c->boxst, seed);
    return PMC_OK;
}

int pmc_run(pmc_ctx *c, int64_t n_trials) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    // a displaced particle is folded back with ONE box length (fp32 Box-Muller proposals end at 6.7 sigma)
    for (auto &m : c->pool)
        if (m.kind == PMC_MOVE_DISPLACEMENT && !(7.0 * m.sigma < c->min_box))
            return fail(PMC_ERR_INVALID, "Displacement sigma %g is too wide for a box of length %g (need 7 sigma < L)", m.sigma, c->min_box);
    if (n_trials < 0) return fail(PMC_ERR_INVALID, "n_trials must be >= 0");
    if (n_trials == 0) return PMC_OK;
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.mode == PMC_MODE_BOX) {
        CU(cudaEventRecord(c->ev0, c->stream));
        rc = pmc::box_run(c->boxst, n_trials);
        if (rc) return fail(rc, "%s", pmc::box_error());
        CU(cudaEventRecord(c->ev1, c->stream));
        c->have_run_events = true;
        c->launches += pmc::box_take_launches(c->boxst);
        return PMC_OK;
    }
    return sweep(c, n_trials, nullptr, nullptr, nullptr, nullptr, false);
}

int pmc_sync(pmc_ctx *c) {
    if (!c) return fail(PMC_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->boxst) {  // surfaces asynchronous device-side conditions (stencil overflow, inter-GPU barrier timeout)
        int rc = pmc::box_check(c->boxst);
        if (rc) return fail(rc, "%s", pmc::box_error());
    }
    return PMC_OK;
}

static int traced_or_replay(pmc_ctx *c, int64_t n, const pmc_trial *in, pmc_trial *out, uint8_t *acc, double *dE) {
    int rc = check_ready(c, !in);
    if (rc) return rc;
    if (in && !c->energy_set) return fail(PMC_ERR_STATE, "pmc_init_energy has not been called");
    if (c->cfg.mode != PMC_MODE_CHAINS) return fail(PMC_ERR_UNSUPPORTED, "trace/replay need PMC_MODE_CHAINS");
    if (n < 1) return fail(PMC_ERR_INVALID, "n_trials must be >= 1");
    CU(cudaSetDevice(c->cfg.device));
    const size_t tot = (size_t)c->cfg.n_chains * (size_t)n;
    pmc_trial *d_tr = nullptr;
    uint8_t *d_acc = nullptr;
    double *d_dE = nullptr;
    std::vector<pmc_trial> fixed;
    bool replay_swaps = false;
    if (in) {
        fixed.assign(in, in + tot);
        for (auto &t : fixed) {
            if (t.kind == PMC_MOVE_DISPLACEMENT) t.j = -1;
            if (t.kind == PMC_MOVE_SWAP || t.kind == PMC_MOVE_FLIP) replay_swaps = true;
            const bool ok = t.move >= 0 && t.move < PMC_MAX_MOVES && t.i >= 0 && t.i < c->cfg.n_particles &&
                            (t.kind == PMC_MOVE_DISPLACEMENT ||
                             ((t.kind == PMC_MOVE_SWAP || t.kind == PMC_MOVE_FLIP) && t.j >= 0 && t.j < c->cfg.n_particles && t.j != t.i));
            if (!ok) return fail(PMC_ERR_INVALID, "replay trial out of range (kind %d, move %d, i %d, j %d)", t.kind, t.move, t.i, t.j);
        }
    }
    cudaError_t e = cudaMalloc((void **)&d_tr, sizeof(pmc_trial) * tot);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_acc, tot);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_dE, sizeof(double) * tot);
    if (e == cudaSuccess && in) e = cudaMemcpyAsync(d_tr, fixed.data(), sizeof(pmc_trial) * tot, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        const size_t saved_moves = c->pool.size();
        if (in && c->pool.empty()) c->pool.push_back(pmc_move{PMC_MOVE_DISPLACEMENT, 0, 0, 0, 1.0, 1.0});
        rc = in ? sweep(c, n, d_tr, nullptr, d_acc, d_dE, true, replay_swaps) : sweep(c, n, nullptr, d_tr, d_acc, d_dE, false);
        if (in && saved_moves == 0) c->pool.clear();
        if (rc == PMC_OK) {
            if (out) e = cudaMemcpyAsync(out, d_tr, sizeof(pmc_trial) * tot, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess && acc) e = cudaMemcpyAsync(acc, d_acc, tot, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess && dE) e = cudaMemcpyAsync(dE, d_dE, sizeof(double) * tot, cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        }
    }
    cudaStreamSynchronize(c->stream);
    if (d_tr) cudaFree(d_tr);
    if (d_acc) cudaFree(d_acc);
    if (d_dE) cudaFree(d_dE);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(PMC_ERR_CUDA, "trace/replay: %s", cudaGetErrorString(e));
    return PMC_OK;
}

int pmc_run_traced(pmc_ctx *c, int64_t n, pmc_trial *trials, uint8_t *accepted, double *dE) {
    if (!trials) return fail(PMC_ERR_INVALID, "null argument");
    return traced_or_replay(c, n, nullptr, trials, accepted, dE);
}

int pmc_replay(pmc_ctx *c, int64_t n, const pmc_trial *trials, uint8_t *accepted, double *dE) {
    if (!trials) return fail(PMC_ERR_INVALID, "null argument");
    return traced_or_replay(c, n, trials, nullptr, accepted, dE);
}

int pmc_energy(pmc_ctx *c, double *out) {
    if (!c || !out) return fail(PMC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.mode == PMC_MODE_BOX) {
        int rc = pmc::box_energy(c->boxst, out);
        if (rc) return fail(rc, "%s", pmc::box_error());
        return PMC_OK;
    }
    CU(cudaMemcpyAsync(out, c->energy, sizeof(double) * c->cfg.n_chains, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PMC_OK;
}

int pmc_total_energy(pmc_ctx *c, double *out) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (!out) return fail(PMC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.mode == PMC_MODE_BOX) {
        rc = pmc::box_total_energy(c->boxst, out);
        if (rc) return fail(rc, "%s", pmc::box_error());
        c->launches += pmc::box_take_launches(c->boxst);
        return PMC_OK;
    }
    rc = run_energy(c);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, c->etot, sizeof(double) * c->cfg.n_chains, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PMC_OK;
}

int pmc_local_energy(pmc_ctx *c, int32_t chain, double *out) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (!out || chain < 0 || chain >= c->cfg.n_chains) return fail(PMC_ERR_INVALID, "bad chain index or null output");
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.mode == PMC_MODE_BOX) {
        rc = pmc::box_local_energy(c->boxst, out);
        if (rc) return fail(rc, "%s", pmc::box_error());
        c->launches += pmc::box_take_launches(c->boxst);
        return PMC_OK;
    }
    rc = run_energy(c);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, c->eloc + (size_t)chain * c->Npad, sizeof(double) * c->cfg.n_particles, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PMC_OK;
}

int pmc_download(pmc_ctx *c, int32_t first, int32_t count, double *position, int64_t *species) {
    if (!c || !position || !species) return fail(PMC_ERR_INVALID, "null argument");
    if (!c->uploaded) return fail(PMC_ERR_STATE, "nothing uploaded yet");
    if (first < 0 || count < 1 || first + count > c->cfg.n_chains) return fail(PMC_ERR_INVALID, "chain range outside 0..%d", c->cfg.n_chains);
    CU(cudaSetDevice(c->cfg.device));
    if (c->cfg.mode == PMC_MODE_BOX) {
        int rc = pmc::box_download(c->boxst, position, species);
        if (rc) return fail(rc, "%s", pmc::box_error());
        c->launches += pmc::box_take_launches(c->boxst);
        return PMC_OK;
    }
    const int N = c->cfg.n_particles, d = c->cfg.dim;
    int rc = ensure_staging(c, (size_t)count);
    if (rc) return rc;
    k_egress<<<count, 256, 0, c->stream>>>(c->x, c->img, c->sp, c->box, first, N, c->Npad, d, c->raw_pos, c->raw_sp);
    CU(cudaGetLastError());
    c->launches++;
    CU(cudaMemcpyAsync(position, c->raw_pos, sizeof(double) * (size_t)count * N * d, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(species, c->raw_sp, sizeof(long long) * (size_t)count * N, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return PMC_OK;
}

int pmc_pair_histogram(pmc_ctx *c, int32_t sa, int32_t sb, double rmax, int32_t nbins, uint64_t *hist) {
    if (!c || !hist) return fail(PMC_ERR_INVALID, "null argument");
    if (!c->uploaded) return fail(PMC_ERR_STATE, "nothing uploaded yet");
    if (nbins < 1 || nbins > 8192 || !(rmax > 0.0)) return fail(PMC_ERR_INVALID, "need 1 <= nbins <= 8192 and rmax > 0");
    if (sa < 0 || sa > c->cfg.n_species || sb < 0 || sb > c->cfg.n_species) return fail(PMC_ERR_INVALID, "species labels must lie in 0..%d", c->cfg.n_species);
    CU(cudaSetDevice(c->cfg.device));
    unsigned long long *d_hist = nullptr;
    CU(cudaMalloc((void **)&d_hist, sizeof(unsigned long long) * nbins));
    cudaError_t e = cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, c->stream);
    int rc = PMC_OK;
    if (e == cudaSuccess) {
        if (c->cfg.mode == PMC_MODE_BOX) {
            rc = pmc::box_pair_histogram(c->boxst, sa - 1, sb - 1, rmax, nbins, d_hist);
            c->launches += pmc::box_take_launches(c->boxst);
        } else {
            std::vector<double> hb((size_t)c->cfg.n_chains * 3);
            e = cudaMemcpyAsync(hb.data(), c->box, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            for (int k = 0; e == cudaSuccess && k < c->cfg.n_chains; k++)
                for (int a = 0; a < c->cfg.dim; a++)
                    if (rmax > 0.5 * hb[(size_t)k * 3 + a]) rc = PMC_ERR_INVALID;
            if (rc == PMC_OK && e == cudaSuccess) {
                e = pmc::launch_chain_pair_histogram(c->cfg.dim, c->cfg.n_chains, c->cfg.n_particles, c->Npad, c->x, c->sp, c->box,
                                                     sa - 1, sb - 1, rmax, nbins, d_hist, c->stream);
                c->launches++;
            }
        }
    }
    if (e == cudaSuccess && rc == PMC_OK) e = cudaMemcpyAsync(hist, d_hist, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_hist);
    if (rc == PMC_ERR_INVALID) return fail(rc, "rmax must not exceed half the box length (minimum image)");
    if (rc) return fail(rc, "%s", pmc::box_error());
    if (e != cudaSuccess) return fail(PMC_ERR_CUDA, "pair histogram: %s", cudaGetErrorString(e));
    return PMC_OK;
}

int pmc_chain_correlation(pmc_ctx *c, double *out) {
    if (!c || !out) return fail(PMC_ERR_INVALID, "null argument");
    if (c->cfg.mode != PMC_MODE_CHAINS) return fail(PMC_ERR_UNSUPPORTED, "chain_correlation is defined for PMC_MODE_CHAINS contexts");
    if (!c->uploaded) return fail(PMC_ERR_STATE, "nothing uploaded yet");
    if (c->n_mol < 1) return fail(PMC_ERR_STATE, "pmc_set_molecules first");
    if (c->mol_uniform_len == 0) return fail(PMC_ERR_INVALID, "All chains must have the same length");
    if (c->mol_uniform_len < 2) return fail(PMC_ERR_INVALID, "Chains must have at least two particles");
    if (c->mol_uniform_len > 64) return fail(PMC_ERR_UNSUPPORTED, "molecules of more than 64 sites");
    CU(cudaSetDevice(c->cfg.device));
    const int M = c->cfg.n_chains, len = c->mol_uniform_len;
    double *d_out = nullptr;
    CU(cudaMalloc((void **)&d_out, sizeof(double) * M));
    k_chain_correlation<<<M, 128, sizeof(double) * len * (len - 1) / 2, c->stream>>>(c->sp, c->mol_start, c->n_mol, len, c->Npad, d_out);
    cudaError_t e = cudaGetLastError();
    c->launches++;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(double) * M, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(PMC_ERR_CUDA, "chain correlation: %s", cudaGetErrorString(e));
    return PMC_OK;
}

int pmc_energy_histogram(pmc_ctx *c, double emin, double emax, int32_t nbins, int32_t per_particle, uint64_t *hist) {
    if (!c || !hist) return fail(PMC_ERR_INVALID, "null argument");
    if (!c->energy_set) return fail(PMC_ERR_STATE, "pmc_init_energy has not been called");
    if (nbins < 1 || nbins > (1 << 20) || !(emax > emin)) return fail(PMC_ERR_INVALID, "need 1 <= nbins <= 2^20 and emax > emin");
    CU(cudaSetDevice(c->cfg.device));
    const int M = c->cfg.mode == PMC_MODE_BOX ? 1 : c->cfg.n_chains;
    std::vector<double> e_host;
    const double *d_energy = c->energy;
    double *d_tmp = nullptr;
    if (c->cfg.mode == PMC_MODE_BOX) {  // the box keeps its running energy on the host side of the API
        e_host.resize(1);
        int rc = pmc_energy(c, e_host.data());
        if (rc) return rc;
        CU(cudaMalloc((void **)&d_tmp, sizeof(double)));
        CU(cudaMemcpyAsync(d_tmp, e_host.data(), sizeof(double), cudaMemcpyHostToDevice, c->stream));
        d_energy = d_tmp;
    }
    unsigned long long *d_hist = nullptr;
    CU(cudaMalloc((void **)&d_hist, sizeof(unsigned long long) * nbins));
    cudaError_t e = cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, c->stream);
    if (e == cudaSuccess) {
        k_energy_histogram<<<(M + 255) / 256, 256, 0, c->stream>>>(d_energy, M, per_particle ? 1.0 / c->cfg.n_particles : 1.0, emin,
                                                                  nbins / (emax - emin), nbins, d_hist);
        e = cudaGetLastError();
        c->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hist, d_hist, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_hist);
    if (d_tmp) cudaFree(d_tmp);
    if (e != cudaSuccess) return fail(PMC_ERR_CUDA, "energy histogram: %s", cudaGetErrorString(e));
    return PMC_OK;
}

int pmc_counters(pmc_ctx *c, int64_t *calls, int64_t *accepted) {
    if (!c || !calls || !accepted) return fail(PMC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->cfg.device));
    const int M = c->cfg.n_chains, nm = (int)c->pool.size();
    if (c->cfg.mode == PMC_MODE_BOX) {
        int rc = pmc::box_counters(c->boxst, calls, accepted);
        if (rc) return fail(rc, "%s", pmc::box_error());
        return PMC_OK;
    }
    std::vector<unsigned long long> hc((size_t)M * PMC_MAX_MOVES), ha((size_t)M * PMC_MAX_MOVES);
    CU(cudaMemcpyAsync(hc.data(), c->calls, sizeof(unsigned long long) * hc.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(ha.data(), c->accepted, sizeof(unsigned long long) * ha.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < M; k++)
        for (int m = 0; m < nm; m++) {
            calls[(size_t)k * nm + m] = (int64_t)hc[(size_t)k * PMC_MAX_MOVES + m];
            accepted[(size_t)k * nm + m] = (int64_t)ha[(size_t)k * PMC_MAX_MOVES + m];
        }
    return PMC_OK;
}

int pmc_box_peer_export(pmc_ctx *c, uint8_t *handle) {
    if (!c || !handle) return fail(PMC_ERR_INVALID, "null argument");
    if (c->cfg.mode != PMC_MODE_BOX) return fail(PMC_ERR_INVALID, "peer attachment exists only in PMC_MODE_BOX");
    CU(cudaSetDevice(c->cfg.device));
    int rc = pmc::box_peer_export(c->boxst, handle);
    if (rc) return fail(rc, "%s", pmc::box_error());
    return PMC_OK;
}

int pmc_box_peer_attach(pmc_ctx *c, int32_t rank, int32_t world, const uint8_t *handles) {
    if (!c || !handles) return fail(PMC_ERR_INVALID, "null argument");
    if (c->cfg.mode != PMC_MODE_BOX) return fail(PMC_ERR_INVALID, "peer attachment exists only in PMC_MODE_BOX");
    CU(cudaSetDevice(c->cfg.device));
    int rc = pmc::box_peer_attach(c->boxst, rank, world, handles);
    if (rc) return fail(rc, "%s", pmc::box_error());
    return PMC_OK;
}

int64_t pmc_launch_count(const pmc_ctx *c) { return c ? c->launches : 0; }

int pmc_last_run_ms(pmc_ctx *c, float *ms) {
    if (!c || !ms) return fail(PMC_ERR_INVALID, "null argument");
    if (!c->have_run_events) return fail(PMC_ERR_STATE, "no pmc_run has been issued yet");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PMC_OK;
}

int pmc_work_counters(pmc_ctx *c, int32_t enable, uint64_t *out) {
    if (!c) return fail(PMC_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->cfg.device));
    if (enable == 1) {
        CU(cudaMemsetAsync(c->stats, 0, 4 * sizeof(unsigned long long), c->stream));
        c->stats_on = true;
    } else if (enable == 0 || enable == 2) {
        if (!out) return fail(PMC_ERR_INVALID, "null argument");
        CU(cudaMemcpyAsync(out, c->stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (enable == 0) c->stats_on = false;
    } else {
        return fail(PMC_ERR_INVALID, "enable must be 0, 1 or 2");
    }
    if (c->boxst) pmc::box_set_stats(c->boxst, c->stats_on ? c->stats : nullptr);
    return PMC_OK;
}

int pmc_measure_fma_peak(int32_t device, int32_t fp64, double *tflops) {
    if (!tflops) return fail(PMC_ERR_INVALID, "null argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    void *out = nullptr;
    CU(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(e0));
        if (fp64)
            k_fma_burst<double><<<blocks, threads>>>((double *)out, iters);
        else
            k_fma_burst<float><<<blocks, threads>>>((float *)out, iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / ((double)best * 1e-3) / 1e12;
    return PMC_OK;
}

}  // extern "C"
