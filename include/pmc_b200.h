/*
 * pmc_b200.h -- C ABI of the B200-native Metropolis hot path of ParticlesMC.
 *
 * This is the drop-in boundary: the entry points a Julia `ccall` (or any FFI) binds in place of the
 * reference's per-move generic functions.  The reference plugs into Arianna.jl per MOVE
 * (Arianna.perform_action! src/moves.jl:11, revert_action! :76/:201, invert_action! :88/:212,
 * sample_action! :120/:238, log_proposal_density :110/:231, delta_log_target_density src/utils.jl:8);
 * a GPU needs a coarser seam, so this library replaces the `Metropolis` algorithm entry of the
 * algorithm list (src/ParticlesMC.jl:246: pool, seed, parallel, sweepstep) and advances every chain by
 * a block of trials per call.  Each function below names the reference code it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no CUDA/torch types; all pointers are HOST pointers owned by the caller,
 *     copied during the call and never retained;
 *   - positions are AoS float64 [N][dim] exactly like Vector{SVector{dim,Float64}} (src/atoms.jl:19);
 *   - species are int64 labels 1..n_species exactly like Vector{Int} (src/atoms.jl:20);
 *   - particle indices crossing this ABI are 0-BASED (the Julia shim subtracts 1);
 *   - every function returns PMC_OK (0) or an error code; pmc_last_error() gives the message
 *     (thread-local).  The library never exits or aborts.
 *   - a context is used from one host thread at a time.
 */
#ifndef PMC_B200_H
#define PMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMC_ABI_VERSION 1
#define PMC_NPAR 12        /* float64 slots per species pair */
#define PMC_MAX_SPECIES 4
#define PMC_MAX_MOVES 8
#define PMC_MAX_BONDS 6    /* bonded partners per site */

enum pmc_status {
    PMC_OK = 0,
    PMC_ERR_INVALID = 1,     /* bad argument / inconsistent configuration */
    PMC_ERR_CUDA = 2,        /* CUDA runtime error (message in pmc_last_error) */
    PMC_ERR_NONFINITE = 3,   /* "Initial configuration has infinite or NaN energy." (src/atoms.jl:53-55) */
    PMC_ERR_UNSUPPORTED = 4, /* shape does not fit this build (e.g. N too large for the chain kernels) */
    PMC_ERR_STATE = 5        /* call order violated (e.g. run before upload) */
};

/* src/models.jl: LennardJones :99, SoftSpheres :52, SmoothLennardJones :137, GeneralKG :183 */
enum pmc_model { PMC_MODEL_LJ = 1, PMC_MODEL_SOFT = 2, PMC_MODEL_SMOOTHLJ = 3, PMC_MODEL_KG = 4 };

/* Parameter slots of one species pair, flattened from the model struct fields:
 *   all models : [0] rcut  [1] rcut2  [2] eps4 (eps for SoftSpheres)  [3] sigma2  [4] shift
 *   SoftSpheres: [5] ndiv2                                      (src/models.jl:52-70)
 *   SmoothLJ   : [4] 0  [5] C0  [6] C2/sigma2  [7] C4/sigma4   (src/models.jl:137-158)
 *   GeneralKG  : [5] eps4bond [6] sigma2bond [7] rcut2bond [8] shiftbond [9] kr02 [10] r02 (:183-217) */
enum pmc_param {
    PMC_P_RCUT = 0, PMC_P_RCUT2 = 1, PMC_P_EPS = 2, PMC_P_SIG2 = 3, PMC_P_SHIFT = 4,
    PMC_P_NDIV2 = 5,
    PMC_P_C0 = 5, PMC_P_C2S2 = 6, PMC_P_C4S4 = 7,
    PMC_P_EPS4B = 5, PMC_P_SIG2B = 6, PMC_P_RCUT2B = 7, PMC_P_SHIFTB = 8, PMC_P_KR02 = 9, PMC_P_R02 = 10
};

enum pmc_mode {
    PMC_MODE_CHAINS = 0, /* many independent chains, one chain per CTA, state resident in shared memory */
    PMC_MODE_BOX = 1     /* one large box, cell lists in HBM, checkerboard sweeps */
};

enum pmc_precision {
    PMC_FP64 = 0,  /* every pair term in float64 (parity tolerance 1e-12 relative) */
    PMC_MIXED = 1  /* float32 pair terms, float64 accumulation (parity tolerance 1e-6 relative) */
};

/* src/moves.jl: Displacement :34 (+SimpleGaussian :105), DiscreteSwap :137 (+DoubleUniform :226),
 * MoleculeFlip :291 (+DoubleUniform :336-352): species exchange between two unlike sites of one molecule.
 * The per-species id lists behind DiscreteSwap (SpeciesList, src/utils.jl:31-49) follow every accepted swap as
 * update_species_list! does (:175-179).  Molecules carry no such list in the reference; here DiscreteSwap on Molecules
 * is an extension: a pool with MoleculeFlip and DiscreteSwap exchanges the two list entries of an accepted flip in
 * place, a pool without DiscreteSwap leaves the lists alone and rebuilds them (ids ascending per species, as at
 * upload) when pmc_run returns. */
enum pmc_move_kind { PMC_MOVE_DISPLACEMENT = 0, PMC_MOVE_SWAP = 1, PMC_MOVE_FLIP = 2 };

typedef struct pmc_ctx pmc_ctx;

typedef struct pmc_config {
    int32_t device;       /* CUDA device ordinal */
    int32_t mode;         /* enum pmc_mode */
    int32_t precision;    /* enum pmc_precision */
    int32_t n_chains;     /* length(chains) (src/IO/IO.jl:320-327); 1 in PMC_MODE_BOX */
    int32_t n_particles;  /* system.N, identical for all chains (src/IO/IO.jl:236-238) */
    int32_t dim;          /* system.d: 2 or 3 */
    int32_t n_species;    /* size(model_matrix, 1) */
    int32_t model_kind;   /* enum pmc_model; one kind per model matrix */
    int32_t molecules;    /* 0 = Atoms (src/atoms.jl:18), 1 = Molecules (src/molecules.jl:24) */
    int32_t chain_offset; /* global index of local chain 0: the RNG stream of a chain is keyed by its GLOBAL
                             index, so results do not depend on how chains are sharded over GPUs */
    int32_t threads;      /* CTA size of the sweep kernels; 0 = library default */
    int32_t prefilter;    /* chain kernels: 0 = integer fixed-point distance prefilter when all boxes are cubic
                             (same fp64 pair terms, far fewer fp64 distance evaluations), with four consecutive
                             trials of a chain evaluated speculatively per round where supported (same sequential
                             chain, see csrc/chains_spec.cuh); 1 = prefilter, one trial at a time; -1 = always
                             visit every candidate in fp64 (the reference-equivalent amount of work) */
    int32_t reserved[4];
} pmc_config;

/* One entry of the move pool == Arianna `Move(action, policy, parameters, probability)` as built at
 * src/ParticlesMC.jl:192-245 */
typedef struct pmc_move {
    int32_t kind;       /* enum pmc_move_kind */
    int32_t species_a;  /* DiscreteSwap.species[1] (label, 1-based) */
    int32_t species_b;  /* DiscreteSwap.species[2] */
    int32_t reserved;
    double probability; /* Move probability (need not be normalised; cumulative selection) */
    double sigma;       /* SimpleGaussian sigma */
} pmc_move;

/* One recorded / injected trial: what sample_action! drew plus the acceptance uniform. */
typedef struct pmc_trial {
    int32_t kind;    /* enum pmc_move_kind */
    int32_t move;    /* index into the pool (for the counters) */
    int32_t i;       /* Displacement.i / DiscreteSwap.i / MoleculeFlip.i, 0-based */
    int32_t j;       /* DiscreteSwap.j / MoleculeFlip.j, 0-based; -1 for displacements */
    double delta[3]; /* Displacement.delta (unused components 0) */
    double u;        /* the uniform compared with the acceptance probability */
} pmc_trial;

/* ---- life cycle ------------------------------------------------------------------------------- */
int pmc_abi_version(void);
const char *pmc_last_error(void);
/* Allocates device state for cfg (replaces building chains::Vector{System}, src/IO/IO.jl:320-327). */
int pmc_create(const pmc_config *cfg, pmc_ctx **out);
void pmc_destroy(pmc_ctx *ctx);
/* Launch all work of this context on the caller's CUDA stream (a cudaStream_t passed as void*;
 * NULL = the context's own stream).  Lets a host framework time the kernels with its own events. */
int pmc_set_stream(pmc_ctx *ctx, void *cuda_stream);

/* ---- system definition ------------------------------------------------------------------------ */
/* model_matrix (src/atoms.jl:25) flattened to [n_species][n_species][PMC_NPAR]. */
int pmc_set_model(pmc_ctx *ctx, const double *params);
/* Molecules.start_mol / length_mol (src/molecules.jl:28-29; 0-based starts), needed by MoleculeFlip. */
int pmc_set_molecules(pmc_ctx *ctx, int32_t n_molecules, const int32_t *start, const int32_t *length);
/* Molecules.bonds (src/molecules.jl:40) as CSR over sites; the topology is shared by all chains. */
int pmc_set_bonds(pmc_ctx *ctx, const int32_t *bond_offsets /*[N+1]*/, const int32_t *bond_index);
/* State of chains [first, first+count): position [count][N][dim], species [count][N], box [count][dim]
 * (system.box, src/atoms.jl:45), temperature [count].  Positions may lie outside the box (the
 * reference never re-wraps, src/moves.jl:46-48); pmc_download returns them unwrapped again. */
int pmc_upload(pmc_ctx *ctx, int32_t first, int32_t count, const double *position, const int64_t *species,
               const double *box, const double *temperature);
/* The tail of System(...) (src/atoms.jl:50-56, src/molecules.jl:88-94): (re)build the device neighbour
 * structure, set energy[1] = sum_i e_i / 2 for every chain; PMC_ERR_NONFINITE if any is Inf/NaN. */
int pmc_init_energy(pmc_ctx *ctx);

/* ---- Metropolis ------------------------------------------------------------------------------- */
/* The move pool of the Metropolis algorithm entry (src/ParticlesMC.jl:192-246). */
int pmc_set_moves(pmc_ctx *ctx, const pmc_move *pool, int32_t n_moves);
/* `seed` of the algorithm entry; chain k draws from Philox4x32-10 keyed by (seed, chain_offset + k). */
int pmc_seed(pmc_ctx *ctx, uint64_t seed);
/* Advance every chain by n_trials attempted moves (Arianna mc_sweep!(system, pool, rng; mc_steps),
 * benchmark/particles_benchmarks.jl:29).  Asynchronous; pair with pmc_sync. */
int pmc_run(pmc_ctx *ctx, int64_t n_trials);
int pmc_sync(pmc_ctx *ctx);
/* Same as pmc_run + pmc_sync, but records every trial of every chain: trials/accepted/delta_e are
 * [n_chains][n_trials].  Test hook: the oracle replays exactly these proposals. */
int pmc_run_traced(pmc_ctx *ctx, int64_t n_trials, pmc_trial *trials, uint8_t *accepted, double *delta_e);
/* Replays injected trials ([n_chains][n_trials]) through the SAME sweep kernel with the reference's
 * acceptance arithmetic min(1, exp(-(e2-e1)/T)) > u.  Outputs as in pmc_run_traced. */
int pmc_replay(pmc_ctx *ctx, int64_t n_trials, const pmc_trial *trials, uint8_t *accepted, double *delta_e);

/* ---- observables / state ---------------------------------------------------------------------- */
/* system.energy[1] of every chain (running value, src/moves.jl:11-20). */
int pmc_energy(pmc_ctx *ctx, double *energy /*[n_chains]*/);
/* Recomputed sum_i e_i / 2 of every chain (does not touch the running value). */
int pmc_total_energy(pmc_ctx *ctx, double *energy /*[n_chains]*/);
/* compute_energy_particle(system, i) for all i of one chain (src/atoms.jl:81-88, molecules.jl:206-215). */
int pmc_local_energy(pmc_ctx *ctx, int32_t chain, double *e /*[N]*/);
int pmc_download(pmc_ctx *ctx, int32_t first, int32_t count, double *position, int64_t *species);
/* Pair-distance histogram of the current configurations (the raw counts behind g(r)): unordered
 * pairs (i < j) with species (a, b) in either order (label 0 = any species), minimum-image distance < rmax,
 * nbins equal bins on [0, rmax), summed over all chains of this context.  rmax must not exceed half the box
 * (PMC_MODE_BOX: the cell side).  Multi-GPU: sum the histograms of the ranks (an all-reduce of nbins integers). */
int pmc_pair_histogram(pmc_ctx *ctx, int32_t species_a, int32_t species_b, double rmax, int32_t nbins,
                       uint64_t *hist /*[nbins]*/);
/* The `chain_correlation` callback (compute_chain_correlation, src/molecules.jl:224-246) of every chain, from the
 * species field on the device: molecules of equal length (pmc_set_molecules), species 2 counted as -1, the squared
 * cross terms of all site pairs summed.  Same error texts as the reference's assertions. */
int pmc_chain_correlation(pmc_ctx *ctx, double *out /*[n_chains]*/);
/* Histogram of the running energies energy[1] of all chains (per particle if per_particle != 0, i.e. the `energy`
 * callback, src/utils.jl:51-53): nbins equal bins on [emin, emax), values outside are dropped.  The caller
 * accumulates successive calls (and ranks) by adding the integer arrays. */
int pmc_energy_histogram(pmc_ctx *ctx, double emin, double emax, int32_t nbins, int32_t per_particle,
                         uint64_t *hist /*[nbins]*/);
/* Move.total_calls / accepted_calls per chain and pool entry: [n_chains][n_moves]. */
int pmc_counters(pmc_ctx *ctx, int64_t *calls, int64_t *accepted);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t pmc_launch_count(const pmc_ctx *ctx);
/* Device time of the most recent pmc_run, measured with CUDA events on the launch stream (ms). */
int pmc_last_run_ms(pmc_ctx *ctx, float *ms);
/* Work actually done by the sweep kernels, counted on the device (bench.py's roofline.frac_actual): enable = 1 resets
 * and starts counting in the following pmc_run calls, 0 stops and reads, 2 reads.  out[0] = candidate particles that
 * passed the integer prefilter and were evaluated in fp64 (each against the old and the new position), out[1] = trial
 * evaluations including those a speculative round had to repeat, out[2..3] = 0.  Counting costs a few instructions per
 * trial, so it is off by default; implemented by the default kernels of both modes (speculative chains, box sweep). */
int pmc_work_counters(pmc_ctx *ctx, int32_t enable, uint64_t *out /*[4]*/);

/* ---- multi-GPU, single large box (PMC_MODE_BOX) ---------------------------------------------------------- */
/* The cell grid is cut into slabs along x, one per rank (one process per GPU).  Every rank uploads the SAME state
 * and uses the same seed; it keeps the particle-order positions of the whole box, but cell lists only for its slab
 * plus the halo planes its stencils reach, and sweeps only its slab's cells.  Accepted moves are stored from inside
 * the sweep kernel straight into the peers' memory over NVLink (positions: every peer; cell-list entries and the
 * per-cell completion stamps: the peers whose halo holds the cell); a cell waits only for the stamps of its own
 * neighbour cells, and sweeps are chained by per-rank flags -- no host round trip, no barrier kernel and no
 * collective library on the data path.  The result is bit-identical to the single-GPU run.  Call order: pmc_upload
 * on every rank, exchange the 64-byte handles (any host channel, e.g. an all-gather), pmc_box_peer_attach (before
 * the first pmc_run), then pmc_init_energy / pmc_run as usual; every rank must issue the same pmc_run calls. */
#define PMC_IPC_HANDLE_BYTES 64
#define PMC_MAX_RANKS 8
int pmc_box_peer_export(pmc_ctx *ctx, uint8_t *handle /*[PMC_IPC_HANDLE_BYTES]*/);
int pmc_box_peer_attach(pmc_ctx *ctx, int32_t rank, int32_t world, const uint8_t *handles /*[world][64]*/);

/* ---- device micro-benchmarks (roofline denominators) ----------------------------- */
/* Burst FMA throughput of the CUDA-core pipes in TFLOP/s (fp64 = 1: DFMA, 0: FFMA). */
int pmc_measure_fma_peak(int32_t device, int32_t fp64, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* PMC_B200_H */
