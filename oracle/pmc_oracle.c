/*
 * oracle/pmc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, fp64, single-threaded-per-chain restatement of the CPU algorithm of
 * TheDisorderedOrganization/ParticlesMC's Metropolis hot path.  It exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs can CHECK (and time beside) the CUDA path.  Nothing under particlesmc_b200/
 * may import, link or call it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).  The arithmetic keeps the reference's operation order
 * (Julia neither reassociates nor contracts): compile with -ffp-contract=off.
 *
 * PARITY PINNING
 *   pinned  : energy arithmetic (geometry, potentials, neighbour lists, local and
 *             total energies) against the two known answers the reference's own
 *             tests hold: test/runtests.jl:36-38 (config_0 + JBB = -2.676832/N,
 *             atol 1e-6, EmptyList == LinkedList) and test/runtests.jl:148-149
 *             (molecule + Trimer = 25.65865662277199/N); see tests/test_oracle.py.
 *   UNPINNED: the accept/reject rule, move selection and RNG stream live in the
 *             un-vendored third-party package Arianna.jl (Project.toml:6, compat
 *             0.2 at Project.toml:21, no pinned patch version).  orc_step_* restate
 *             the published Metropolis-Hastings rule `min(1, exp(dlogp + dlogq)) >
 *             rand(rng)` as used through the reference's call sites
 *             (benchmark/particles_benchmarks.jl:28-29, src/ParticlesMC.jl:246,
 *             src/utils.jl:8-10); no reference test records a decision trace, so
 *             bit-level accept/reject parity is "parity unpinned" and only pinned
 *             statistically (examples/lj-mixture/calculated-energies.csv).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NPAR 12
enum { ORC_LJ = 1, ORC_SOFT = 2, ORC_SMOOTHLJ = 3, ORC_KG = 4 };
enum { ORC_EMPTYLIST = 0, ORC_LINKEDLIST = 1 };
enum { ORC_MOVE_DISPLACEMENT = 0, ORC_MOVE_SWAP = 1, ORC_MOVE_FLIP = 2 };

/* parameter slots of one species pair (same layout as include/pmc_b200.h) */
enum {
    P_RCUT = 0, P_RCUT2 = 1, P_EPS = 2 /* eps4, or eps for SoftSpheres */, P_SIG2 = 3, P_SHIFT = 4,
    P_NDIV2 = 5,                       /* SoftSpheres */
    P_C0 = 5, P_C2S2 = 6, P_C4S4 = 7,  /* SmoothLennardJones */
    P_EPS4B = 5, P_SIG2B = 6, P_RCUT2B = 7, P_SHIFTB = 8, P_KR02 = 9, P_R02 = 10 /* GeneralKG */
};

typedef struct orc_system {
    int N, d, ns, kind, list_type, is_molecule;
    double box[3], temperature, energy;
    double *pos;      /* [N][d] AoS, as Vector{SVector{d,Float64}} (atoms.jl:19) */
    int64_t *species; /* labels 1..ns (atoms.jl:20) */
    double *par;      /* [ns][ns][ORC_NPAR] */
    /* LinkedList (neighbours.jl:224-232); particle ids 0-based here, -1 = empty */
    double cell[3];
    int ncells[3], ncell_tot;
    int *cs, *head, *list;
    int *nb_count, *nb_cells; /* stencil table [ncell_tot][27] (neighbours.jl:94-111) */
    /* SpeciesList (utils.jl:31-49) */
    int *sp_off, *sp_ids, *sp_heads;
    /* bonds, CSR (molecules.jl:40) */
    int *bond_off, *bond_idx;
} orc_system;

/* ---------------------------------------------------------------- geometry */

/* utils.jl:12  fold_back(x, box) = x - fld(x, box) * box */
static double fold_back(double x, double box) { return x - floor(x / box) * box; }

/* integer fold used by cell_index (neighbours.jl:80) */
static int fold_int(int m, int n) {
    int r = m % n;
    return r < 0 ? r + n : r;
}

/* utils.jl:15-18  dx - round(dx / L) * L ; Julia round = ties-to-even = nearbyint */
static double vector_1d(double c1, double c2, double L) {
    double dx = c1 - c2;
    return dx - nearbyint(dx / L) * L;
}

/* utils.jl:24-28  sum(abs2, dx), left fold */
static double nearest_image_r2(const double *xi, const double *xj, const double *box, int d) {
    double r2 = 0.0;
    for (int a = 0; a < d; a++) {
        double dx = vector_1d(xi[a], xj[a], box[a]);
        r2 = (a == 0) ? dx * dx : r2 + dx * dx;
    }
    return r2;
}

/* ---------------------------------------------------------------- potentials */

/* models.jl:30-34 */
static double lennard_jones(double r2, double eps4, double sig2) {
    double x = sig2 * (1.0 / r2);
    double x3 = x * x * x;
    return eps4 * (x3 * x3 - x3);
}

/* models.jl:28 with Julia's integer power when ndiv2 is integral (models.jl:67) */
static double inverse_power(double r2, double eps, double sig2, double ndiv2) {
    return eps * pow(sig2 / r2, ndiv2);
}

/* potential(r2, model): models.jl:72-74 (SoftSpheres), :121-123 (LennardJones),
 * :160-166 (SmoothLennardJones, muladd -> fma), :207-209 (GeneralKG non-bonded) */
static double pair_potential(int kind, const double *p, double r2) {
    switch (kind) {
    case ORC_LJ:
    case ORC_KG:
        return lennard_jones(r2, p[P_EPS], p[P_SIG2]) - p[P_SHIFT];
    case ORC_SOFT:
        return inverse_power(r2, p[P_EPS], p[P_SIG2], p[P_NDIV2]) - p[P_SHIFT];
    case ORC_SMOOTHLJ: {
        double lj = lennard_jones(r2, p[P_EPS], p[P_SIG2]);
        double shift = p[P_EPS] * (p[P_C0] + r2 * fma(r2, p[P_C4S4], p[P_C2S2]));
        return lj + shift;
    }
    }
    return NAN;
}

/* models.jl:36 fene, :219-226 bond_potential */
static double bond_potential(const double *p, double r2) {
    double u_fene = (r2 <= p[P_R02]) ? p[P_KR02] * log(1.0 - r2 * (1.0 / p[P_R02])) : INFINITY;
    double u_lj = 0.0;
    if (r2 <= p[P_RCUT2B]) u_lj += lennard_jones(r2, p[P_EPS4B], p[P_SIG2B]) - p[P_SHIFTB];
    return u_fene + u_lj;
}

static const double *pair_par(const orc_system *s, int i, int j) {
    /* ParticlesMC.jl:55 get_model: model_matrix[species[i], species[j]] */
    return s->par + ((size_t)(s->species[i] - 1) * s->ns + (size_t)(s->species[j] - 1)) * ORC_NPAR;
}

/* ---------------------------------------------------------------- neighbour list */

/* neighbours.jl:79-88  scalar cell index, last dimension fastest (0-based here) */
static int cell_index(const int *ncells, const int *mc, int d) {
    int c = 0, stride = 1;
    for (int a = d - 1; a >= 0; a--) {
        c += fold_int(mc[a], ncells[a]) * stride;
        stride *= ncells[a];
    }
    return c;
}

/* neighbours.jl:178-180 get_cell: fld(x, cell) per axis */
static void get_cell(const orc_system *s, const double *x, int *mc) {
    for (int a = 0; a < s->d; a++) mc[a] = (int)floor(x[a] / s->cell[a]);
}

static int get_cell_index(const orc_system *s, const double *x) {
    int mc[3];
    get_cell(s, x, mc);
    return cell_index(s->ncells, mc, s->d);
}

/* neighbours.jl:236-244 constructor + :94-111 build_neighbour_cells.
 * ncells is computed as fld(box, rcut) directly (the reference's Int(box/cell) equals it
 * whenever it does not throw InexactError). Stencil enumeration: Iterators.product over
 * (x-1:x+1) per axis, FIRST axis fastest, duplicates removed keeping first occurrence. */
static void linked_list_init(orc_system *s, double rcut) {
    int d = s->d;
    s->ncell_tot = 1;
    for (int a = 0; a < d; a++) {
        int n = (int)floor(s->box[a] / rcut);
        if (n < 1) n = 1;
        s->ncells[a] = n;
        s->cell[a] = s->box[a] / (double)n;
        s->ncell_tot *= n;
    }
    s->cs = (int *)calloc(s->N, sizeof(int));
    s->list = (int *)calloc(s->N, sizeof(int));
    s->head = (int *)malloc(sizeof(int) * s->ncell_tot);
    s->nb_count = (int *)calloc(s->ncell_tot, sizeof(int));
    s->nb_cells = (int *)malloc(sizeof(int) * 27 * (size_t)s->ncell_tot);
    int nst = (d == 3) ? 27 : (d == 2 ? 9 : 3);
    int mc[3] = {0, 0, 0};
    for (int c_lin = 0; c_lin < s->ncell_tot; c_lin++) {
        /* decode c_lin -> mc (any enumeration order of centre cells gives the same table) */
        int rem = c_lin;
        for (int a = d - 1; a >= 0; a--) {
            mc[a] = rem % s->ncells[a];
            rem /= s->ncells[a];
        }
        int *out = s->nb_cells + 27 * (size_t)c_lin, cnt = 0;
        for (int k = 0; k < nst; k++) {
            int mc2[3], kk = k;
            for (int a = 0; a < d; a++) { /* first axis fastest */
                mc2[a] = mc[a] + (kk % 3) - 1;
                kk /= 3;
            }
            int c2 = cell_index(s->ncells, mc2, d), dup = 0;
            for (int q = 0; q < cnt; q++) dup |= (out[q] == c2);
            if (!dup) out[cnt++] = c2;
        }
        s->nb_count[c_lin] = cnt;
    }
}

/* neighbours.jl:251-270 build: push every particle at the head of its cell */
static void linked_list_build(orc_system *s) {
    for (int c = 0; c < s->ncell_tot; c++) s->head[c] = -1;
    for (int i = 0; i < s->N; i++) {
        int c = get_cell_index(s, s->pos + (size_t)i * s->d);
        s->list[i] = s->head[c];
        s->head[c] = i;
        s->cs[i] = c;
    }
}

/* neighbours.jl:276-302 relink particle i if its cell changed */
static void linked_list_update(orc_system *s, int i) {
    int c = s->cs[i];
    int c2 = get_cell_index(s, s->pos + (size_t)i * s->d);
    if (c == c2) return;
    if (s->head[c] == i) {
        s->head[c] = s->list[i];
    } else {
        int j = s->head[c];
        while (s->list[j] != i) j = s->list[j];
        s->list[j] = s->list[i];
    }
    s->list[i] = s->head[c2];
    s->head[c2] = i;
    s->cs[i] = c2;
}

static void update_neighbour_list(orc_system *s, int i) {
    if (s->list_type == ORC_LINKEDLIST) linked_list_update(s, i);
}

/* ---------------------------------------------------------------- local energy */

static int is_bonded(const orc_system *s, int i, int j) {
    for (int b = s->bond_off[i]; b < s->bond_off[i + 1]; b++)
        if (s->bond_idx[b] == j) return 1;
    return 0;
}

/* atoms.jl:66-75 compute_energy_ij ; molecules.jl:163-171,194-198 non-bonded variant */
static double energy_ij(const orc_system *s, int i, int j, const double *xi) {
    if (i == j) return 0.0;
    if (s->is_molecule && is_bonded(s, i, j)) return 0.0;
    const double *p = pair_par(s, i, j);
    double r2 = nearest_image_r2(xi, s->pos + (size_t)j * s->d, s->box, s->d);
    if (r2 > p[P_RCUT2]) return 0.0;
    return pair_potential(s->kind, p, r2);
}

/* atoms.jl:81-88 / molecules.jl:206-215 compute_energy_particle.  The candidate order is the
 * list's iteration order: EmptyList 1..N (neighbours.jl:42-44); LinkedList stencil cells of the
 * cell recomputed from the CURRENT position (neighbours.jl:371-377), each cell head->list
 * (neighbours.jl:332-365). */
static double local_energy(const orc_system *s, int i) {
    const double *xi = s->pos + (size_t)i * s->d;
    double e = 0.0;
    if (s->is_molecule) { /* molecules.jl:173-179 bonded part first */
        for (int b = s->bond_off[i]; b < s->bond_off[i + 1]; b++) {
            int j = s->bond_idx[b];
            double r2 = nearest_image_r2(xi, s->pos + (size_t)j * s->d, s->box, s->d);
            e += bond_potential(pair_par(s, i, j), r2);
        }
    }
    if (s->list_type == ORC_EMPTYLIST) {
        for (int j = 0; j < s->N; j++) e += energy_ij(s, i, j, xi);
    } else {
        int c = get_cell_index(s, xi);
        const int *nb = s->nb_cells + 27 * (size_t)c;
        for (int q = 0; q < s->nb_count[c]; q++)
            for (int j = s->head[nb[q]]; j != -1; j = s->list[j]) e += energy_ij(s, i, j, xi);
    }
    return e;
}

/* atoms.jl:51-52 / molecules.jl:89-90: energy = sum(local)/2 */
static double total_energy(const orc_system *s) {
    double sum = 0.0;
    for (int i = 0; i < s->N; i++) sum += local_energy(s, i);
    return sum / 2;
}

/* ---------------------------------------------------------------- species list */

/* utils.jl:36-49 SpeciesList(species): ids per species ascending, heads = slot in own list */
static void species_list_build(orc_system *s) {
    s->sp_off = (int *)calloc(s->ns + 1, sizeof(int));
    s->sp_ids = (int *)malloc(sizeof(int) * s->N);
    s->sp_heads = (int *)malloc(sizeof(int) * s->N);
    for (int i = 0; i < s->N; i++) s->sp_off[s->species[i]]++;
    for (int k = 0; k < s->ns; k++) s->sp_off[k + 1] += s->sp_off[k];
    int *cur = (int *)calloc(s->ns, sizeof(int));
    for (int i = 0; i < s->N; i++) {
        int k = (int)s->species[i] - 1;
        s->sp_heads[i] = cur[k];
        s->sp_ids[s->sp_off[k] + cur[k]++] = i;
    }
    free(cur);
}

static void species_list_build_free(orc_system *s) {
    free(s->sp_off); free(s->sp_ids); free(s->sp_heads);
    species_list_build(s);
}

/* moves.jl:175-179 update_species_list!(species_list, (A,B), i, j) */
static void species_list_update(orc_system *s, int A, int B, int i, int j) {
    s->sp_ids[s->sp_off[A - 1] + s->sp_heads[i]] = j;
    s->sp_ids[s->sp_off[B - 1] + s->sp_heads[j]] = i;
    int t = s->sp_heads[i];
    s->sp_heads[i] = s->sp_heads[j];
    s->sp_heads[j] = t;
}

/* ---------------------------------------------------------------- public: system */

void orc_destroy(orc_system *s) {
    if (!s) return;
    free(s->pos); free(s->species); free(s->par); free(s->cs); free(s->head); free(s->list);
    free(s->nb_count); free(s->nb_cells); free(s->sp_off); free(s->sp_ids); free(s->sp_heads);
    free(s->bond_off); free(s->bond_idx); free(s);
}

/* atoms.jl:40-58 / molecules.jl:76-96 System(...): box is passed in explicitly (the caller
 * evaluates (N/density)^(1/d), atoms.jl:45); bond_off == NULL selects Atoms.  Returns NULL when
 * the initial energy is Inf/NaN (atoms.jl:53-55). */
orc_system *orc_create(int N, int d, int ns, int kind, const double *par, const double *pos,
                       const int64_t *species, const double *box, double temperature, int list_type,
                       const int32_t *bond_off, const int32_t *bond_idx) {
    orc_system *s = (orc_system *)calloc(1, sizeof(orc_system));
    s->N = N; s->d = d; s->ns = ns; s->kind = kind; s->list_type = list_type;
    s->temperature = temperature;
    for (int a = 0; a < d; a++) s->box[a] = box[a];
    s->pos = (double *)malloc(sizeof(double) * (size_t)N * d);
    memcpy(s->pos, pos, sizeof(double) * (size_t)N * d);
    s->species = (int64_t *)malloc(sizeof(int64_t) * N);
    memcpy(s->species, species, sizeof(int64_t) * N);
    s->par = (double *)malloc(sizeof(double) * (size_t)ns * ns * ORC_NPAR);
    memcpy(s->par, par, sizeof(double) * (size_t)ns * ns * ORC_NPAR);
    if (bond_off) {
        s->is_molecule = 1;
        s->bond_off = (int *)malloc(sizeof(int) * (N + 1));
        memcpy(s->bond_off, bond_off, sizeof(int) * (N + 1));
        s->bond_idx = (int *)malloc(sizeof(int) * (bond_off[N] > 0 ? bond_off[N] : 1));
        memcpy(s->bond_idx, bond_idx, sizeof(int) * bond_off[N]);
    }
    double maxcut = 0.0; /* atoms.jl:46 */
    for (int k = 0; k < ns * ns; k++) maxcut = fmax(maxcut, par[(size_t)k * ORC_NPAR + P_RCUT]);
    if (list_type == ORC_LINKEDLIST) {
        linked_list_init(s, maxcut);
        linked_list_build(s);
    }
    species_list_build(s);
    s->energy = total_energy(s);
    if (isinf(s->energy) || isnan(s->energy)) {
        orc_destroy(s);
        return NULL;
    }
    return s;
}

double orc_energy(const orc_system *s) { return s->energy; }
double orc_total_energy(const orc_system *s) { return total_energy(s); }
double orc_local_energy(const orc_system *s, int i) { return local_energy(s, i); }
void orc_local_energies(const orc_system *s, double *out) {
    for (int i = 0; i < s->N; i++) out[i] = local_energy(s, i);
}
void orc_get_state(const orc_system *s, double *pos, int64_t *species) {
    memcpy(pos, s->pos, sizeof(double) * (size_t)s->N * s->d);
    memcpy(species, s->species, sizeof(int64_t) * s->N);
}
void orc_get_ncells(const orc_system *s, int *out) {
    for (int a = 0; a < s->d; a++) out[a] = s->ncells[a];
}
int orc_species_count(const orc_system *s, int A) { return s->sp_off[A] - s->sp_off[A - 1]; }
int orc_species_member(const orc_system *s, int A, int k) { return s->sp_ids[s->sp_off[A - 1] + k]; }

/* geometry / potential probes for unit tests */
double orc_nearest_image_r2(const double *xi, const double *xj, const double *box, int d) {
    return nearest_image_r2(xi, xj, box, d);
}
double orc_fold_back(double x, double box) { return fold_back(x, box); }
double orc_pair_potential(int kind, const double *p, double r2) { return pair_potential(kind, p, r2); }
double orc_bond_potential(const double *p, double r2) { return bond_potential(p, r2); }
double orc_lennard_jones(double r2, double eps4, double sig2) { return lennard_jones(r2, eps4, sig2); }

/* ---------------------------------------------------------------- public: moves */

/* Metropolis rule as reached through Arianna.mc_step! (external, parity unpinned -- see header):
 *   dlogp = -(e2 - e1) / T                    (utils.jl:8-10)
 *   alpha = min(1, exp(dlogp + logq_bwd - logq_fwd)); accept iff alpha > u
 * logq_bwd - logq_fwd is exactly 0.0 for Displacement/SimpleGaussian (moves.jl:110-112: depends on
 * |delta| only) and for DiscreteSwap/DoubleUniform (moves.jl:231-233). */
static int metropolis_accept(double e1, double e2, double T, double u) {
    double dlogp = -(e2 - e1) / T;
    double ex = exp(dlogp + 0.0);
    double alpha = (ex != ex) ? ex : fmin(1.0, ex); /* Julia's min propagates NaN (C fmin does not) */
    return alpha > u;
}

/* One Displacement trial: moves.jl:57-67 (perform), :11-20 (energy bookkeeping wrapper),
 * :88-90 (invert), :76-81 (revert).  revert_mode 0 = reference arithmetic (x+d)+(-d), (E+de)-de;
 * revert_mode 1 = restore x and E exactly (what the production GPU kernels do). */
int orc_step_displacement(orc_system *s, int i, const double *delta, double u, int revert_mode,
                          double *e1_out, double *e2_out) {
    int d = s->d;
    double *xi = s->pos + (size_t)i * d;
    double xold[3], Eold = s->energy, de;
    for (int a = 0; a < d; a++) xold[a] = xi[a];
    double e1 = local_energy(s, i);
    for (int a = 0; a < d; a++) xi[a] = xi[a] + delta[a]; /* moves.jl:46-48, no re-wrap */
    update_neighbour_list(s, i);
    double e2 = local_energy(s, i);
    if (isinf(e1) || isinf(e2)) {
        de = 0.0;
    } else {
        de = e2 - e1;
        s->energy += de;
    }
    int acc = metropolis_accept(e1, e2, s->temperature, u);
    if (!acc) {
        for (int a = 0; a < d; a++) xi[a] = xi[a] + (-delta[a]);
        s->energy -= de;
        if (revert_mode == 1) {
            for (int a = 0; a < d; a++) xi[a] = xold[a];
            s->energy = Eold;
        }
        update_neighbour_list(s, i);
    }
    if (e1_out) *e1_out = e1;
    if (e2_out) *e2_out = e2;
    return acc;
}

/* One DiscreteSwap trial between particle i (species A) and j (species B): moves.jl:159-167
 * (four local energies), :187-194 (perform), :212-214 (invert), :201-207 (revert). */
int orc_step_swap(orc_system *s, int A, int B, int i, int j, double u, int revert_mode,
                  double *e1_out, double *e2_out) {
    double Eold = s->energy, de;
    int64_t spi = s->species[i], spj = s->species[j];
    double e1i = local_energy(s, i), e1j = local_energy(s, j);
    s->species[i] = spj;
    s->species[j] = spi;
    double e2i = local_energy(s, i), e2j = local_energy(s, j);
    double e1 = e1i + e1j, e2 = e2i + e2j;
    species_list_update(s, A, B, i, j);
    if (isinf(e1) || isinf(e2)) {
        de = 0.0;
    } else {
        de = e2 - e1;
        s->energy += de;
    }
    int acc = metropolis_accept(e1, e2, s->temperature, u);
    if (!acc) { /* after invert_action! the roles are (j, i) */
        int ii = j, jj = i;
        int64_t a = s->species[ii], b = s->species[jj];
        s->species[jj] = a;
        s->species[ii] = b;
        species_list_update(s, A, B, ii, jj);
        s->energy -= de;
        if (revert_mode == 1) s->energy = Eold;
    }
    if (e1_out) *e1_out = e1;
    if (e2_out) *e2_out = e2;
    return acc;
}

/* One MoleculeFlip trial between two unlike sites i, j of one molecule: moves.jl:301-307 (perform, the same
 * four-energy swap_particle_species!), :312-314 (invert), :319-323 (revert).  Molecules carry no species list. */
int orc_step_flip(orc_system *s, int i, int j, double u, int revert_mode, double *e1_out, double *e2_out) {
    double Eold = s->energy, de;
    int64_t spi = s->species[i], spj = s->species[j];
    double e1i = local_energy(s, i), e1j = local_energy(s, j);
    s->species[i] = spj;
    s->species[j] = spi;
    double e2i = local_energy(s, i), e2j = local_energy(s, j);
    double e1 = e1i + e1j, e2 = e2i + e2j;
    if (isinf(e1) || isinf(e2)) {
        de = 0.0;
    } else {
        de = e2 - e1;
        s->energy += de;
    }
    int acc = metropolis_accept(e1, e2, s->temperature, u);
    if (!acc) {
        s->species[i] = spi;
        s->species[j] = spj;
        s->energy -= de;
        if (revert_mode == 1) s->energy = Eold;
    } else {
        /* keep the oracle's own species list usable if swaps are mixed in (not part of the reference path) */
        species_list_build_free(s);
    }
    if (e1_out) *e1_out = e1;
    if (e2_out) *e2_out = e2;
    return acc;
}

/* DoubleUniform draw (moves.jl:238-241): i = sp_ids[A][ka], j = sp_ids[B][kb] */
int orc_step_swap_draw(orc_system *s, int A, int B, int ka, int kb, double u, int revert_mode,
                       int *i_out, int *j_out, double *e1_out, double *e2_out) {
    int i = s->sp_ids[s->sp_off[A - 1] + ka], j = s->sp_ids[s->sp_off[B - 1] + kb];
    if (i_out) *i_out = i;
    if (j_out) *j_out = j;
    return orc_step_swap(s, A, B, i, j, u, revert_mode, e1_out, e2_out);
}

/* Replay a recorded proposal stream (the trace layout of pmc_trial in include/pmc_b200.h):
 * kind[t] 0 = Displacement(i, delta), 1 = DiscreteSwap(i, j) between species (spA[t], spB[t]).
 * Records the decision, e2 - e1 and the running system.energy[1] after every trial. */
void orc_replay(orc_system *s, int64_t n, const int32_t *kind, const int32_t *ii, const int32_t *jj,
                const int32_t *spA, const int32_t *spB, const double *delta /* [n][3] */,
                const double *u, int revert_mode, uint8_t *accepted, double *dE, double *E) {
    for (int64_t t = 0; t < n; t++) {
        double e1, e2;
        int acc;
        if (kind[t] == ORC_MOVE_DISPLACEMENT)
            acc = orc_step_displacement(s, ii[t], delta + 3 * t, u[t], revert_mode, &e1, &e2);
        else if (kind[t] == ORC_MOVE_FLIP)
            acc = (ii[t] >= 0 && jj[t] >= 0 && ii[t] != jj[t]) ? orc_step_flip(s, ii[t], jj[t], u[t], revert_mode, &e1, &e2)
                                                               : (e1 = e2 = 0.0, 0);
        else
            acc = orc_step_swap(s, spA[t], spB[t], ii[t], jj[t], u[t], revert_mode, &e1, &e2);
        if (accepted) accepted[t] = (uint8_t)acc;
        if (dE) dE[t] = e2 - e1;
        if (E) E[t] = s->energy;
    }
}

/* ---------------------------------------------------------------- own RNG: Philox4x32-10 */

/* Counter-based generator of Salmon et al. (SC'11), restated from the published algorithm.
 * The proposal transforms below are the SAME definitions the CUDA kernels use
 * (particlesmc_b200/csrc/rng.cuh); libm vs CUDA libdevice differ in the last ulp of logf/sincosf,
 * so bit-level parity with the GPU is established by replaying the GPU's own trace, not this. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
void orc_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof(c));
}

typedef struct {
    int32_t kind;     /* ORC_MOVE_* */
    double prob;      /* selection probability */
    double sigma;     /* Displacement */
    int32_t spA, spB; /* DiscreteSwap */
} orc_move;

/* One proposal for move number t of chain `chain` under `seed`.
 * block A = philox(t_lo, t_hi, chain, 0): A0 -> move selection, A1 -> particle index (or ka),
 *           (A2, A3) -> 53-bit acceptance uniform;
 * block B = philox(t_lo, t_hi, chain, 1): (B0, B1), (B2, B3) -> Box-Muller pairs in fp32 (or B0 -> kb). */
static void draw_trial(const orc_system *s, uint64_t seed, uint32_t chain, uint64_t t, const orc_move *pool,
                       int n_moves, int *m_out, int *i_out, int *j_out, double *delta, double *u_out) {
    uint32_t A[4] = {(uint32_t)t, (uint32_t)(t >> 32), chain, 0u};
    uint32_t B[4] = {(uint32_t)t, (uint32_t)(t >> 32), chain, 1u};
    philox4x32_10(A, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(B, (uint32_t)seed, (uint32_t)(seed >> 32));
    double um = A[0] * 0x1p-32, cum = 0.0;
    int m = n_moves - 1;
    for (int k = 0; k < n_moves; k++) {
        cum += pool[k].prob;
        if (um < cum) { m = k; break; }
    }
    *m_out = m;
    *u_out = (double)(((uint64_t)(A[2] >> 5) << 26) | (uint64_t)(A[3] >> 6)) * 0x1p-53;
    if (pool[m].kind == ORC_MOVE_DISPLACEMENT) {
        *i_out = (int)(((uint64_t)A[1] * (uint64_t)s->N) >> 32);
        *j_out = -1;
        float sg = (float)pool[m].sigma;
        float r0 = sqrtf(-2.0f * logf(((float)B[0] + 1.0f) * 0x1p-32f));
        float th0 = 6.28318530717958647692f * ((float)B[1] * 0x1p-32f);
        float r1 = sqrtf(-2.0f * logf(((float)B[2] + 1.0f) * 0x1p-32f));
        float th1 = 6.28318530717958647692f * ((float)B[3] * 0x1p-32f);
        delta[0] = (double)(sg * (r0 * cosf(th0)));
        delta[1] = (double)(sg * (r0 * sinf(th0)));
        delta[2] = (double)(sg * (r1 * cosf(th1)));
    } else {
        int A_ = pool[m].spA, B_ = pool[m].spB;
        int nA = s->sp_off[A_] - s->sp_off[A_ - 1], nB = s->sp_off[B_] - s->sp_off[B_ - 1];
        int ka = (int)(((uint64_t)A[1] * (uint64_t)nA) >> 32), kb = (int)(((uint64_t)B[0] * (uint64_t)nB) >> 32);
        *i_out = (nA > 0) ? s->sp_ids[s->sp_off[A_ - 1] + ka] : -1;
        *j_out = (nB > 0) ? s->sp_ids[s->sp_off[B_ - 1] + kb] : -1;
        delta[0] = delta[1] = delta[2] = 0.0;
    }
}

/* Run n_trials Metropolis trials of one chain with the Philox stream (mc_sweep! semantics:
 * pick a move from the pool by probability, mc_step! it, count calls/acceptances). */
void orc_run(orc_system *s, uint64_t seed, uint32_t chain, uint64_t t0, int64_t n_trials, const orc_move *pool,
             int n_moves, int revert_mode, int64_t *calls, int64_t *accepted) {
    for (int64_t q = 0; q < n_trials; q++) {
        int m, i, j, acc = 0;
        double delta[3], u;
        draw_trial(s, seed, chain, t0 + (uint64_t)q, pool, n_moves, &m, &i, &j, delta, &u);
        if (pool[m].kind == ORC_MOVE_DISPLACEMENT) {
            acc = orc_step_displacement(s, i, delta, u, revert_mode, NULL, NULL);
        } else if (i >= 0 && j >= 0) {
            acc = orc_step_swap(s, pool[m].spA, pool[m].spB, i, j, u, revert_mode, NULL, NULL);
        }
        if (calls) calls[m] += 1;
        if (accepted) accepted[m] += acc;
    }
}

/* The reference's only parallelism: independent chains over host threads (Arianna `parallel=true`,
 * src/ParticlesMC.jl:164,246).  Used by bench.py's cpu_baseline / --impl reference legs. */
int orc_run_chains(orc_system **chains, int n_chains, uint64_t seed, uint64_t t0, int64_t n_trials,
                   const orc_move *pool, int n_moves, int revert_mode, int n_threads) {
    int used = 1;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int k = 0; k < n_chains; k++)
        orc_run(chains[k], seed, (uint32_t)k, t0, n_trials, pool, n_moves, revert_mode, NULL, NULL);
    return used;
}
