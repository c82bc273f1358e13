// box.cuh -- PMC_MODE_BOX: one large periodic box, cell lists in HBM rebuilt on the device by sorting
// particles by cell (replaces src/neighbours.jl:251-270), checkerboard (non-interacting sub-cell)
// parallel sweeps.  Host-side interface used by api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pmc_b200.h"

namespace pmc {

struct BoxState;

const char *box_error();
int box_create(BoxState **out, const pmc_config &cfg);
void box_destroy(BoxState *b);
void box_set_stream(BoxState *b, cudaStream_t st);
int box_set_model(BoxState *b, const double *params);
void box_set_sigma(BoxState *b, double sigma);
void box_set_stats(BoxState *b, unsigned long long *stats);
void box_seed(BoxState *b, uint64_t seed);
int box_upload(BoxState *b, const double *pos_aos, const int64_t *species, const double *box3, double temperature);
int box_init_energy(BoxState *b, double *e_out);
int box_total_energy(BoxState *b, double *e_out);
int box_local_energy(BoxState *b, double *eloc_out);
int box_energy(BoxState *b, double *e_out);
int box_run(BoxState *b, int64_t n_trials);
int box_download(BoxState *b, double *pos_aos, int64_t *species);
int box_counters(BoxState *b, int64_t *calls, int64_t *accepted);
int64_t box_take_launches(BoxState *b);
int box_peer_export(BoxState *b, unsigned char *handle64);
int box_peer_attach(BoxState *b, int rank, int world, const unsigned char *handles);
int box_check(BoxState *b);
int box_pair_histogram(BoxState *b, int sa, int sb, double rmax, int nbins, unsigned long long *d_hist);

}  // namespace pmc
