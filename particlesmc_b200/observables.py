"""On-device observables used to establish statistical parity: pair-distance histograms -> g(r).

The counting runs on the GPU (``pmc_pair_histogram``); this module only normalises.  With chains sharded over
GPUs the raw counts of the ranks are summed with one all-reduce (``sharding.allreduce_sum``) -- the only place a
collective appears outside the timing of bench.py.
"""
from __future__ import annotations

import numpy as np


def radial_distribution(ctx, n_a: int, n_b: int, volume: float, species_a: int = 0, species_b: int = 0,
                        rmax: float = 3.0, nbins: int = 60, n_configs: int | None = None, allreduce=None):
    """g_ab(r) of the configurations currently held by ``ctx``.

    n_a, n_b: particles of species a / b per configuration (n_b = n_a for a == b); volume: box volume (d-dim);
    n_configs: configurations contributing (default: all chains of ctx); allreduce: optional callable summing a
    numpy array over ranks."""
    counts = ctx.pair_histogram(species_a, species_b, rmax, nbins).astype(np.float64)
    n_configs = ctx.n_chains if n_configs is None else n_configs
    if allreduce is not None:
        counts = allreduce(counts)
    edges = np.linspace(0.0, rmax, nbins + 1)
    d = ctx.dim
    shell = (4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)) if d == 3 else (np.pi * (edges[1:] ** 2 - edges[:-1] ** 2))
    same = species_a == species_b
    n_pairs = n_a * (n_a - 1) / 2.0 if same else n_a * n_b  # unordered pairs per configuration
    ideal = n_pairs * shell / volume
    r = 0.5 * (edges[1:] + edges[:-1])
    return r, counts / (n_configs * ideal)


def chain_correlation(ctx):
    """The ``chain_correlation`` callback (src/molecules.jl:244-246) for every chain held by ``ctx``; computed on
    the device from the species field (``pmc_chain_correlation``)."""
    return ctx.chain_correlation()


class EnergyHistogram:
    """Accumulates ``pmc_energy_histogram`` calls over a run (the observable behind the energy-distribution checks
    of test/gerhard_energy_distribution.jl); ``add`` at every schedule point, ``density`` at the end."""

    def __init__(self, emin: float, emax: float, nbins: int, per_particle: bool = True):
        self.emin, self.emax, self.nbins, self.per_particle = float(emin), float(emax), int(nbins), per_particle
        self.counts = np.zeros(nbins, dtype=np.uint64)

    def add(self, ctx, allreduce=None):
        h = ctx.energy_histogram(self.emin, self.emax, self.nbins, self.per_particle)
        if allreduce is not None:  # sums come back as float64 (exact for counts < 2^53)
            h = np.rint(np.asarray(allreduce(h))).astype(np.uint64)
        self.counts += h

    @property
    def edges(self):
        return np.linspace(self.emin, self.emax, self.nbins + 1)

    def density(self):
        w = (self.emax - self.emin) / self.nbins
        tot = self.counts.sum()
        return self.counts / (tot * w) if tot else self.counts.astype(np.float64)
