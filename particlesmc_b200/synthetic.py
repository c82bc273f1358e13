"""Seeded synthetic configurations of the BASELINE shapes (the reference ships no KA configurations).

``ka_lattice`` follows what examples/lj-mixture/run-validation.py:50-64 does for its inputs: a simple-cubic
lattice filling the box, species assigned by a seeded shuffle.  Used by bench.py and the tests.
"""
from __future__ import annotations

import numpy as np


def lattice(N: int, d: int, density: float, seed: int = 0, fractions=(0.8, 0.2)):
    """N particles on the first N sites (seeded shuffle) of the smallest simple-cubic lattice with >= N sites.
    Returns (position [N,d] in [0,L), species [N] int64 1-based, box [d])."""
    rng = np.random.default_rng(seed)
    L = (N / density) ** (1.0 / d)
    m = int(np.ceil(N ** (1.0 / d) - 1e-9))
    while m ** d < N:
        m += 1
    grid = np.stack(np.meshgrid(*[np.arange(m)] * d, indexing="ij"), axis=-1).reshape(-1, d)
    sites = rng.permutation(len(grid))[:N] if m ** d > N else np.arange(N)
    pos = (grid[sites] + 0.5) * (L / m)
    counts = [int(round(f * N)) for f in fractions]
    counts[0] = N - sum(counts[1:])
    species = np.concatenate([np.full(c, k + 1, dtype=np.int64) for k, c in enumerate(counts)])
    rng.shuffle(species)
    return pos.astype(np.float64), species, np.full(d, L)


def ka_lattice(N: int = 1000, density: float = 1.2, seed: int = 0):
    """3-D Kob-Andersen 80:20 mixture on a lattice (BASELINE configs 1-3)."""
    return lattice(N, 3, density, seed, (0.8, 0.2))
