"""System types of the hot path: host-side mirror of ``src/atoms.jl`` and ``src/molecules.jl``.

``Atoms`` / ``Molecules`` keep the reference's field names and meaning (position, species, density,
temperature, energy[1], model_matrix, N, d, box, neighbour_list, bonds ...) so code written against the
reference reads the same.  They hold HOST copies only; all energies are computed on the GPU through the
C ABI (``DeviceContext``) -- there is no CPU energy path in this package.

Neighbour-list types (``EmptyList``, ``LinkedList``, ``CellList``, ``VerletList``, src/neighbours.jl) are
accepted as ``list_type`` for source compatibility.  They all give identical energies in the reference
(test/runtests.jl:36-38,90-91); on the device the structure is chosen by system size instead: shared-memory
resident chains (all candidates visited) or sorted cell lists in HBM (PMC_MODE_BOX).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .device import DeviceContext
from .models import flatten_model_matrix, max_cutoff, model_kind

# shared-memory budget of the chain kernels (bytes per CTA) used to choose the device mode
_CHAIN_SMEM_LIMIT = 200 * 1024


class NeighbourList:
    """abstract type NeighbourList (src/neighbours.jl:3)."""


class EmptyList(NeighbourList):
    pass


class LinkedList(NeighbourList):
    pass


class CellList(NeighbourList):
    pass


class VerletList(NeighbourList):
    pass


def fold_back(x, box):
    """fold_back(x, box) = x - fld(x, box) * box (src/utils.jl:12)."""
    x = np.asarray(x, dtype=np.float64)
    return x - np.floor(x / box) * box


class Particles:
    """abstract type Particles <: AriannaSystem (src/ParticlesMC.jl:14)."""

    position: np.ndarray
    species: np.ndarray
    N: int
    d: int

    def __len__(self):
        return self.N

    def __iter__(self):
        return iter(self.position)

    def __getitem__(self, i):
        return self.position[i], self.species[i]


class Atoms(Particles):
    """struct Atoms (src/atoms.jl:18-30)."""

    bonds = None

    def __init__(self, position, species, density, energy, temperature, model_matrix, N, d, box, neighbour_list):
        self.position = position
        self.species = species
        self.density = density
        self.energy = energy
        self.temperature = temperature
        self.model_matrix = model_matrix
        self.N = N
        self.d = d
        self.box = box
        self.neighbour_list = neighbour_list
        self.species_list = None


class Molecules(Particles):
    """struct Molecules (src/molecules.jl:24-41)."""

    def __init__(self, position, species, molecule, molecule_species, start_mol, length_mol, density, temperature,
                 energy, model_matrix, d, N, Nmol, box, neighbour_list, bonds):
        self.position = position
        self.species = species
        self.molecule = molecule
        self.molecule_species = molecule_species
        self.start_mol = start_mol
        self.length_mol = length_mol
        self.density = density
        self.temperature = temperature
        self.energy = energy
        self.model_matrix = model_matrix
        self.d = d
        self.N = N
        self.Nmol = Nmol
        self.box = box
        self.neighbour_list = neighbour_list
        self.bonds = bonds


def get_first_and_counts(vec: Sequence[int]):
    """src/molecules.jl:112-139 (1-based firsts, as in the reference)."""
    firsts, counts = [], []
    if len(vec) == 0:
        return firsts, counts
    current, count = vec[0], 1
    firsts.append(1)
    for i in range(1, len(vec)):
        if vec[i] != current:
            counts.append(count)
            firsts.append(i + 1)
            current, count = vec[i], 1
        else:
            count += 1
    counts.append(count)
    return firsts, counts


def chain_smem_estimate(N: int, d: int, molecules: bool) -> int:
    npad = (N + 31) // 32 * 32
    return npad * (8 * d + 1 + 4 + (2 * L.PMC_MAX_BONDS if molecules else 0)) + 16384


def choose_mode(N: int, d: int, molecules: bool) -> int:
    return L.MODE_CHAINS if chain_smem_estimate(N, d, molecules) <= _CHAIN_SMEM_LIMIT else L.MODE_BOX


def make_context(systems: Sequence[Particles], *, device: int = 0, chain_offset: int = 0, threads: int = 0,
                 mode: Optional[int] = None) -> DeviceContext:
    """Device context holding ``systems`` (all of one shape and model), uploaded but energy not initialised."""
    s0 = systems[0]
    mol = isinstance(s0, Molecules)
    params = flatten_model_matrix(s0.model_matrix)
    for s in systems:
        if (s.N, s.d) != (s0.N, s0.d) or isinstance(s, Molecules) != mol:
            raise ValueError("all chains must have the same N, d and system type")
        # the device holds ONE model table and ONE bond topology for all chains of a context
        if s is not s0 and not np.array_equal(flatten_model_matrix(s.model_matrix), params):
            raise ValueError("all chains of one context must share the model matrix")
        if mol and s is not s0 and (list(map(list, s.bonds)) != list(map(list, s0.bonds)) or
                                    list(s.start_mol) != list(s0.start_mol) or list(s.length_mol) != list(s0.length_mol)):
            raise ValueError("all chains of one context must share the bond topology and molecule layout")
    ns = params.shape[0]
    if mode is None:
        mode = choose_mode(s0.N, s0.d, mol)
    ctx = DeviceContext(len(systems), s0.N, s0.d, ns, model_kind(s0.model_matrix), mode=mode, molecules=mol,
                        device=device, chain_offset=chain_offset, threads=threads)
    try:
        ctx.set_model(params)
        if mol:
            ctx.set_bonds([[j - 1 for j in b] for b in s0.bonds])
            ctx.set_molecules([f - 1 for f in s0.start_mol], s0.length_mol)
        ctx.upload(np.stack([s.position for s in systems]), np.stack([s.species for s in systems]),
                   np.stack([s.box for s in systems]), np.array([s.temperature for s in systems]))
    except Exception:
        ctx.close()
        raise
    return ctx


def _initial_energy(system: Particles, device: int) -> float:
    with make_context([system], device=device) as ctx:
        try:
            ctx.init_energy()
        except L.PMCError as e:
            if e.code == L.PMC_ERR_NONFINITE:
                raise ValueError("Initial configuration has infinite or NaN energy.") from e
            raise
        return float(ctx.energy()[0])


def System(*args, molecule_species=None, list_type=EmptyList, list_parameters=None, device: int = 0,
           compute_energy: bool = True):
    """``System(position, species, density, temperature, model_matrix; list_type)`` -> Atoms (src/atoms.jl:40-58)
    ``System(position, species, molecule, density, temperature, model_matrix, bonds; ...)`` -> Molecules
    (src/molecules.jl:76-96).  ``bonds`` lists 1-based partner indices per site, as the reference stores them.

    The cubic box is recomputed from the density, ``(N / density)^(1/d)``, exactly as the reference does, and
    the initial energy ``sum_i e_i / 2`` is evaluated on the GPU (``compute_energy=False`` defers it to
    ``Simulation``, which initialises all chains in one launch)."""
    if len(args) == 5:
        position, species, density, temperature, model_matrix = args
        molecule = bonds = None
    elif len(args) == 7:
        position, species, molecule, density, temperature, model_matrix, bonds = args
    else:
        raise TypeError("System expects 5 (Atoms) or 7 (Molecules) positional arguments")
    position = np.array(position, dtype=np.float64)
    species = np.array(species, dtype=np.int64)
    assert len(position) == len(species)
    N, d = position.shape
    density, temperature = float(density), float(temperature)
    box = np.full(d, (N / density) ** (1 / d))
    energy = np.zeros(1)
    max_cutoff(model_matrix)  # validates the matrix (atoms.jl:46)
    nl = list_type() if isinstance(list_type, type) else list_type
    if molecule is None:
        system = Atoms(position, species, density, energy, temperature, model_matrix, N, d, box, nl)
    else:
        molecule = np.array(molecule, dtype=np.int64)
        Nmol = len(np.unique(molecule))
        start_mol, length_mol = get_first_and_counts(list(molecule))
        if molecule_species is None:
            molecule_species = np.ones(N, dtype=np.int64)
        system = Molecules(position, species, molecule, molecule_species, start_mol, length_mol, density, temperature,
                           energy, model_matrix, d, N, Nmol, box, nl, [list(b) for b in bonds])
    if compute_energy:
        system.energy[0] = _initial_energy(system, device)
    return system


def compute_energy_particle(system: Particles, ids=None, device: int = 0):
    """compute_energy_particle(system, i) / (system, ids) (src/ParticlesMC.jl:102-112); indices 1-based."""
    with make_context([system], device=device) as ctx:
        e = ctx.local_energy(0)
    if ids is None:
        return e
    if np.isscalar(ids):
        return float(e[int(ids) - 1])
    return e[np.asarray(ids, dtype=np.int64) - 1]


def get_start_end_mol(system: Molecules, i: int):
    """src/molecules.jl:104 (1-based)."""
    return system.start_mol[i - 1], system.start_mol[i - 1] + system.length_mol[i - 1] - 1


def energy(system: Particles) -> float:
    """The ``energy`` callback: energy[1] / N (src/utils.jl:51-53)."""
    return float(system.energy[0]) / len(system)


def bonds_from_pairs(N: int, pairs: np.ndarray) -> List[List[int]]:
    """Bond pair table (1-based, as in the XYZ/EXYZ bond section, src/IO/IO.jl:158-199) -> per-site lists."""
    bonds: List[List[int]] = [[] for _ in range(N)]
    for a, b in np.asarray(pairs, dtype=np.int64):
        bonds[a - 1].append(int(b))
        bonds[b - 1].append(int(a))
    return bonds
