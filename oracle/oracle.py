"""ctypes front-end of the CPU oracle (``oracle/pmc_oracle.c``).

TEST INFRASTRUCTURE ONLY: importable from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package
``particlesmc_b200`` never imports this module.

The arithmetic lives in C; this file only marshals arrays.  Indices are 0-based here,
species labels are 1..ns as in the reference arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpmc_oracle.so")
NPAR = 12
EMPTYLIST, LINKEDLIST = 0, 1
MOVE_DISPLACEMENT, MOVE_SWAP, MOVE_FLIP = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with the Makefile committed beside it (gcc, no GPU needed)."""
    src = os.path.join(_HERE, "pmc_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


class _Move(C.Structure):
    _fields_ = [("kind", C.c_int32), ("prob", C.c_double), ("sigma", C.c_double),
                ("spA", C.c_int32), ("spB", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, lp, dp, C.c_double, C.c_int, ip, ip]
        L.orc_destroy.argtypes = [C.c_void_p]
        for name in ("orc_energy", "orc_total_energy"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_local_energy.restype = C.c_double
        L.orc_local_energy.argtypes = [C.c_void_p, C.c_int]
        L.orc_local_energies.argtypes = [C.c_void_p, dp]
        L.orc_get_state.argtypes = [C.c_void_p, dp, lp]
        L.orc_get_ncells.argtypes = [C.c_void_p, ip]
        L.orc_species_count.restype = C.c_int
        L.orc_species_count.argtypes = [C.c_void_p, C.c_int]
        L.orc_species_member.restype = C.c_int
        L.orc_species_member.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_nearest_image_r2.restype = C.c_double
        L.orc_nearest_image_r2.argtypes = [dp, dp, dp, C.c_int]
        L.orc_fold_back.restype = C.c_double
        L.orc_fold_back.argtypes = [C.c_double, C.c_double]
        L.orc_pair_potential.restype = C.c_double
        L.orc_pair_potential.argtypes = [C.c_int, dp, C.c_double]
        L.orc_bond_potential.restype = C.c_double
        L.orc_bond_potential.argtypes = [dp, C.c_double]
        L.orc_lennard_jones.restype = C.c_double
        L.orc_lennard_jones.argtypes = [C.c_double] * 3
        L.orc_step_displacement.restype = C.c_int
        L.orc_step_displacement.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_int, dp, dp]
        L.orc_step_swap.restype = C.c_int
        L.orc_step_swap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp]
        L.orc_step_flip.restype = C.c_int
        L.orc_step_flip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp]
        L.orc_step_swap_draw.restype = C.c_int
        L.orc_step_swap_draw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                         ip, ip, dp, dp]
        L.orc_replay.argtypes = [C.c_void_p, C.c_int64, ip, ip, ip, ip, ip, dp, dp, C.c_int,
                                 C.POINTER(C.c_uint8), dp, dp]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_run.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int64, C.POINTER(_Move), C.c_int,
                              C.c_int, lp, lp]
        L.orc_run_chains.restype = C.c_int
        L.orc_run_chains.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint64, C.c_uint64, C.c_int64,
                                     C.POINTER(_Move), C.c_int, C.c_int, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def make_pool(moves: Sequence[dict]):
    """moves: dicts with kind ('displacement'|'swap'), prob, sigma | species=(A,B)."""
    arr = (_Move * len(moves))()
    for k, m in enumerate(moves):
        if m["kind"] == "displacement":
            arr[k] = _Move(MOVE_DISPLACEMENT, m["prob"], m["sigma"], 0, 0)
        else:
            arr[k] = _Move(MOVE_SWAP, m["prob"], 0.0, m["species"][0], m["species"][1])
    return arr


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    u32 = C.POINTER(C.c_uint32)
    lib().orc_philox4x32_10(c.ctypes.data_as(u32), k.ctypes.data_as(u32), out.ctypes.data_as(u32))
    return out


class OracleSystem:
    """One reference ``Atoms``/``Molecules`` system (atoms.jl:40-58, molecules.jl:76-96)."""

    def __init__(self, position, species, box, temperature, kind: int, params, list_type: int = LINKEDLIST,
                 bonds: Optional[Sequence[Sequence[int]]] = None):
        pos = np.ascontiguousarray(position, dtype=np.float64)
        self.N, self.d = pos.shape
        sp = np.ascontiguousarray(species, dtype=np.int64)
        par = np.ascontiguousarray(params, dtype=np.float64)
        self.ns = par.shape[0]
        assert par.shape == (self.ns, self.ns, NPAR)
        bx = np.ascontiguousarray(np.broadcast_to(np.asarray(box, dtype=np.float64), (self.d,)))
        self.box = bx.copy()
        self.temperature = float(temperature)
        if bonds is not None:
            off = np.zeros(self.N + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(b) for b in bonds])
            idx = np.asarray([j for b in bonds for j in b], dtype=np.int32)
            if idx.size == 0:
                idx = np.zeros(1, dtype=np.int32)
            boff, bidx = _ip(off), _ip(idx)
        else:
            boff = bidx = None
        self._h = lib().orc_create(self.N, self.d, self.ns, kind, _dp(par), _dp(pos), _lp(sp), _dp(bx),
                                   self.temperature, list_type, boff, bidx)
        if not self._h:
            raise ValueError("Initial configuration has infinite or NaN energy.")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- energies ---------------------------------------------------------------
    @property
    def energy(self) -> float:
        """system.energy[1] (running bookkeeping value)."""
        return lib().orc_energy(self._h)

    def total_energy(self) -> float:
        return lib().orc_total_energy(self._h)

    def local_energy(self, i: int) -> float:
        return lib().orc_local_energy(self._h, i)

    def local_energies(self) -> np.ndarray:
        out = np.zeros(self.N)
        lib().orc_local_energies(self._h, _dp(out))
        return out

    def state(self):
        pos = np.zeros((self.N, self.d))
        sp = np.zeros(self.N, dtype=np.int64)
        lib().orc_get_state(self._h, _dp(pos), _lp(sp))
        return pos, sp

    def ncells(self):
        out = np.zeros(3, dtype=np.int32)
        lib().orc_get_ncells(self._h, _ip(out))
        return out[: self.d]

    # -- moves ------------------------------------------------------------------
    def step_displacement(self, i, delta, u, revert_mode=0):
        dl = np.zeros(3)
        dl[: self.d] = delta
        e1, e2 = C.c_double(), C.c_double()
        acc = lib().orc_step_displacement(self._h, int(i), _dp(dl), float(u), revert_mode, C.byref(e1), C.byref(e2))
        return bool(acc), e1.value, e2.value

    def step_swap(self, A, B, i, j, u, revert_mode=0):
        e1, e2 = C.c_double(), C.c_double()
        acc = lib().orc_step_swap(self._h, A, B, int(i), int(j), float(u), revert_mode, C.byref(e1), C.byref(e2))
        return bool(acc), e1.value, e2.value

    def step_swap_draw(self, A, B, ka, kb, u, revert_mode=0):
        """DoubleUniform draw resolved through the oracle's OWN species lists (src/moves.jl:238-241, utils.jl:31-49):
        slots (ka, kb) -> particles (i, j), then the swap.  Returns (accepted, i, j, e1, e2)."""
        e1, e2 = C.c_double(), C.c_double()
        i, j = C.c_int(), C.c_int()
        acc = lib().orc_step_swap_draw(self._h, A, B, int(ka), int(kb), float(u), revert_mode, C.byref(i), C.byref(j),
                                       C.byref(e1), C.byref(e2))
        return bool(acc), i.value, j.value, e1.value, e2.value

    def step_flip(self, i, j, u, revert_mode=0):
        e1, e2 = C.c_double(), C.c_double()
        acc = lib().orc_step_flip(self._h, int(i), int(j), float(u), revert_mode, C.byref(e1), C.byref(e2))
        return bool(acc), e1.value, e2.value

    def replay(self, kind, i, j, spA, spB, delta, u, revert_mode=0):
        n = len(kind)
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        i = np.ascontiguousarray(i, dtype=np.int32)
        j = np.ascontiguousarray(j, dtype=np.int32)
        spA = np.ascontiguousarray(spA, dtype=np.int32)
        spB = np.ascontiguousarray(spB, dtype=np.int32)
        delta = np.ascontiguousarray(delta, dtype=np.float64).reshape(n, 3)
        u = np.ascontiguousarray(u, dtype=np.float64)
        acc = np.zeros(n, dtype=np.uint8)
        dE = np.zeros(n)
        E = np.zeros(n)
        lib().orc_replay(self._h, n, _ip(kind), _ip(i), _ip(j), _ip(spA), _ip(spB), _dp(delta), _dp(u), revert_mode,
                         acc.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dE), _dp(E))
        return acc, dE, E

    def run(self, seed, chain, t0, n_trials, pool, revert_mode=1):
        calls = np.zeros(len(pool), dtype=np.int64)
        accepted = np.zeros(len(pool), dtype=np.int64)
        lib().orc_run(self._h, seed, chain, t0, n_trials, pool, len(pool), revert_mode, _lp(calls), _lp(accepted))
        return calls, accepted


def run_chains(systems: Sequence[OracleSystem], seed, t0, n_trials, pool, revert_mode=1, n_threads=0) -> int:
    """Independent chains over host threads (the reference's `parallel=true`). Returns threads used."""
    hs = (C.c_void_p * len(systems))(*[s._h for s in systems])
    return lib().orc_run_chains(hs, len(systems), seed, t0, n_trials, pool, len(pool), revert_mode, n_threads)


def chain_correlation(species, start_mol, length_mol) -> float:
    """compute_chain_correlation (src/molecules.jl:224-243), numpy restatement; species 1-based labels,
    start_mol 0-based first site of each molecule."""
    species = np.asarray(species)
    length_mol = np.asarray(length_mol)
    n = int(length_mol[0])
    assert np.all(length_mol == n), "All chains must have the same length"
    assert n > 1, "Chains must have at least two particles"
    nmol = len(start_mol)
    arr = np.zeros((nmol, n))
    for m, s0 in enumerate(start_mol):
        arr[m, :] = species[s0:s0 + n]
    arr[arr == 2] = -1
    cross = [np.sum(arr[:, i] * arr[:, j]) / nmol for i in range(n - 1) for j in range(i + 1, n)]
    return float(np.sum(np.square(cross)))
