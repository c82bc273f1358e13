"""``DeviceContext`` -- thin object wrapper over one ``pmc_ctx`` of the C ABI (include/pmc_b200.h).

Everything here is marshalling: numpy arrays in the caller's (reference) layout go straight to the
library, which owns all device memory.  The Julia shim (julia/ParticlesMCB200.jl) performs exactly the
same sequence of calls with ``ccall``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib as L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


TRIAL_DTYPE = np.dtype([("kind", np.int32), ("move", np.int32), ("i", np.int32), ("j", np.int32),
                        ("delta", np.float64, (3,)), ("u", np.float64)], align=True)
assert TRIAL_DTYPE.itemsize == C.sizeof(L.Trial)


class DeviceContext:
    def __init__(self, n_chains: int, n_particles: int, dim: int, n_species: int, model_kind: int, *,
                 mode: int = L.MODE_CHAINS, precision: int = L.FP64, molecules: bool = False, device: int = 0,
                 chain_offset: int = 0, threads: int = 0, prefilter: int = 0):
        self.lib = L.load()
        self.n_chains, self.N, self.dim, self.ns = n_chains, n_particles, dim, n_species
        self.mode = mode
        self.n_moves = 0
        cfg = L.Config(device=device, mode=mode, precision=precision, n_chains=n_chains, n_particles=n_particles,
                       dim=dim, n_species=n_species, model_kind=model_kind, molecules=int(molecules),
                       chain_offset=chain_offset, threads=threads, prefilter=prefilter)
        h = C.c_void_p()
        L.check(self.lib.pmc_create(C.byref(cfg), C.byref(h)))
        self._h = h

    # -- life cycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.pmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: Optional[int]):
        """Launch on the given cudaStream_t (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        L.check(self.lib.pmc_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- system definition --------------------------------------------------------------------------
    def set_model(self, params: np.ndarray):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.shape != (self.ns, self.ns, L.PMC_NPAR):
            raise ValueError(f"params must have shape ({self.ns}, {self.ns}, {L.PMC_NPAR})")
        L.check(self.lib.pmc_set_model(self._h, _dp(p)))

    def set_bonds(self, bonds: Sequence[Sequence[int]]):
        off = np.zeros(self.N + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(b) for b in bonds])
        idx = np.asarray([j for b in bonds for j in b] or [0], dtype=np.int32)
        L.check(self.lib.pmc_set_bonds(self._h, _ip(off), _ip(idx)))

    def set_molecules(self, start, length):
        """Molecules.start_mol / length_mol with 0-based starts (needed by MoleculeFlip)."""
        st = np.ascontiguousarray(start, dtype=np.int32)
        ln = np.ascontiguousarray(length, dtype=np.int32)
        L.check(self.lib.pmc_set_molecules(self._h, len(st), _ip(st), _ip(ln)))

    def upload(self, position, species, box, temperature, first: int = 0):
        pos = np.ascontiguousarray(position, dtype=np.float64)
        if pos.ndim == 2:
            pos = pos[None]
        count = pos.shape[0]
        if pos.shape != (count, self.N, self.dim):
            raise ValueError(f"position must have shape (count, {self.N}, {self.dim})")
        sp = np.ascontiguousarray(species, dtype=np.int64).reshape(count, self.N)
        bx = np.ascontiguousarray(np.broadcast_to(np.asarray(box, dtype=np.float64), (count, self.dim)))
        tt = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype=np.float64), (count,)))
        L.check(self.lib.pmc_upload(self._h, first, count, _dp(pos), _lp(sp), _dp(bx), _dp(tt)))

    def upload_raw(self, pos_ptr: int, sp_ptr: int, box_ptr: int, temp_ptr: int, first: int, count: int):
        """Pointer form (pinned host buffers of a host framework): no numpy conversion on the way."""
        L.check(self.lib.pmc_upload(self._h, first, count, C.cast(pos_ptr, C.POINTER(C.c_double)),
                                    C.cast(sp_ptr, C.POINTER(C.c_int64)), C.cast(box_ptr, C.POINTER(C.c_double)),
                                    C.cast(temp_ptr, C.POINTER(C.c_double))))

    def init_energy(self):
        L.check(self.lib.pmc_init_energy(self._h))

    # -- Metropolis ---------------------------------------------------------------------------------
    def set_moves(self, moves: Sequence[dict]):
        arr = (L.MoveSpec * len(moves))()
        for k, m in enumerate(moves):
            if m["kind"] in ("displacement", L.MOVE_DISPLACEMENT):
                arr[k] = L.MoveSpec(L.MOVE_DISPLACEMENT, 0, 0, 0, float(m["prob"]), float(m["sigma"]))
            elif m["kind"] in ("flip", L.MOVE_FLIP):
                arr[k] = L.MoveSpec(L.MOVE_FLIP, 0, 0, 0, float(m["prob"]), 0.0)
            else:
                a, b = m["species"]
                arr[k] = L.MoveSpec(L.MOVE_SWAP, int(a), int(b), 0, float(m["prob"]), 0.0)
        L.check(self.lib.pmc_set_moves(self._h, arr, len(moves)))
        self.n_moves = len(moves)

    def seed(self, seed: int):
        L.check(self.lib.pmc_seed(self._h, C.c_uint64(seed)))

    def run(self, n_trials: int, sync: bool = True):
        L.check(self.lib.pmc_run(self._h, n_trials))
        if sync:
            self.sync()

    def sync(self):
        L.check(self.lib.pmc_sync(self._h))

    def run_traced(self, n_trials: int):
        tot = self.n_chains * n_trials
        trials = np.zeros(tot, dtype=TRIAL_DTYPE)
        acc = np.zeros(tot, dtype=np.uint8)
        dE = np.zeros(tot, dtype=np.float64)
        L.check(self.lib.pmc_run_traced(self._h, n_trials, trials.ctypes.data_as(C.POINTER(L.Trial)),
                                        acc.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dE)))
        shape = (self.n_chains, n_trials)
        return trials.reshape(shape), acc.reshape(shape), dE.reshape(shape)

    def replay(self, trials: np.ndarray):
        tr = np.ascontiguousarray(trials, dtype=TRIAL_DTYPE).reshape(self.n_chains, -1)
        n = tr.shape[1]
        acc = np.zeros(tr.size, dtype=np.uint8)
        dE = np.zeros(tr.size, dtype=np.float64)
        L.check(self.lib.pmc_replay(self._h, n, tr.ctypes.data_as(C.POINTER(L.Trial)),
                                    acc.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dE)))
        return acc.reshape(tr.shape), dE.reshape(tr.shape)

    # -- observables --------------------------------------------------------------------------------
    def energy(self) -> np.ndarray:
        e = np.zeros(self.n_chains)
        L.check(self.lib.pmc_energy(self._h, _dp(e)))
        return e

    def energy_into(self, ptr: int):
        L.check(self.lib.pmc_energy(self._h, C.cast(ptr, C.POINTER(C.c_double))))

    def total_energy(self) -> np.ndarray:
        e = np.zeros(self.n_chains)
        L.check(self.lib.pmc_total_energy(self._h, _dp(e)))
        return e

    def local_energy(self, chain: int = 0) -> np.ndarray:
        e = np.zeros(self.N)
        L.check(self.lib.pmc_local_energy(self._h, chain, _dp(e)))
        return e

    def download(self, first: int = 0, count: Optional[int] = None):
        count = self.n_chains - first if count is None else count
        pos = np.zeros((count, self.N, self.dim))
        sp = np.zeros((count, self.N), dtype=np.int64)
        L.check(self.lib.pmc_download(self._h, first, count, _dp(pos), _lp(sp)))
        return pos, sp

    def download_raw(self, pos_ptr: int, sp_ptr: int, first: int, count: int):
        """Pointer form of ``download`` (pinned host buffers)."""
        L.check(self.lib.pmc_download(self._h, first, count, C.cast(pos_ptr, C.POINTER(C.c_double)),
                                      C.cast(sp_ptr, C.POINTER(C.c_int64))))

    def pair_histogram(self, species_a: int = 0, species_b: int = 0, rmax: float = 3.0, nbins: int = 60) -> np.ndarray:
        """Raw pair-distance counts (i < j, summed over chains); label 0 = any species."""
        h = np.zeros(nbins, dtype=np.uint64)
        L.check(self.lib.pmc_pair_histogram(self._h, species_a, species_b, float(rmax), nbins,
                                            h.ctypes.data_as(C.POINTER(C.c_uint64))))
        return h

    def chain_correlation(self) -> np.ndarray:
        """``chain_correlation`` callback of every chain (src/molecules.jl:224-246)."""
        out = np.zeros(self.n_chains)
        L.check(self.lib.pmc_chain_correlation(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def energy_histogram(self, emin: float, emax: float, nbins: int, per_particle: bool = True) -> np.ndarray:
        """Counts of the chains' running energies (per particle by default) in nbins equal bins on [emin, emax)."""
        h = np.zeros(nbins, dtype=np.uint64)
        L.check(self.lib.pmc_energy_histogram(self._h, float(emin), float(emax), nbins, 1 if per_particle else 0,
                                              h.ctypes.data_as(C.POINTER(C.c_uint64))))
        return h

    def counters(self):
        nm = max(self.n_moves, 1)
        calls = np.zeros((self.n_chains, nm), dtype=np.int64)
        acc = np.zeros((self.n_chains, nm), dtype=np.int64)
        L.check(self.lib.pmc_counters(self._h, _lp(calls), _lp(acc)))
        return calls, acc

    # -- one large box over several GPUs (slabs of the cell grid, halo pushes over NVLink peer memory) -----------
    def box_peer_handle(self) -> np.ndarray:
        h = np.zeros(64, dtype=np.uint8)
        L.check(self.lib.pmc_box_peer_export(self._h, h.ctypes.data_as(C.POINTER(C.c_uint8))))
        return h

    def box_peer_attach(self, rank: int, world: int, handles: np.ndarray):
        hs = np.ascontiguousarray(handles, dtype=np.uint8).reshape(world, 64)
        L.check(self.lib.pmc_box_peer_attach(self._h, rank, world, hs.ctypes.data_as(C.POINTER(C.c_uint8))))

    def launch_count(self) -> int:
        return int(self.lib.pmc_launch_count(self._h))

    def work_counters(self, enable: int):
        """pmc_work_counters: 1 = reset and start counting, 0 = stop and read, 2 = read.  Returns
        (fp64-evaluated candidates, trial evaluations) for 0 / 2."""
        out = np.zeros(4, dtype=np.uint64)
        L.check(self.lib.pmc_work_counters(self._h, int(enable), out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return int(out[0]), int(out[1])

    def last_run_ms(self) -> float:
        ms = C.c_float()
        L.check(self.lib.pmc_last_run_ms(self._h, C.byref(ms)))
        return float(ms.value)


def measure_fma_peak(fp64: bool = True, device: int = 0) -> float:
    """Burst FMA throughput of the CUDA-core pipe in TFLOP/s (roofline denominator)."""
    t = C.c_double()
    L.check(L.load().pmc_measure_fma_peak(device, int(fp64), C.byref(t)))
    return float(t.value)
