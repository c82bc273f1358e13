"""ctypes binding of ``include/pmc_b200.h`` -- the same entry points the Julia shim ``ccall``s.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a context
is created, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

PMC_NPAR = 12
PMC_MAX_SPECIES = 4
PMC_MAX_MOVES = 8
PMC_MAX_BONDS = 6
PMC_OK, PMC_ERR_INVALID, PMC_ERR_CUDA, PMC_ERR_NONFINITE, PMC_ERR_UNSUPPORTED, PMC_ERR_STATE = range(6)
MODE_CHAINS, MODE_BOX = 0, 1
FP64, MIXED = 0, 1
MOVE_DISPLACEMENT, MOVE_SWAP, MOVE_FLIP = 0, 1, 2


class PMCError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[pmc error {code}] {message}")
        self.code = code
        self.message = message


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("mode", C.c_int32), ("precision", C.c_int32), ("n_chains", C.c_int32),
                ("n_particles", C.c_int32), ("dim", C.c_int32), ("n_species", C.c_int32), ("model_kind", C.c_int32),
                ("molecules", C.c_int32), ("chain_offset", C.c_int32), ("threads", C.c_int32),
                ("prefilter", C.c_int32), ("reserved", C.c_int32 * 4)]


class MoveSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species_a", C.c_int32), ("species_b", C.c_int32), ("reserved", C.c_int32),
                ("probability", C.c_double), ("sigma", C.c_double)]


class Trial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("move", C.c_int32), ("i", C.c_int32), ("j", C.c_int32),
                ("delta", C.c_double * 3), ("u", C.c_double)]


# every symbol include/pmc_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "pmc_abi_version", "pmc_last_error", "pmc_create", "pmc_destroy", "pmc_set_stream", "pmc_set_model",
    "pmc_set_bonds", "pmc_set_molecules", "pmc_upload", "pmc_init_energy", "pmc_set_moves", "pmc_seed", "pmc_run", "pmc_sync",
    "pmc_run_traced", "pmc_replay", "pmc_energy", "pmc_total_energy", "pmc_local_energy", "pmc_download",
    "pmc_pair_histogram", "pmc_chain_correlation", "pmc_energy_histogram", "pmc_counters", "pmc_launch_count", "pmc_last_run_ms", "pmc_work_counters", "pmc_measure_fma_peak", "pmc_box_peer_export",
    "pmc_box_peer_attach",
]

_lib = None


def load():
    """Load libpmc_b200.so (built in-tree by ``particlesmc_b200.build``). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PMCError(PMC_ERR_STATE, f"{LIB_PATH} is missing: run `python -m particlesmc_b200.build` "
                                      "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, dp, lp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    u8p = C.POINTER(C.c_uint8)
    L.pmc_abi_version.restype = C.c_int
    L.pmc_last_error.restype = C.c_char_p
    L.pmc_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.pmc_destroy.argtypes = [vp]
    L.pmc_destroy.restype = None
    L.pmc_set_stream.argtypes = [vp, vp]
    L.pmc_set_model.argtypes = [vp, dp]
    L.pmc_set_bonds.argtypes = [vp, ip, ip]
    L.pmc_set_molecules.argtypes = [vp, C.c_int32, ip, ip]
    L.pmc_upload.argtypes = [vp, C.c_int32, C.c_int32, dp, lp, dp, dp]
    L.pmc_init_energy.argtypes = [vp]
    L.pmc_set_moves.argtypes = [vp, C.POINTER(MoveSpec), C.c_int32]
    L.pmc_seed.argtypes = [vp, C.c_uint64]
    L.pmc_run.argtypes = [vp, C.c_int64]
    L.pmc_sync.argtypes = [vp]
    L.pmc_run_traced.argtypes = [vp, C.c_int64, C.POINTER(Trial), u8p, dp]
    L.pmc_replay.argtypes = [vp, C.c_int64, C.POINTER(Trial), u8p, dp]
    L.pmc_energy.argtypes = [vp, dp]
    L.pmc_total_energy.argtypes = [vp, dp]
    L.pmc_local_energy.argtypes = [vp, C.c_int32, dp]
    L.pmc_download.argtypes = [vp, C.c_int32, C.c_int32, dp, lp]
    L.pmc_pair_histogram.argtypes = [vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.POINTER(C.c_uint64)]
    L.pmc_chain_correlation.argtypes = [vp, C.POINTER(C.c_double)]
    L.pmc_energy_histogram.argtypes = [vp, C.c_double, C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
    L.pmc_counters.argtypes = [vp, lp, lp]
    L.pmc_launch_count.argtypes = [vp]
    L.pmc_launch_count.restype = C.c_int64
    L.pmc_last_run_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.pmc_measure_fma_peak.argtypes = [C.c_int32, C.c_int32, dp]
    L.pmc_work_counters.argtypes = [vp, C.c_int32, C.POINTER(C.c_uint64)]
    L.pmc_box_peer_export.argtypes = [vp, u8p]
    L.pmc_box_peer_attach.argtypes = [vp, C.c_int32, C.c_int32, u8p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("pmc_last_error", "pmc_destroy", "pmc_launch_count", "pmc_abi_version"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != PMC_OK:
        raise PMCError(rc, load().pmc_last_error().decode())
