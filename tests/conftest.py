import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_config0():
    """test/config_0.xyz of the reference (N=1290, d=2, ternary) prepared as load_chains does
    (src/IO/IO.jl:239,284): density from the file box, positions folded into [0, L)."""
    g = np.load(os.path.join(GOLDEN, "config_0.npz"))
    pos, sp, box = g["position"], g["species"], g["box"]
    N = len(sp)
    density = N / np.prod(box)
    pos = pos - np.floor(pos / box) * box
    return dict(position=pos, species=sp, density=density, temperature=float(g["temperature"]), N=N, d=2,
                box=np.full(2, (N / density) ** (1 / 2)), ref=float(g["energy_per_particle_ref"]))


def load_molecule():
    """test/molecule.xyz of the reference (1000 trimers)."""
    g = np.load(os.path.join(GOLDEN, "molecule.npz"))
    pos, sp, box = g["position"], g["species"], g["box"]
    N = len(sp)
    density = N / np.prod(box)
    pos = pos - np.floor(pos / box) * box
    bonds = [[] for _ in range(N)]
    for a, b in g["bonds"]:
        bonds[a - 1].append(int(b))  # 1-based partners, as the reference stores them
        bonds[b - 1].append(int(a))
    return dict(position=pos, species=sp, molecule=g["molecule"], bonds=bonds, density=density,
                temperature=float(g["temperature"]), N=N, d=3, box=np.full(3, (N / density) ** (1 / 3)),
                ref=float(g["energy_per_particle_ref"]))


@pytest.fixture(scope="session")
def config0():
    return load_config0()


@pytest.fixture(scope="session")
def molecule():
    return load_molecule()
