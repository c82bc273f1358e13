"""Regenerates tests/golden/*.npz from the reference's own test fixtures.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_fixtures.py

Inputs (data files of the reference's test suite, not source code):
  test/config_0.xyz   N=1290, d=2, ternary, cell 32.8962^2      (test/runtests.jl:22-38)
  test/molecule.xyz   3000 sites = 1000 trimers + 3000 bonds     (test/runtests.jl:136-149)
  examples/lj-mixture/calculated-energies.csv  23 state points   (statistical golden data)
The positions are stored exactly as written in the files (no fold): folding into [0, L) is part
of the path under test (IO.jl:284 -> utils.jl:12).  Known answers pinned by the reference:
  config_0 + JBB     : energy/N = -2.676832        (atol 1e-6)   test/runtests.jl:36-38
  molecule + Trimer  : energy/N = 25.65865662277199 (atol 1e-6)  test/runtests.jl:148-149
"""
import csv
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_xyz(path):
    with open(path) as f:
        lines = f.read().split("\n")
    n = int(lines[0])
    meta = lines[1].split()
    cols = [m for m in meta if m.startswith("columns:")][0].split(":")[1].split(",")
    cell = [m for m in meta if m.startswith("cell:")][0].split(":")[1].split(",")
    box = np.array([float(c) for c in cell])
    d = len(box)
    rows = [l.split() for l in lines[2:2 + n]]
    k = 0
    molecule = None
    if "molecule" in cols:
        molecule = np.array([int(r[0]) for r in rows], dtype=np.int64)
        k = 1
    species = np.array([int(r[k]) for r in rows], dtype=np.int64)
    pos = np.array([[float(v) for v in r[k + 1:k + 1 + d]] for r in rows], dtype=np.float64)
    bonds = None
    rest = [l for l in lines[2 + n:] if l.strip()]
    if rest:
        nb = int(rest[0])
        assert rest[1].startswith("columns:bond")
        bonds = np.array([[int(v) for v in l.split()[:2]] for l in rest[2:2 + nb]], dtype=np.int64)
    return dict(N=n, d=d, box=box, species=species, position=pos, molecule=molecule, bonds=bonds)


def main():
    c = read_xyz(os.path.join(REF, "test/config_0.xyz"))
    np.savez_compressed(os.path.join(OUT, "config_0.npz"), position=c["position"], species=c["species"], box=c["box"],
                        temperature=0.231, energy_per_particle_ref=-2.676832, atol_ref=1e-6)
    m = read_xyz(os.path.join(REF, "test/molecule.xyz"))
    np.savez_compressed(os.path.join(OUT, "molecule.npz"), position=m["position"], species=m["species"], box=m["box"],
                        molecule=m["molecule"], bonds=m["bonds"], temperature=2.0,
                        energy_per_particle_ref=25.65865662277199, atol_ref=1e-6)
    rows = []
    with open(os.path.join(REF, "examples/lj-mixture/calculated-energies.csv")) as f:
        for r in csv.DictReader(f):
            rows.append([float(r[k]) for k in ("t", "x", "density", "energy", "energy_err",
                                               "acceptance_rate_displacement", "acceptance_rate_swap")])
    np.savez_compressed(os.path.join(OUT, "lj_mixture_table.npz"), table=np.array(rows),
                        columns=np.array(["t", "x", "density", "energy", "energy_err", "acc_displacement", "acc_swap"]))
    print("config_0:", c["N"], c["d"], c["box"], np.bincount(c["species"]))
    print("molecule:", m["N"], m["d"], m["box"], len(m["bonds"]))
    print("lj table:", len(rows))


if __name__ == "__main__":
    main()
