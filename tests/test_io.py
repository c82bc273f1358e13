"""File formats either side of the hot path (particlesmc_b200/io.py): the reference's XYZ / EXYZ / LAMMPS grammar
(src/IO/*.jl).  Host-side only; systems are built with compute_energy=False so no GPU is needed."""
import io as _io
import os

import numpy as np
import pytest

from particlesmc_b200 import io as IO
from particlesmc_b200 import models as M
from particlesmc_b200.systems import Atoms, Molecules, System

REF = "/root/reference/test"


def make_atoms(d, N=12, seed=0):
    rng = np.random.default_rng(seed)
    sp = rng.integers(1, 4, N)
    pos = rng.uniform(0, 3.0, (N, d))
    return System(pos, sp, 0.9, 0.75, M.JBB(), compute_energy=False)


def make_molecules(nmol=5):
    rng = np.random.default_rng(1)
    N = 3 * nmol
    pos = rng.uniform(0, 3.0, (N, 3))
    sp = np.tile([1, 2, 3], nmol)
    mol = np.repeat(np.arange(1, nmol + 1), 3)
    bonds = [[] for _ in range(N)]
    for m in range(nmol):
        a = 3 * m + 1
        for i, j in ((a, a + 1), (a + 1, a + 2)):
            bonds[i - 1].append(j)
            bonds[j - 1].append(i)
    return System(pos, sp, mol, 1.2, 2.0, M.Trimer(), bonds, compute_energy=False)


@pytest.mark.parametrize("fmt", [IO.XYZ(), IO.EXYZ(), IO.LAMMPS()])
@pytest.mark.parametrize("d", [2, 3])
def test_atoms_roundtrip_all_formats(fmt, d):
    s = make_atoms(d)
    buf = _io.StringIO()
    IO.store_trajectory(buf, s, 7, fmt)
    IO.store_trajectory(buf, s, 8, fmt)  # two frames: the second is selected with m = 2
    lines = buf.getvalue().splitlines()
    for m in (1, 2):
        c = IO.load_configuration(lines, fmt, m=m)
        assert c["N"] == s.N and c["d"] == d
        assert np.allclose(c["box"], s.box, rtol=1e-15)
        assert np.array_equal(c["species"], s.species)
        assert np.max(np.abs(c["position"] - s.position)) <= 0.5e-6  # six decimals
    assert len(lines) == 2 * (s.N + (9 if isinstance(fmt, IO.LAMMPS) else 2))


def test_header_grammar_matches_the_reference_strings():
    s = make_atoms(2)
    b = _io.StringIO()
    IO.write_header(b, s, 3, IO.XYZ())
    n, meta = b.getvalue().splitlines()
    assert n == "12"
    L = repr(float(s.box[0]))
    assert meta == f"step:3 columns:species,position dt:1 cell:{L},{L} rho:0.9 T:0.75"  # xyz.jl:79-84
    b = _io.StringIO()
    IO.write_header(b, s, 3, IO.EXYZ())
    assert b.getvalue().splitlines()[1] == f'Lattice="{L} 0.0 0.0 0.0 {L} 0.0 0.0 0.0 0.0" Properties=:species:S:1:pos:R:2 Time=3'
    b = _io.StringIO()
    IO.write_header(b, s, 3, IO.LAMMPS())
    assert b.getvalue().splitlines() == ["ITEM: TIMESTEP", "3", "ITEM: NUMBER OF ATOMS", "12", "ITEM: BOX BOUNDS pp pp pp",
                                         f"0.0 {L}", f"0.0 {L}", "-0.1 0.1", "ITEM: ATOMS  type x y"]  # lammps.jl:88-105
    m = make_molecules()
    b = _io.StringIO()
    IO.write_header(b, m, 0, IO.XYZ())
    assert "columns:molecule,species,position" in b.getvalue()
    b = _io.StringIO()
    IO.write_header(b, m, 0, IO.EXYZ())
    assert "Properties=molecule:I:1:species:S:1:pos:R:3 Time=0" in b.getvalue()


@pytest.mark.parametrize("fmt", [IO.XYZ(), IO.EXYZ()])
def test_molecules_lastframe_roundtrip_with_bonds(fmt):
    s = make_molecules()
    buf = _io.StringIO()
    IO.store_lastframe(buf, s, 0, fmt)
    lines = buf.getvalue().splitlines()
    assert lines[s.N + 2] == "10" and lines[s.N + 3] == ("columns:bond" if isinstance(fmt, IO.XYZ) else "Properties=bond:I:2")
    c = IO.load_configuration(lines, fmt)
    assert np.array_equal(c["molecule"], s.molecule) and np.array_equal(c["species"], s.species)
    assert [sorted(b) for b in c["bond"]] == [sorted(b) for b in s.bonds]
    with pytest.raises(ValueError, match="frame index"):
        IO.load_configuration(lines, fmt, m=2)
    with pytest.raises(ValueError, match="does not support bonds"):
        IO.store_lastframe(_io.StringIO(), s, 0, IO.LAMMPS())


def test_load_chains_overrides_and_replicas(tmp_path):
    s = make_atoms(2, N=20)
    p = tmp_path / "conf"
    p.mkdir()
    for k in range(2):
        with open(p / f"c{k}.xyz", "w") as f:
            IO.store_trajectory(f, s, 0, IO.XYZ())
    with pytest.raises(KeyError, match="model"):
        IO.load_chains(str(p), compute_energy=False)
    chains = IO.load_chains(str(p), args=dict(model="JBB", nsim=3, temperature=0.3), filename=".xyz", compute_energy=False)
    assert len(chains) == 6 and all(isinstance(c, Atoms) for c in chains)
    assert all(c.temperature == 0.3 and abs(c.density - 0.9) < 1e-5 for c in chains)
    assert np.all(chains[0].position >= 0) and np.all(chains[0].position < chains[0].box)
    dense = IO.load_chains(str(p / "c0.xyz"), args=dict(model="JBB", density=1.8), compute_energy=False)[0]
    assert abs(dense.density - 1.8) < 1e-12 and np.allclose(dense.box, chains[0].box / np.sqrt(2.0), rtol=1e-5)
    table = {"1-1": dict(name="LennardJones", epsilon=1.0, sigma=1.0), "1-2": dict(name="LennardJones", epsilon=1.5, sigma=0.8, rcut=2.0),
             "1-3": dict(name="LennardJones", epsilon=1.0, sigma=1.0), "2-2": dict(name="LennardJones", epsilon=0.5, sigma=0.88),
             "2-3": dict(name="LennardJones", epsilon=1.0, sigma=1.0), "3-3": dict(name="LennardJones", epsilon=1.0, sigma=1.0, shift_potential=False)}
    lj = IO.load_chains(str(p / "c0.xyz"), args=dict(model=table), compute_energy=False)[0]
    assert lj.model_matrix[0][1].rcut == 2.0 and lj.model_matrix[1][0].eps == 1.5 and lj.model_matrix[2][2].shift_potential is False


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_fixtures_parse_identically_in_all_three_formats(config0, molecule):
    """test/runtests.jl:22-38 reads config_0 from .xyz, .exyz and .lmp; all three must give the committed fixture."""
    got = [IO.load_configuration(os.path.join(REF, "config_0" + ext)) for ext in (".xyz", ".exyz", ".lmp")]
    for c in got:
        assert c["N"] == 1290 and c["d"] == 2
        assert np.allclose(c["box"], 32.8962, rtol=1e-12)
        assert np.array_equal(c["species"], config0["species"])
        assert np.array_equal(c["position"], got[0]["position"])
    chains = IO.load_chains(os.path.join(REF, "config_0.xyz"), args=dict(model="JBB"), compute_energy=False)
    assert chains[0].temperature == 0.231 and np.allclose(chains[0].position, config0["position"], atol=1e-12)
    for ext in (".xyz", ".exyz"):
        c = IO.load_configuration(os.path.join(REF, "molecule" + ext))
        assert c["N"] == 3000 and np.array_equal(c["molecule"], molecule["molecule"])
        assert [sorted(b) for b in c["bond"]] == [sorted(b) for b in molecule["bonds"]]
    mol = IO.load_chains(os.path.join(REF, "molecule.xyz"), args=dict(model="Trimer"), compute_energy=False)[0]
    assert isinstance(mol, Molecules) and mol.Nmol == 1000 and mol.temperature == 2.0
