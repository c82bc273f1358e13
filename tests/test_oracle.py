"""CPU tests of the oracle (oracle/pmc_oracle.c) against the reference's own known answers.

These pin the oracle: everything the GPU parity tests compare against is first checked here against the
golden vectors the reference's test-suite holds (test/runtests.jl:22-38, :136-149, :90-91, :129).
"""
import numpy as np
import pytest

from oracle import oracle as O
from particlesmc_b200 import models as M
from particlesmc_b200.synthetic import ka_lattice


def zero_based(bonds):
    return [[j - 1 for j in b] for b in bonds]


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [hex(v) for v in O.philox4x32_10([0] * 4, [0] * 2)] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(v) for v in O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(v) for v in O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                            [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_config0_known_answer(config0):
    """test/runtests.jl:36-38: JBB energy per particle -2.676832 (atol 1e-6) for every list type."""
    par = M.flatten_model_matrix(M.JBB())
    es = []
    for lt in (O.EMPTYLIST, O.LINKEDLIST):
        s = O.OracleSystem(config0["position"], config0["species"], config0["box"], config0["temperature"],
                           M.MODEL_SMOOTHLJ, par, lt)
        es.append(s.energy / config0["N"])
        assert abs(es[-1] - config0["ref"]) < 1e-6
    assert abs(es[0] - es[1]) < 1e-13
    assert list(s.ncells()) == [13, 13]  # 32.8962 / 2.5 -> 13 cells per side


def test_molecule_known_answer(molecule):
    """test/runtests.jl:148-149: Trimer energy per site 25.65865662277199 (atol 1e-6)."""
    par = M.flatten_model_matrix(M.Trimer())
    for lt in (O.EMPTYLIST, O.LINKEDLIST):
        s = O.OracleSystem(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"],
                           M.MODEL_KG, par, lt, bonds=zero_based(molecule["bonds"]))
        assert abs(s.energy / molecule["N"] - molecule["ref"]) < 1e-6
        assert abs(s.energy / molecule["N"] - 25.65865662277199) < 1e-12


def test_nearest_image_matches_numpy():
    rng = np.random.default_rng(1)
    box = np.array([3.0, 4.5, 7.25])
    for _ in range(200):
        xi, xj = rng.uniform(-10, 10, 3), rng.uniform(-10, 10, 3)
        dx = xi - xj
        dx = dx - np.round(dx / box) * box
        import ctypes as C
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        r2 = O.lib().orc_nearest_image_r2(dp(xi), dp(xj), dp(box), 3)
        assert r2 == (dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]
    assert O.lib().orc_fold_back(-0.25, 2.0) == 1.75


@pytest.mark.parametrize("name", ["KobAndersen", "BHHP", "JBB", "Trimer"])
def test_potentials_match_host_models(name):
    """The C potentials and the Python mirror of models.jl evaluate the same formulas."""
    import ctypes as C
    mm = M.NAMED_MODELS[name]()
    kind = M.model_kind(mm)
    rng = np.random.default_rng(2)
    for row in mm:
        for m in row:
            p = m.flat()
            pp = p.ctypes.data_as(C.POINTER(C.c_double))
            for r2 in rng.uniform(0.7, m.rcut2, 20):
                ref = m.potential(float(r2))
                got = O.lib().orc_pair_potential(kind, pp, float(r2))
                assert abs(got - ref) <= 4e-16 * max(1.0, abs(ref))
            if name == "Trimer" and m.r0 > 0:
                for r2 in rng.uniform(0.8, m.r02 * 0.99, 10):
                    assert abs(O.lib().orc_bond_potential(pp, float(r2)) - m.bond_potential(float(r2))) < 1e-12
                assert np.isinf(O.lib().orc_bond_potential(pp, m.r02 * 1.01))
    if name == "KobAndersen":  # models.jl:125-133: shifted at 2.5 sigma
        m = mm[0][1]
        assert m.rcut == 2.5 * 0.8 and abs(m.potential(m.rcut2)) < 1e-15


def _run_pair(pos, sp, box, T, kind, par, pool, n_trials, bonds=None, seed=10):
    out = []
    for lt in (O.EMPTYLIST, O.LINKEDLIST):
        s = O.OracleSystem(pos, sp, box, T, kind, par, lt, bonds=bonds)
        es = []
        for blk in range(4):
            s.run(seed, 0, blk * n_trials, n_trials, pool, revert_mode=0)
            es.append(s.energy)
        out.append((np.array(es), s))
    return out


def test_list_equivalence_displacement(config0):
    """test/runtests.jl:40-91: same seed => EmptyList and LinkedList energy series agree (atol 1e-6)."""
    par = M.flatten_model_matrix(M.JBB())
    pool = O.make_pool([dict(kind="displacement", prob=1.0, sigma=0.05)])
    (e0, s0), (e1, s1) = _run_pair(config0["position"], config0["species"], config0["box"], config0["temperature"],
                                   M.MODEL_SMOOTHLJ, par, pool, 2 * config0["N"])
    assert np.allclose(e0 / config0["N"], e1 / config0["N"], atol=1e-6, rtol=0)
    assert np.array_equal(s0.state()[0], s1.state()[0])
    # bookkeeping energy tracks the recomputed one
    assert abs(s1.energy - s1.total_energy()) < 1e-8


def test_list_equivalence_swaps(config0):
    """test/runtests.jl:93-129: Displacement 0.2 + DiscreteSwap (1,3) 0.4 + (2,3) 0.4."""
    par = M.flatten_model_matrix(M.JBB())
    pool = O.make_pool([dict(kind="displacement", prob=0.2, sigma=0.05), dict(kind="swap", prob=0.4, species=(1, 3)),
                        dict(kind="swap", prob=0.4, species=(2, 3))])
    (e0, s0), (e1, s1) = _run_pair(config0["position"], config0["species"], config0["box"], config0["temperature"],
                                   M.MODEL_SMOOTHLJ, par, pool, config0["N"])
    assert np.allclose(e0 / config0["N"], e1 / config0["N"], atol=1e-6, rtol=0)
    assert np.array_equal(s0.state()[1], s1.state()[1])
    sp = s1.state()[1]
    assert np.bincount(sp)[1:].tolist() == [600, 330, 360]  # swaps conserve composition
    for A in (1, 2, 3):  # SpeciesList stays consistent (utils.jl:31-49, moves.jl:175-179)
        members = sorted(O.lib().orc_species_member(s1._h, A, k) for k in range(O.lib().orc_species_count(s1._h, A)))
        assert members == sorted(np.nonzero(sp == A)[0].tolist())
    assert abs(s1.energy - s1.total_energy()) < 1e-8


def test_list_equivalence_molecules(molecule):
    """test/runtests.jl:151-191."""
    par = M.flatten_model_matrix(M.Trimer())
    pool = O.make_pool([dict(kind="displacement", prob=1.0, sigma=0.05)])
    (e0, _), (e1, s1) = _run_pair(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"],
                                  M.MODEL_KG, par, pool, 1500, bonds=zero_based(molecule["bonds"]))
    assert np.allclose(e0 / molecule["N"], e1 / molecule["N"], atol=1e-6, rtol=0)
    assert abs(s1.energy - s1.total_energy()) < 1e-7


def test_revert_modes_and_rejection():
    """Reference revert (x+d)+(-d) vs exact restore differ by rounding only; an overlap is rejected."""
    pos, sp, box = ka_lattice(216, 1.2, seed=3)
    par = M.flatten_model_matrix(M.KobAndersen())
    a = O.OracleSystem(pos, sp, box, 0.5, M.MODEL_LJ, par, O.LINKEDLIST)
    b = O.OracleSystem(pos, sp, box, 0.5, M.MODEL_LJ, par, O.LINKEDLIST)
    pool = O.make_pool([dict(kind="displacement", prob=1.0, sigma=0.1)])
    ca, aa = a.run(7, 0, 0, 2000, pool, revert_mode=0)
    cb, ab = b.run(7, 0, 0, 2000, pool, revert_mode=1)
    assert ca[0] == cb[0] == 2000 and aa[0] == ab[0] and 0 < aa[0] < 2000
    assert np.allclose(a.state()[0], b.state()[0], atol=1e-12, rtol=0)
    # move particle 0 on top of particle 1: e2 = +Inf (or NaN) => rejected, state untouched
    x0, x1 = b.state()[0][0].copy(), b.state()[0][1].copy()
    acc, e1, e2 = b.step_displacement(0, x1 - x0, 0.0, revert_mode=1)
    assert not acc and not np.isfinite(e2)
    assert np.array_equal(b.state()[0][0], x0)


def test_initial_overlap_raises():
    pos, sp, box = ka_lattice(64, 1.2, seed=1)
    pos[1] = pos[0]
    with pytest.raises(ValueError, match="infinite or NaN"):
        O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, M.flatten_model_matrix(M.KobAndersen()), O.EMPTYLIST)


def test_small_box_stencil_dedup():
    """Fewer than 3 cells per side: the stencil de-duplicates (neighbours.jl:106-108) and equals all pairs."""
    pos, sp, box = ka_lattice(125, 1.2, seed=5)  # L = 4.70 -> 1 cell of side >= 2.5
    par = M.flatten_model_matrix(M.KobAndersen())
    rng = np.random.default_rng(0)
    pos = pos + rng.normal(0, 0.05, pos.shape)
    a = O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, par, O.EMPTYLIST)
    b = O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, par, O.LINKEDLIST)
    assert list(b.ncells()) == [1, 1, 1]
    assert np.allclose(a.local_energies(), b.local_energies(), rtol=1e-13, atol=1e-13)


def test_molecule_flip_is_reversible(molecule):
    """MoleculeFlip (moves.jl:301-323): exchanging the species of two unlike sites of a trimer twice restores the
    energy; a rejected flip leaves species and energy untouched."""
    par = M.flatten_model_matrix(M.Trimer())
    s = O.OracleSystem(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"],
                       M.MODEL_KG, par, O.LINKEDLIST, bonds=zero_based(molecule["bonds"]))
    e0 = s.energy
    acc, e1, e2 = s.step_flip(0, 2, 0.0, revert_mode=0)          # u = 0: accepted unless the energy is infinite
    if acc:
        assert s.state()[1][0] == 3 and s.state()[1][2] == 1
        assert abs(s.energy - s.total_energy()) < 1e-8
        acc2, f1, f2 = s.step_flip(0, 2, 0.0, revert_mode=0)
        assert acc2 and abs(s.energy - e0) < 1e-8 and abs((f2 - f1) + (e2 - e1)) < 1e-8
    acc3, g1, g2 = s.step_flip(3, 4, 1.0, revert_mode=0)        # u = 1: always rejected
    assert not acc3 and s.state()[1][3] == 1 and s.state()[1][4] == 2 and abs(s.energy - e0) < 1e-8
