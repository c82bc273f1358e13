"""CPU tests of the host-side mirror of the reference interface (no GPU, no oracle needed)."""
import numpy as np
import pytest

import particlesmc_b200 as P
from particlesmc_b200 import models as M
from particlesmc_b200.moves import pool_to_specs
from particlesmc_b200.sharding import shard_range
from particlesmc_b200.systems import choose_mode, get_first_and_counts
from particlesmc_b200 import _lib as L


def test_kob_andersen_parameters():
    """models.jl:125-133: eps [1 1.5; 1.5 0.5], sigma [1 .8; .8 .88], rcut 2.5 sigma, shifted."""
    mm = P.KobAndersen()
    assert [[m.eps for m in r] for r in mm] == [[1.0, 1.5], [1.5, 0.5]]
    assert [[m.rcut for m in r] for r in mm] == [[2.5, 2.0], [2.0, 2.5 * 0.88]]
    flat = P.flatten_model_matrix(mm)
    assert flat.shape == (2, 2, 12)
    assert flat[0, 1, 2] == 6.0 and flat[0, 1, 3] == 0.8 * 0.8 and flat[0, 1, 1] == 4.0
    assert abs(mm[0][0].potential(mm[0][0].rcut2)) < 1e-16
    assert M.model_kind(mm) == M.MODEL_LJ


def test_other_models():
    b = P.BHHP()
    assert b[1][1].sigma == 1.4 and b[0][0].ndiv2 == 6 and isinstance(b[0][0].ndiv2, int)
    assert abs(b[0][1].potential(b[0][1].rcut2)) < 1e-16
    j = P.JBB()
    assert len(j) == 3 and j[0][2].eps == 0.75 and j[2][2].sigma == 0.94
    assert j[0][1].C4_sig4 == 0.00062012616 / (0.8 * 0.8) ** 2
    t = P.Trimer()
    assert t[0][1].k == 33.241 and t[0][1].r0 == 1.425 and t[0][0].k == 0.0
    assert t[1][2].kr02 == -27.210884 * 1.575 * 1.575 / 2
    assert np.isinf(t[0][1].bond_potential(1.425 ** 2 * 1.01))
    lj = P.get_model({"name": "LennardJones", "epsilon": 1.1523, "sigma": 1.0339, "rcut": 4.0, "shift_potential": False})
    assert lj.shift == 0.0 and lj.rcut2 == 16.0
    with pytest.raises(ValueError):
        P.get_model({"name": "Nope"})


def test_build_schedule():
    assert P.build_schedule(100, 0, [0, 1, 2, 4, 8])[:8] == [0, 1, 2, 4, 8, 9, 10, 12]
    assert P.build_schedule(100, 0, [0, 1, 2, 4, 8])[-1] == 100
    assert P.build_schedule(100, 0, 10) == list(range(0, 101, 10))
    assert P.build_schedule(100, 50, 25) == [50, 75, 100]


def test_pool_to_specs_and_swap_constructor():
    class S:
        species = np.array([1, 1, 3, 2, 3, 3])
        d = 2
        temperature = 0.5
    sw = P.DiscreteSwap.from_system([1, 3], S)
    assert sw.particles_per_species == (2, 3) and sw.species == (1, 3)
    pool = (P.Move(P.Displacement(0, np.zeros(2), 0.0), P.SimpleGaussian(), {"sigma": 0.05}, 0.2),
            P.Move(sw, P.DoubleUniform(), [], 0.8))
    specs = pool_to_specs(pool)
    assert specs[0] == {"kind": "displacement", "prob": 0.2, "sigma": 0.05}
    assert specs[1] == {"kind": "swap", "prob": 0.8, "species": (1, 3)}
    # log q is symmetric for both policies (moves.jl:110-112, 231-233)
    a = P.Displacement(1, np.array([0.1, -0.2]), 0.0)
    b = P.Displacement(1, -np.array([0.1, -0.2]), 0.0)
    assert P.log_proposal_density(a, P.SimpleGaussian(), {"sigma": 0.05}, S) == \
        P.log_proposal_density(b, P.SimpleGaussian(), {"sigma": 0.05}, S)
    assert P.log_proposal_density(sw, P.DoubleUniform(), [], S) == -np.log(6)
    assert P.delta_log_target_density(1.0, 2.0, S) == -2.0
    with pytest.raises(NotImplementedError):
        pool_to_specs([P.Move(object(), P.DoubleUniform(), [], 1.0)])


def test_system_helpers():
    assert get_first_and_counts([1, 1, 1, 2, 2, 3]) == ([1, 4, 6], [3, 2, 1])
    assert get_first_and_counts([]) == ([], [])
    assert np.allclose(P.fold_back(np.array([-0.25, 2.5]), 2.0), [1.75, 0.5])
    assert choose_mode(1000, 3, False) == L.MODE_CHAINS
    assert choose_mode(3000, 3, True) == L.MODE_CHAINS
    assert choose_mode(1 << 20, 3, False) == L.MODE_BOX
    assert P.bonds_from_pairs(3, np.array([[1, 2], [1, 3], [2, 3]])) == [[2, 3], [1, 3], [1, 2]]


def test_shard_ranges_cover_all_chains():
    for n, w in [(4096, 8), (10, 4), (3, 8), (4096, 1)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (o0, c0), (o1, _) in zip(parts, parts[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
