"""Edge cases of the hot path through the C ABI: degenerate sizes, boxes smaller than two cutoffs (the reference's
own tiny system, test/ss142d.jl), kernel-selection boundaries, many chains with different thermodynamic states."""
import numpy as np
import pytest

from oracle import oracle as O
from particlesmc_b200 import models as M
from particlesmc_b200.device import DeviceContext
from particlesmc_b200.synthetic import ka_lattice, lattice

pytestmark = pytest.mark.gpu


def trace_vs_oracle(ctx, orc, n, labels=None):
    tr, acc, dE = ctx.run_traced(n)
    t = tr[0]
    labels = labels or {0: (0, 0)}
    spA = np.array([labels[m][0] for m in t["move"]], dtype=np.int32)
    spB = np.array([labels[m][1] for m in t["move"]], dtype=np.int32)
    o_acc, o_dE, _ = orc.replay(t["kind"], t["i"], np.maximum(t["j"], 0), spA, spB, t["delta"], t["u"], 1)
    assert np.array_equal(o_acc, acc[0])
    assert abs(ctx.energy()[0] - orc.energy) <= 1e-11 * max(1.0, abs(orc.energy))
    return acc[0]


def test_single_particle_and_pair():
    par = M.flatten_model_matrix(M.KobAndersen())
    with DeviceContext(1, 1, 3, 2, M.MODEL_LJ) as ctx:  # one particle: no partners, every move accepted
        ctx.set_model(par)
        ctx.upload(np.array([[0.3, 0.4, 0.5]]), np.array([1]), [5.0] * 3, 1.0)
        ctx.init_energy()
        assert ctx.energy()[0] == 0.0
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.1)])
        ctx.seed(1)
        _, acc, dE = ctx.run_traced(200)
        assert acc.all() and np.all(dE == 0.0)
        pos, _ = ctx.download()
        assert np.all(np.isfinite(pos))
    pos = np.array([[1.0, 1.0, 1.0], [2.1, 1.0, 1.0]])
    sp = np.array([1, 2])
    orc = O.OracleSystem(pos, sp, [6.0] * 3, 0.7, M.MODEL_LJ, par, O.LINKEDLIST)
    with DeviceContext(1, 2, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(pos, sp, [6.0] * 3, 0.7)
        ctx.init_energy()
        assert abs(ctx.energy()[0] - orc.energy) < 1e-14
        ctx.set_moves([dict(kind="displacement", prob=0.7, sigma=0.1), dict(kind="swap", prob=0.3, species=(1, 2))])
        ctx.seed(2)
        trace_vs_oracle(ctx, orc, 400, {0: (0, 0), 1: (1, 2)})


def test_box_smaller_than_two_cutoffs_ss142d():
    """test/ss142d.jl:10-24: BHHP soft spheres, d=2, N=8, rho=0.5 (L=4 < 2 rcut=7): one cell, minimum image only,
    100 chains; sigma=0.065."""
    N, nch = 8, 100
    rng = np.random.default_rng(4)
    mm = M.BHHP()
    par = M.flatten_model_matrix(mm)
    pos, sp, box = lattice(N, 2, 0.5, seed=1, fractions=(0.5, 0.5))
    allpos = np.stack([pos + rng.normal(0, 0.05, pos.shape) for _ in range(nch)])
    with DeviceContext(nch, N, 2, 2, M.MODEL_SOFT) as ctx:
        ctx.set_model(par)
        ctx.upload(allpos, np.stack([sp] * nch), box, 1.0)
        ctx.init_energy()
        e = ctx.energy()
        for c in (0, 57, 99):
            p = allpos[c] - np.floor(allpos[c] / box) * box
            orc = O.OracleSystem(p, sp, box, 1.0, M.MODEL_SOFT, par, O.LINKEDLIST)
            assert list(orc.ncells()) == [1, 1]
            assert abs(e[c] - orc.energy) <= 1e-12 * abs(orc.energy)
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.065)])
        ctx.seed(3)
        p0 = allpos[0] - np.floor(allpos[0] / box) * box
        orc0 = O.OracleSystem(p0, sp, box, 1.0, M.MODEL_SOFT, par, O.EMPTYLIST)
        acc = trace_vs_oracle(ctx, orc0, 3000)
        assert 0.2 < acc.mean() < 0.99
        assert np.max(np.abs(ctx.energy() - ctx.total_energy()) / np.abs(ctx.total_energy())) < 1e-11


@pytest.mark.parametrize("N", [1024, 1025, 2000])
def test_kernel_selection_boundaries(N):
    """N = 1024 is the last size of the register-resident kernel, 1025 the first of the general one."""
    pos, sp, box = ka_lattice(N, 1.2, seed=N)
    pos = pos + np.random.default_rng(N).normal(0, 0.04, pos.shape)
    pos -= np.floor(pos / box) * box
    par = M.flatten_model_matrix(M.KobAndersen())
    orc = O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, par, O.LINKEDLIST)
    with DeviceContext(1, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(pos, sp, box, 1.0)
        ctx.init_energy()
        assert abs(ctx.energy()[0] - orc.energy) <= 1e-12 * abs(orc.energy)
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(N)
        trace_vs_oracle(ctx, orc, 1200)


def test_chains_with_different_states_do_not_interfere():
    """Per-chain box, temperature and composition (load_chains allows per-chain temperature/density,
    src/IO/IO.jl:255-268): each chain must evolve exactly as it does alone."""
    par = M.flatten_model_matrix(M.KobAndersen())
    N = 343
    cfgs = []
    for k, (rho, T, fa) in enumerate([(1.2, 1.0, 0.8), (0.9, 0.4, 0.5), (0.6, 2.5, 0.2)]):
        pos, sp, box = lattice(N, 3, rho, seed=k, fractions=(fa, 1 - fa))
        cfgs.append((pos, sp, box, T))
    pool = [dict(kind="displacement", prob=0.9, sigma=0.06), dict(kind="swap", prob=0.1, species=(1, 2))]
    with DeviceContext(3, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), np.stack([c[2] for c in cfgs]),
                   [c[3] for c in cfgs])
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(5)
        ctx.run(3000)
        together, sp_t = ctx.download()
        e_t = ctx.energy()
    for k, (pos, sp, box, T) in enumerate(cfgs):
        with DeviceContext(1, N, 3, 2, M.MODEL_LJ, chain_offset=k) as one:
            one.set_model(par)
            one.upload(pos, sp, box, T)
            one.init_energy()
            one.set_moves(pool)
            one.seed(5)
            one.run(3000)
            p1, s1 = one.download()
            assert np.array_equal(p1[0], together[k]) and np.array_equal(s1[0], sp_t[k])
            assert one.energy()[0] == e_t[k]
