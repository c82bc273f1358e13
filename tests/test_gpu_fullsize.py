"""BASELINE.json's full sizes (config 2: 4096 chains x N = 1000; config 3: one box of N = 2^20; configs 4 and 5 at the
chain counts bench/other_configs.py times), checked through
size-independent properties: the running energy equals a recomputation from the final configuration, chains do not
influence each other (a subset run alone reproduces the same chains bit for bit), composition and counters are
conserved, the same seed gives the same trajectory."""
import numpy as np
import pytest

from oracle import oracle as O
from particlesmc_b200 import _lib as L
from particlesmc_b200 import models as M
from particlesmc_b200.device import DeviceContext
from particlesmc_b200.synthetic import ka_lattice

pytestmark = pytest.mark.gpu


def ka_chains(ctx, pos, sp, box, n, T=1.0):
    ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
    ctx.upload(np.broadcast_to(pos, (n,) + pos.shape).copy(), np.broadcast_to(sp, (n,) + sp.shape).copy(), box, T)
    ctx.init_energy()
    ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
    ctx.seed(42)


def test_config2_4096_chains_of_1000():
    N, nch, sweeps = 1000, 4096, 3
    pos, sp, box = ka_lattice(N, 1.2, seed=0)
    with DeviceContext(nch, N, 3, 2, M.MODEL_LJ) as ctx:
        ka_chains(ctx, pos, sp, box, nch)
        e0 = ctx.energy()
        assert np.all(e0 == e0[0])  # identical replicas, identical initial energies
        ctx.run(sweeps * N)
        e_run, e_tot = ctx.energy(), ctx.total_energy()
        assert np.max(np.abs(e_run - e_tot) / np.abs(e_tot)) < 1e-11
        calls, acc = ctx.counters()
        assert np.all(calls[:, 0] == sweeps * N) and np.all(acc[:, 0] > 0) and np.all(acc[:, 0] < sweeps * N)
        assert len(np.unique(e_run)) == nch  # every chain drew its own stream
        p_all, s_all = ctx.download(0, nch)
        assert np.array_equal(s_all, np.broadcast_to(sp, s_all.shape))
        # per-chain checksum of the coordinates, and a checksum of the checksums
        chk = np.frombuffer(np.ascontiguousarray(p_all).tobytes(), dtype=np.uint64).reshape(nch, -1).sum(axis=1)
    # the last 5 chains alone (global indices through chain_offset): same streams, same coordinates
    with DeviceContext(5, N, 3, 2, M.MODEL_LJ, chain_offset=nch - 5) as sub:
        ka_chains(sub, pos, sp, box, 5)
        sub.run(sweeps * N)
        p_sub, _ = sub.download(0, 5)
        chk_sub = np.frombuffer(np.ascontiguousarray(p_sub).tobytes(), dtype=np.uint64).reshape(5, -1).sum(axis=1)
        assert np.array_equal(chk_sub, chk[-5:])
        assert np.array_equal(sub.energy(), e_run[-5:])


def test_config3_box_of_2_to_the_20():
    N = 1 << 20
    pos, sp, box = ka_lattice(N, 1.2, seed=0)
    par = M.flatten_model_matrix(M.KobAndersen())
    finals = []
    for _ in range(2):
        with DeviceContext(1, N, 3, 2, M.MODEL_LJ, mode=L.MODE_BOX) as ctx:
            ctx.set_model(par)
            ctx.upload(pos, sp, box, 1.0)
            ctx.init_energy()
            ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
            ctx.seed(42)
            ctx.run(2 * N)
            e_run, e_tot = ctx.energy()[0], ctx.total_energy()[0]
            assert abs(e_run - e_tot) / abs(e_tot) < 1e-11
            calls, acc = ctx.counters()
            assert calls[0, 0] == 2 * N and 0.1 < acc[0, 0] / calls[0, 0] < 0.9
            p, s = ctx.download()
            assert np.array_equal(np.bincount(s[0]), np.bincount(sp))
            finals.append(p[0])
            e_final = e_run
    assert np.array_equal(finals[0], finals[1])
    # the oracle (reference arithmetic: LinkedList cells, division minimum image) on the final configuration of the
    # full-size box: total energy within 1e-12 of the device's running energy and of its recomputation
    orc = O.OracleSystem(finals[0] - np.floor(finals[0] / box) * box, sp, box, 1.0, M.MODEL_LJ, par, O.LINKEDLIST)
    assert abs(orc.energy - e_final) / abs(orc.energy) < 1e-12, (orc.energy, e_final)


def test_config2_block_of_chains_replayed_by_the_oracle():
    """32 consecutive chains picked at random out of the 4096 of BASELINE config 2: (a) the production run (persistent
    work queue, no tracing) and a traced run of just those chains (same global chain ids) end in bit-identical
    configurations; (b) the oracle replays the traced proposals of every one of them: same decisions, same energies."""
    N, nch, n_trials, nsub = 1000, 4096, 1500, 32
    pos, sp, box = ka_lattice(N, 1.2, seed=0)
    par = M.flatten_model_matrix(M.KobAndersen())
    first = int(np.random.default_rng(20261017).integers(0, nch - nsub))
    with DeviceContext(nch, N, 3, 2, M.MODEL_LJ) as ctx:
        ka_chains(ctx, pos, sp, box, nch)
        ctx.run(n_trials)
        p_full, _ = ctx.download(first, nsub)
        e_full = ctx.energy()[first:first + nsub]
    with DeviceContext(nsub, N, 3, 2, M.MODEL_LJ, chain_offset=first) as sub:
        ka_chains(sub, pos, sp, box, nsub)
        tr, acc, dE = sub.run_traced(n_trials)
        p_sub, _ = sub.download()
        assert np.array_equal(p_sub, p_full)
        assert np.array_equal(sub.energy(), e_full)
    z = np.zeros(n_trials, dtype=np.int32)
    for c in range(nsub):
        orc = O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, par, O.LINKEDLIST)
        o_acc, o_dE, _ = orc.replay(tr["kind"][c], tr["i"][c], z, z, z, tr["delta"][c], tr["u"][c], 1)
        assert np.array_equal(o_acc, acc[c]), f"chain {first + c}"
        assert np.max(np.abs(o_dE - dE[c])) < 1e-10
        assert abs(orc.energy - e_full[c]) < 1e-10 * abs(orc.energy)


def _replay_block(tr, acc, dE, e_full, make_oracle, pool_labels, first):
    for c in range(tr.shape[0]):
        orc = make_oracle()
        t = tr[c]
        spA = np.array([pool_labels[m][0] for m in t["move"]], dtype=np.int32)
        spB = np.array([pool_labels[m][1] for m in t["move"]], dtype=np.int32)
        o_acc, o_dE, _ = orc.replay(t["kind"], t["i"], np.maximum(t["j"], 0), spA, spB, t["delta"], t["u"], 1)
        assert np.array_equal(o_acc, acc[c]), f"chain {first + c}"
        fin = np.isfinite(o_dE)
        assert np.max(np.abs(o_dE[fin] - dE[c][fin]) / np.maximum(1.0, np.abs(o_dE[fin]))) < 1e-10
        assert abs(orc.energy - e_full[c]) < 1e-10 * abs(orc.energy)


def test_config4_swap_pool_block_of_chains_replayed_by_the_oracle(config0):
    """BASELINE config 4 at the size bench/other_configs.py times: 1184 chains of the reference's 2-D ternary fixture
    (N = 1290, JBB) under Displacement 0.8 + DiscreteSwap (1,3) 0.1 + (2,3) 0.1 (test/gerhard_energy_distribution.jl:
    63-72).  The production run (work queue, no tracing) and a traced run of 8 of its chains alone end bit-identically;
    the oracle replays those chains; composition and energy bookkeeping hold for all 1184."""
    nch, n_trials, nsub = 1184, 3000, 8
    par = M.flatten_model_matrix(M.JBB())
    pool = [dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="swap", prob=0.1, species=(1, 3)),
            dict(kind="swap", prob=0.1, species=(2, 3))]
    pos, sp, box, T = config0["position"], config0["species"], config0["box"], 0.231
    first = int(np.random.default_rng(4).integers(0, nch - nsub))

    def setup(ctx, n):
        ctx.set_model(par)
        ctx.upload(np.stack([pos] * n), np.stack([sp] * n), box, T)
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(42)

    with DeviceContext(nch, 1290, 2, 3, M.MODEL_SMOOTHLJ) as ctx:
        setup(ctx, nch)
        ctx.run(n_trials)
        p_full, s_full = ctx.download(first, nsub)
        e_run, e_tot = ctx.energy(), ctx.total_energy()
        assert np.max(np.abs(e_run - e_tot) / np.abs(e_tot)) < 1e-10
        calls, accepted = ctx.counters()
        assert np.all(calls.sum(axis=1) == n_trials) and accepted[:, 1:].sum() > 0  # some swaps were accepted
        _, s_all = ctx.download()
        assert all(np.array_equal(np.bincount(s, minlength=4), np.bincount(sp, minlength=4)) for s in s_all[::37])
    with DeviceContext(nsub, 1290, 2, 3, M.MODEL_SMOOTHLJ, chain_offset=first) as sub:
        setup(sub, nsub)
        tr, acc, dE = sub.run_traced(n_trials)
        p_sub, s_sub = sub.download()
        assert np.array_equal(p_sub, p_full) and np.array_equal(s_sub, s_full)
        assert np.array_equal(sub.energy(), e_run[first:first + nsub])
    _replay_block(tr, acc, dE, e_run[first:first + nsub],
                  lambda: O.OracleSystem(pos, sp, box, T, M.MODEL_SMOOTHLJ, par, O.LINKEDLIST), {0: (0, 0), 1: (1, 3), 2: (2, 3)}, first)


def test_config5_flip_pool_block_of_chains_replayed_by_the_oracle(molecule):
    """BASELINE config 5 at the size bench/other_configs.py times: 296 chains of the 1000-trimer fixture (N = 3000,
    Trimer / GeneralKG) under the pool examples/ortho-terphenyl runs, Displacement 0.8 + MoleculeFlip 0.2
    (params-template.toml:59-68).  Production run vs a traced run of 4 of its chains (bit-identical), oracle replay of
    those, per-molecule composition and energy bookkeeping for all."""
    nch, n_trials, nsub, n = 296, 2500, 4, 3000
    par = M.flatten_model_matrix(M.Trimer())
    pool = [dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="flip", prob=0.2)]
    pos, sp, box, T = molecule["position"], molecule["species"], molecule["box"], molecule["temperature"]
    bonds = [[j - 1 for j in b] for b in molecule["bonds"]]
    first = int(np.random.default_rng(5).integers(0, nch - nsub))

    def setup(ctx, k):
        ctx.set_model(par)
        ctx.set_bonds(bonds)
        ctx.set_molecules(np.arange(0, n, 3), np.full(n // 3, 3))
        ctx.upload(np.stack([pos] * k), np.stack([sp] * k), box, T)
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(42)

    with DeviceContext(nch, n, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        setup(ctx, nch)
        ctx.run(n_trials)
        p_full, s_full = ctx.download(first, nsub)
        e_run, e_tot = ctx.energy(), ctx.total_energy()
        assert np.max(np.abs(e_run - e_tot) / np.abs(e_tot)) < 1e-10
        calls, accepted = ctx.counters()
        assert np.all(calls.sum(axis=1) == n_trials) and np.all(accepted[:, 1] > 0)  # every chain accepted flips
        _, s_all = ctx.download()
        # a flip exchanges species inside one molecule: every molecule keeps its multiset of species
        assert np.array_equal(np.sort(s_all.reshape(nch, n // 3, 3), axis=2), np.sort(np.broadcast_to(sp, (nch, n)).reshape(nch, n // 3, 3), axis=2))
    with DeviceContext(nsub, n, 3, 3, M.MODEL_KG, molecules=True, chain_offset=first) as sub:
        setup(sub, nsub)
        tr, acc, dE = sub.run_traced(n_trials)
        p_sub, s_sub = sub.download()
        assert np.array_equal(p_sub, p_full) and np.array_equal(s_sub, s_full)
        assert np.array_equal(sub.energy(), e_run[first:first + nsub])
    wrapped = pos - np.floor(pos / box) * box
    _replay_block(tr, acc, dE, e_run[first:first + nsub],
                  lambda: O.OracleSystem(wrapped, sp, box, T, M.MODEL_KG, par, O.LINKEDLIST, bonds=bonds), {0: (0, 0), 1: (0, 0)}, first)
