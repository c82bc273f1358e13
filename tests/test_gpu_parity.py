"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (BASELINE.json north_star): total and local energies within 1e-12 relative of the fp64
reference arithmetic; accept/reject decisions bit-exact when the same proposals and uniforms are replayed.
"""
import numpy as np
import pytest

from oracle import oracle as O
import particlesmc_b200 as P
from particlesmc_b200 import _lib as L
from particlesmc_b200 import models as M
from particlesmc_b200.device import TRIAL_DTYPE, DeviceContext
from particlesmc_b200.synthetic import ka_lattice, lattice

pytestmark = pytest.mark.gpu

RTOL_E = 1e-12


def zero_based(bonds):
    return [[j - 1 for j in b] for b in bonds]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def assert_local(e_gpu, e_ref):
    scale = np.maximum(np.abs(e_ref), 1.0)
    assert np.max(np.abs(e_gpu - e_ref) / scale) < RTOL_E


def ka_config(N=1000, seed=0, jitter=0.05):
    pos, sp, box = ka_lattice(N, 1.2, seed)
    rng = np.random.default_rng(seed + 100)
    pos = pos + rng.normal(0.0, jitter, pos.shape)
    return pos - np.floor(pos / box) * box, sp, box


# ---------------------------------------------------------------------------------------------------
# energies: the reference's known answers + oracle parity
# ---------------------------------------------------------------------------------------------------
def test_config0_energy(config0):
    """test/runtests.jl:22-38 through the device path."""
    s = P.System(config0["position"], config0["species"], config0["density"], config0["temperature"], P.JBB(),
                 list_type=P.LinkedList)
    assert s.N == 1290 and s.d == 2
    assert abs(P.energy(s) - config0["ref"]) < 1e-6
    orc = O.OracleSystem(config0["position"], config0["species"], config0["box"], config0["temperature"],
                         M.MODEL_SMOOTHLJ, M.flatten_model_matrix(M.JBB()), O.LINKEDLIST)
    assert rel(s.energy[0], orc.energy) < RTOL_E
    assert_local(P.compute_energy_particle(s), orc.local_energies())
    assert abs(P.compute_energy_particle(s, 7) - orc.local_energy(6)) < 1e-12 * max(1, abs(orc.local_energy(6)))


def test_molecule_energy(molecule):
    """test/runtests.jl:136-149 through the device path."""
    s = P.System(molecule["position"], molecule["species"], molecule["molecule"], molecule["density"],
                 molecule["temperature"], P.Trimer(), molecule["bonds"], list_type=P.LinkedList)
    assert s.N == 3000 and s.Nmol == 1000
    assert abs(P.energy(s) - molecule["ref"]) < 1e-6
    orc = O.OracleSystem(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"],
                         M.MODEL_KG, M.flatten_model_matrix(M.Trimer()), O.LINKEDLIST,
                         bonds=zero_based(molecule["bonds"]))
    assert rel(s.energy[0], orc.energy) < RTOL_E
    assert_local(P.compute_energy_particle(s), orc.local_energies())


@pytest.mark.parametrize("name,d", [("KobAndersen", 3), ("BHHP", 3), ("BHHP", 2), ("JBB", 2)])
def test_model_energies_vs_oracle(name, d):
    mm = M.NAMED_MODELS[name]()
    ns = len(mm)
    N = 512 if d == 3 else 400
    fr = [1.0 / ns] * ns
    pos, sp, box = lattice(N, d, 1.0 if d == 3 else 0.9, seed=4, fractions=fr)
    pos = pos + np.random.default_rng(5).normal(0, 0.04, pos.shape)
    s = P.System(pos, sp, N / np.prod(box), 1.0, mm)
    orc = O.OracleSystem(pos - np.floor(pos / s.box) * s.box, sp, s.box, 1.0, M.model_kind(mm),
                         M.flatten_model_matrix(mm), O.EMPTYLIST)
    assert rel(s.energy[0], orc.energy) < RTOL_E
    assert_local(P.compute_energy_particle(s), orc.local_energies())


def test_unwrapped_positions_roundtrip():
    """Positions outside the box are legal input (src/moves.jl:46-48 never re-wraps) and come back unwrapped."""
    pos, sp, box = ka_config(216, 2)
    shift = np.random.default_rng(0).integers(-3, 4, pos.shape) * box
    par = M.flatten_model_matrix(M.KobAndersen())
    with DeviceContext(2, 216, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([pos, pos + shift]), np.stack([sp, sp]), box, 1.0)
        ctx.init_energy()
        e = ctx.energy()
        assert rel(e[1], e[0]) < 1e-12
        back, spb = ctx.download()
        assert np.allclose(back[1], pos + shift, rtol=0, atol=1e-12)
        assert np.array_equal(spb[0], sp)


def test_initial_overlap_is_an_error():
    """atoms.jl:53-55: 'Initial configuration has infinite or NaN energy.'"""
    pos, sp, box = ka_config(216, 1)
    pos[5] = pos[9]
    with pytest.raises(ValueError, match="infinite or NaN energy"):
        P.System(pos, sp, 1.2, 1.0, P.KobAndersen())


def test_invalid_arguments_are_errors():
    with pytest.raises(P.PMCError):
        DeviceContext(1, 100, 4, 1, M.MODEL_LJ)  # dim 4
    with DeviceContext(1, 64, 3, 2, M.MODEL_LJ) as ctx:
        pos, sp, box = ka_config(64, 1)
        with pytest.raises(P.PMCError, match="pmc_set_model"):
            ctx.run(10)
        ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
        bad = sp.copy()
        bad[0] = 3
        with pytest.raises(P.PMCError, match="species labels"):
            ctx.upload(pos, bad, box, 1.0)
        ctx.upload(pos, sp, box, 1.0)
        with pytest.raises(P.PMCError, match="pmc_set_moves"):
            ctx.run(10)


# ---------------------------------------------------------------------------------------------------
# decision parity: the production sweep kernel's own proposals replayed by the oracle
# ---------------------------------------------------------------------------------------------------
def check_trace(ctx, orcs, pool_labels, n_trials, atol_x=1e-11):
    tr, acc, dE = ctx.run_traced(n_trials)
    e_run = ctx.energy()
    pos, sp = ctx.download()
    for c, orc in enumerate(orcs):
        t = tr[c]
        spA = np.array([pool_labels[m][0] for m in t["move"]], dtype=np.int32)
        spB = np.array([pool_labels[m][1] for m in t["move"]], dtype=np.int32)
        o_acc, o_dE, o_E = orc.replay(t["kind"], t["i"], np.maximum(t["j"], 0), spA, spB, t["delta"], t["u"], 1)
        assert np.array_equal(o_acc, acc[c]), f"chain {c}: {np.count_nonzero(o_acc != acc[c])} decisions differ"
        fin = np.isfinite(o_dE)
        assert np.max(np.abs(o_dE[fin] - dE[c][fin]) / np.maximum(1.0, np.abs(o_dE[fin]))) < 1e-11
        assert rel(e_run[c], orc.energy) < 1e-11
        opos, osp = orc.state()
        assert np.array_equal(osp, sp[c])
        assert np.max(np.abs(opos - pos[c])) < atol_x
        assert 0 < acc[c].sum() < n_trials
    return tr, acc


def test_trace_parity_kob_andersen():
    """BASELINE config 1/2 shape: KA N=1000, rho=1.2, Displacement sigma=0.05."""
    M_ = 3
    par = M.flatten_model_matrix(M.KobAndersen())
    cfgs = [ka_config(1000, s) for s in range(M_)]
    with DeviceContext(M_, 1000, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), cfgs[0][2], [1.0, 0.5, 2.0])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(42)
        orcs = [O.OracleSystem(c[0] - np.floor(c[0] / c[2]) * c[2], c[1], c[2], T, M.MODEL_LJ, par, O.LINKEDLIST)
                for c, T in zip(cfgs, [1.0, 0.5, 2.0])]
        for c, o in enumerate(orcs):
            assert rel(ctx.energy()[c], o.energy) < RTOL_E
        tr, acc = check_trace(ctx, orcs, {0: (0, 0)}, 3000)
        # the proposal stream itself: i uniform over particles, delta ~ N(0, sigma^2)
        assert tr["i"].min() >= 0 and tr["i"].max() < 1000
        assert abs(tr["delta"].std() - 0.05) < 0.002 and abs(tr["delta"].mean()) < 0.002
        assert 0.0 <= tr["u"].min() and tr["u"].max() < 1.0
        calls, accepted = ctx.counters()
        assert calls[:, 0].tolist() == [3000] * M_ and accepted[:, 0].tolist() == acc.sum(axis=1).tolist()


@pytest.mark.parametrize("name,d,prefilter", [("BHHP", 3, 0), ("BHHP", 2, 0), ("JBB", 2, 0), ("JBB", 2, 1), ("BHHP", 3, -1)])
def test_trace_parity_models_displacement(name, d, prefilter):
    """Every potential family and both dimensions through the sweep kernels the Displacement pool selects
    (prefilter 0: speculative, 1: one trial at a time, -1: every candidate in fp64): decisions bit-exact against the
    oracle replaying the same proposals, dE and final state to rounding."""
    mm = M.NAMED_MODELS[name]()
    ns = len(mm)
    N = 512 if d == 3 else 400
    par = M.flatten_model_matrix(mm)
    pos, sp, box = lattice(N, d, 1.0 if d == 3 else 0.9, seed=7, fractions=[1.0 / ns] * ns)
    cfgs = []
    for k in range(2):
        p = pos + np.random.default_rng(10 + k).normal(0, 0.04, pos.shape)
        cfgs.append(p - np.floor(p / box) * box)
    with DeviceContext(2, N, d, ns, M.model_kind(mm), prefilter=prefilter) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack(cfgs), np.stack([sp, sp]), box, [1.0, 0.4])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.06)])
        ctx.seed(21)
        orcs = [O.OracleSystem(c, sp, box, T, M.model_kind(mm), par, O.LINKEDLIST) for c, T in zip(cfgs, (1.0, 0.4))]
        check_trace(ctx, orcs, {0: (0, 0)}, 2500)


def test_trace_parity_swaps(config0):
    """test/runtests.jl:93-129 pool on the reference's own ternary configuration."""
    par = M.flatten_model_matrix(M.JBB())
    pool = [dict(kind="displacement", prob=0.2, sigma=0.05), dict(kind="swap", prob=0.4, species=(1, 3)),
            dict(kind="swap", prob=0.4, species=(2, 3))]
    with DeviceContext(2, 1290, 2, 3, M.MODEL_SMOOTHLJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([config0["position"]] * 2), np.stack([config0["species"]] * 2), config0["box"],
                   [config0["temperature"], 1.0])
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(10)
        orcs = [O.OracleSystem(config0["position"], config0["species"], config0["box"], T, M.MODEL_SMOOTHLJ, par,
                               O.LINKEDLIST) for T in (config0["temperature"], 1.0)]
        tr, acc = check_trace(ctx, orcs, {0: (0, 0), 1: (1, 3), 2: (2, 3)}, 2500)
        kinds = np.bincount(tr["move"].ravel(), minlength=3) / tr.size
        assert np.all(np.abs(kinds - [0.2, 0.4, 0.4]) < 0.03)
        sw = tr["kind"] == 1
        assert acc[sw].sum() > 0  # some swaps are accepted at these temperatures
        _, sp = ctx.download()
        assert np.bincount(sp[0])[1:].tolist() == [600, 330, 360]


def test_trace_parity_molecules(molecule):
    """BASELINE config 5 shape: 1000 trimers, bonded FENE + non-bonded WCA, Displacement."""
    par = M.flatten_model_matrix(M.Trimer())
    with DeviceContext(1, 3000, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(molecule["bonds"]))
        ctx.upload(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(10)
        orc = O.OracleSystem(molecule["position"], molecule["species"], molecule["box"], molecule["temperature"],
                             M.MODEL_KG, par, O.LINKEDLIST, bonds=zero_based(molecule["bonds"]))
        check_trace(ctx, [orc], {0: (0, 0)}, 2000)


def test_trace_parity_small_molecules(molecule):
    """The first 300 trimers of the fixture in a smaller box (N = 900 <= 1024: the four-warp speculative kernel
    with the bond pass) -- decisions bit-exact against the oracle, as for the full fixture."""
    par = M.flatten_model_matrix(M.Trimer())
    n = 900
    pos = molecule["position"][:n].copy()
    sp = molecule["species"][:n]
    bonds = [[j for j in b if j <= n] for b in molecule["bonds"][:n]]
    box = molecule["box"]  # same box: a dilute system, every bond intact
    with DeviceContext(2, n, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(bonds))
        ctx.upload(np.stack([pos, pos]), np.stack([sp, sp]), box, [2.0, 0.7])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.06)])
        ctx.seed(77)
        orcs = [O.OracleSystem(pos - np.floor(pos / box) * box, sp, box, T, M.MODEL_KG, par, O.LINKEDLIST,
                               bonds=zero_based(bonds)) for T in (2.0, 0.7)]
        for c in range(2):
            assert rel(ctx.energy()[c], orcs[c].energy) < RTOL_E
        check_trace(ctx, orcs, {0: (0, 0)}, 3000)


def test_trace_parity_2d_displacement_two_mask_words(config0):
    """test/config_0 (2-D ternary JBB, N = 1290 -> 2048 padded candidates: two survivor-mask words per lane) with a
    Displacement-only pool: the speculative kernel against the oracle."""
    par = M.flatten_model_matrix(M.JBB())
    with DeviceContext(2, 1290, 2, 3, M.MODEL_SMOOTHLJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([config0["position"]] * 2), np.stack([config0["species"]] * 2), config0["box"],
                   [config0["temperature"], 1.0])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.08)])
        ctx.seed(5)
        orcs = [O.OracleSystem(config0["position"], config0["species"], config0["box"], T, M.MODEL_SMOOTHLJ, par,
                               O.LINKEDLIST) for T in (config0["temperature"], 1.0)]
        check_trace(ctx, orcs, {0: (0, 0)}, 3000)


def test_trace_parity_molecule_flip(molecule):
    """examples/ortho-terphenyl pool: Displacement 0.8 + MoleculeFlip 0.2 (params-template.toml:59-68)."""
    par = M.flatten_model_matrix(M.Trimer())
    with DeviceContext(2, 3000, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(molecule["bonds"]))
        ctx.set_molecules(np.arange(0, 3000, 3), np.full(1000, 3))
        ctx.upload(np.stack([molecule["position"]] * 2), np.stack([molecule["species"]] * 2), molecule["box"],
                   [molecule["temperature"], 8.0])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="flip", prob=0.2)])
        ctx.seed(3)
        orcs = [O.OracleSystem(molecule["position"], molecule["species"], molecule["box"], T, M.MODEL_KG, par,
                               O.LINKEDLIST, bonds=zero_based(molecule["bonds"])) for T in (molecule["temperature"], 8.0)]
        tr, acc = check_trace(ctx, orcs, {0: (0, 0), 1: (0, 0)}, 2500)
        fl = tr["kind"] == 2
        assert abs(fl.mean() - 0.2) < 0.03
        # both sites belong to the same trimer and carried different species when the flip was drawn
        assert np.all(tr["i"][fl] // 3 == tr["j"][fl] // 3) and np.all(tr["i"][fl] != tr["j"][fl])
        assert acc[fl].sum() > 0
        _, sp = ctx.download()
        for c in range(2):  # every trimer still holds one site of each species
            assert np.array_equal(np.sort(sp[c].reshape(1000, 3), axis=1), np.tile([1, 2, 3], (1000, 1)))


def test_replay_of_injected_proposals():
    """Proposals and uniforms recorded elsewhere (here: numpy) replayed through the sweep kernel with the
    reference's acceptance arithmetic: decisions bit-exact vs the oracle running the reference's revert."""
    rng = np.random.default_rng(7)
    pos, sp, box = ka_config(512, 3)
    par = M.flatten_model_matrix(M.KobAndersen())
    n = 4000
    tr = np.zeros((1, n), dtype=TRIAL_DTYPE)
    tr["kind"] = 0
    tr["i"] = rng.integers(0, 512, n)
    tr["j"] = -1
    tr["delta"] = rng.normal(0, 0.08, (1, n, 3))
    tr["u"] = rng.random(n)
    is_swap = rng.random(n) < 0.2
    A_ids, B_ids = np.nonzero(sp == 1)[0], np.nonzero(sp == 2)[0]
    orc = O.OracleSystem(pos - np.floor(pos / box) * box, sp, box, 0.8, M.MODEL_LJ, par, O.LINKEDLIST)
    with DeviceContext(1, 512, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(pos, sp, box, 0.8)
        ctx.init_energy()
        # swaps must name a current (A, B) pair: generate them against the evolving oracle state in chunks
        done = 0
        while done < n:
            m = min(500, n - done)
            cur = orc.state()[1]
            A_ids, B_ids = np.nonzero(cur == 1)[0], np.nonzero(cur == 2)[0]
            blk = tr[0, done:done + m]
            first_swap = True
            for q in range(m):
                if is_swap[done + q] and first_swap:  # one swap per chunk keeps the id lists valid
                    blk["kind"][q], blk["move"][q] = 1, 1
                    blk["i"][q], blk["j"][q] = rng.choice(A_ids), rng.choice(B_ids)
                    blk["delta"][q] = 0.0
                    first_swap = False
            spA = np.where(blk["kind"] == 1, 1, 0).astype(np.int32)
            spB = np.where(blk["kind"] == 1, 2, 0).astype(np.int32)
            o_acc, o_dE, _ = orc.replay(blk["kind"], blk["i"], np.maximum(blk["j"], 0), spA, spB, blk["delta"],
                                        blk["u"], 0)
            g_acc, g_dE = ctx.replay(blk.reshape(1, m))
            assert np.array_equal(o_acc, g_acc[0])
            fin = np.isfinite(o_dE)
            assert np.max(np.abs(o_dE[fin] - g_dE[0][fin]) / np.maximum(1.0, np.abs(o_dE[fin]))) < 1e-11
            done += m
        gpos, gsp = ctx.download()
        opos, osp = orc.state()
        assert np.array_equal(gsp[0], osp)
        assert np.max(np.abs(gpos[0] - opos)) < 1e-11
        assert rel(ctx.energy()[0], orc.energy) < 1e-10
        assert rel(ctx.total_energy()[0], orc.total_energy()) < RTOL_E


# ---------------------------------------------------------------------------------------------------
# stream properties: launch splitting and sharding do not change the chains
# ---------------------------------------------------------------------------------------------------
def _ka_ctx(n_chains, chain_offset=0, first=0, threads=0, prefilter=0):
    par = M.flatten_model_matrix(M.KobAndersen())
    cfgs = [ka_config(216, s) for s in range(first, first + n_chains)]
    ctx = DeviceContext(n_chains, 216, 3, 2, M.MODEL_LJ, chain_offset=chain_offset, threads=threads,
                        prefilter=prefilter)
    ctx.set_model(par)
    ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), cfgs[0][2], 1.0)
    ctx.init_energy()
    ctx.set_moves([dict(kind="displacement", prob=0.8, sigma=0.07), dict(kind="swap", prob=0.2, species=(1, 2))])
    ctx.seed(99)
    return ctx


def test_split_launches_equal_one_launch():
    with _ka_ctx(3) as a, _ka_ctx(3) as b:
        a.run(1500)
        for n in (700, 1, 799):
            b.run(n)
        pa, sa = a.download()
        pb, sb = b.download()
        assert np.array_equal(pa, pb) and np.array_equal(sa, sb)
        assert np.array_equal(a.energy(), b.energy())
        assert np.array_equal(a.counters()[1], b.counters()[1])


def test_sharded_chains_equal_unsharded():
    """Chains keyed by GLOBAL index: ranks holding [0,2) and [2,4) reproduce a single 4-chain context."""
    with _ka_ctx(4) as full, _ka_ctx(2, chain_offset=0, first=0) as r0, _ka_ctx(2, chain_offset=2, first=2) as r1:
        for c in (full, r0, r1):
            c.run(1200)
        pf, sf = full.download()
        p0, s0 = r0.download()
        p1, s1 = r1.download()
        assert np.array_equal(pf[:2], p0) and np.array_equal(pf[2:], p1)
        assert np.array_equal(sf[:2], s0) and np.array_equal(sf[2:], s1)


@pytest.mark.parametrize("threads,prefilter", [(32, 0), (64, 0), (128, 0), (128, 1), (128, -1), (256, -1)])
def test_cta_size_and_prefilter_do_not_change_decisions(threads, prefilter):
    """The fixed-point prefilter only selects WHICH candidates get the fp64 evaluation (a superset of the pairs
    inside the cutoff); decisions must equal the kernel that visits every candidate in fp64.  The default context
    (128 threads, prefilter 0) runs the speculative kernel (four trials of a chain per round, chains_spec.cuh);
    prefilter 1 the one-trial-at-a-time kernel; other CTA sizes and prefilter -1 the general kernel."""
    with _ka_ctx(2) as a, _ka_ctx(2, threads=threads, prefilter=prefilter) as b:
        _, acc_a, _ = a.run_traced(800)
        _, acc_b, _ = b.run_traced(800)
        assert np.array_equal(acc_a, acc_b)
        assert np.max(np.abs(a.download()[0] - b.download()[0])) < 1e-12


def _ka_disp_ctx(n_chains, N, prefilter, T=1.0):
    par = M.flatten_model_matrix(M.KobAndersen())
    cfgs = [ka_config(N, s) for s in range(n_chains)]
    ctx = DeviceContext(n_chains, N, 3, 2, M.MODEL_LJ, prefilter=prefilter)
    ctx.set_model(par)
    ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), cfgs[0][2], T)
    ctx.init_energy()
    ctx.set_moves([dict(kind="displacement", prob=0.7, sigma=0.05), dict(kind="displacement", prob=0.3, sigma=0.12)])
    ctx.seed(7)
    return ctx


@pytest.mark.parametrize("N,sweeps", [(1000, 12), (512, 12), (216, 30)])
def test_speculative_rounds_reproduce_the_sequential_chain(N, sweeps):
    """Speculation must not change the Markov chain (chains_spec.cuh): every decision and every final coordinate
    of 6 chains equal those of the kernel that evaluates one trial at a time (prefilter = 1).  Trials whose
    neighbourhood was touched by an earlier accepted trial of the same round are re-evaluated, so nothing but the
    schedule differs; dE agrees to rounding (same pairs, different summation order)."""
    with _ka_disp_ctx(6, N, 0) as a, _ka_disp_ctx(6, N, 1) as b:
        n = sweeps * N + 37  # not a multiple of the batch or of the round size
        _, acc_a, dE_a = a.run_traced(n)
        _, acc_b, dE_b = b.run_traced(n)
        assert np.array_equal(acc_a, acc_b)
        assert 0.2 < acc_a.mean() < 0.9
        assert np.array_equal(a.download()[0], b.download()[0])
        assert np.max(np.abs(dE_a - dE_b) / np.maximum(1.0, np.abs(dE_b))) < 1e-11
        ca, aa = a.counters()
        cb, ab = b.counters()
        assert np.array_equal(ca, cb) and np.array_equal(aa, ab) and ca.sum() == 6 * n
        assert np.allclose(a.energy(), b.energy(), rtol=1e-12, atol=0)
        assert np.allclose(a.energy(), a.total_energy(), rtol=1e-11, atol=0)
        # a second launch continues the same stream (trial counter, image counters, register copies reloaded)
        a.run(3 * N)
        b.run(3 * N)
        assert np.array_equal(a.download()[0], b.download()[0])


def test_noncubic_box_uses_direct_kernel_and_matches_oracle():
    """Per-axis box lengths are legal at the ABI; the prefilter needs a cubic box, so this exercises the
    direct fp64 kernel against the oracle."""
    rng = np.random.default_rng(3)
    N, box = 300, np.array([6.0, 7.5, 5.5])
    grid = np.stack(np.meshgrid(np.arange(6), np.arange(10), np.arange(5), indexing="ij"), -1).reshape(-1, 3)
    pos = (grid[:N] + 0.5) * (box / np.array([6, 10, 5])) + rng.normal(0, 0.03, (N, 3))
    pos = pos - np.floor(pos / box) * box
    sp = rng.permutation(np.repeat([1, 2], [240, 60])).astype(np.int64)
    par = M.flatten_model_matrix(M.KobAndersen())
    orc = O.OracleSystem(pos, sp, box, 1.0, M.MODEL_LJ, par, O.LINKEDLIST)
    with DeviceContext(1, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(pos, sp, box, 1.0)
        ctx.init_energy()
        assert rel(ctx.energy()[0], orc.energy) < RTOL_E
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(1)
        check_trace(ctx, [orc], {0: (0, 0)}, 1500)


def test_energy_bookkeeping_after_long_run():
    with _ka_ctx(4) as ctx:
        ctx.run(50 * 216)
        e_run, e_tot = ctx.energy(), ctx.total_energy()
        assert np.max(np.abs(e_run - e_tot) / np.abs(e_tot)) < 1e-11
        calls, acc = ctx.counters()
        assert np.all(calls.sum(axis=1) == 50 * 216)
        rate = acc[:, 0] / calls[:, 0]
        assert np.all((rate > 0.05) & (rate < 0.95))


# ---------------------------------------------------------------------------------------------------
# round 2: the holes the round-1 review named
# ---------------------------------------------------------------------------------------------------
def test_swap_slots_resolve_through_the_species_lists_like_the_reference(config0):
    """DoubleUniform draws SLOTS of the per-species id lists (src/moves.jl:238-241); which particle a slot names depends on
    how update_species_list! (src/moves.jl:175-179) rearranged the lists after every accepted swap.  Here the oracle
    draws the particles itself: it receives the slots the GPU's Philox stream produced (recomputed on the host from
    seed, chain and trial number) and resolves them through its OWN lists; the GPU must have picked the same particles
    and taken the same decisions for thousands of trials with hundreds of accepted swaps in between."""
    par = M.flatten_model_matrix(M.JBB())
    pool = [dict(kind="displacement", prob=0.2, sigma=0.05), dict(kind="swap", prob=0.4, species=(1, 3)),
            dict(kind="swap", prob=0.4, species=(2, 3))]
    species_of = {1: (1, 3), 2: (2, 3)}
    seed, n_trials, T = 1234, 4000, 1.0
    counts = np.bincount(config0["species"], minlength=4)
    with DeviceContext(2, 1290, 2, 3, M.MODEL_SMOOTHLJ, chain_offset=5) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack([config0["position"]] * 2), np.stack([config0["species"]] * 2), config0["box"], T)
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(seed)
        tr, acc, dE = ctx.run_traced(n_trials)
        gpos, gsp = ctx.download()
    key = (seed & 0xFFFFFFFF, seed >> 32)
    for c in range(2):
        orc = O.OracleSystem(config0["position"], config0["species"], config0["box"], T, M.MODEL_SMOOTHLJ, par, O.LINKEDLIST)
        n_acc_swaps = 0
        for t in range(n_trials):
            r = tr[c, t]
            if r["kind"] == 0:
                a, _, _ = orc.step_displacement(r["i"], r["delta"][:2], r["u"], 0)
            else:
                A, B = species_of[int(r["move"])]
                blkA = O.philox4x32_10((t, 0, 5 + c, 0), key)
                blkB = O.philox4x32_10((t, 0, 5 + c, 1), key)
                ka = (int(blkA[1]) * int(counts[A])) >> 32
                kb = (int(blkB[0]) * int(counts[B])) >> 32
                a, i, j, _, _ = orc.step_swap_draw(A, B, ka, kb, r["u"], 0)
                assert (i, j) == (int(r["i"]), int(r["j"])), f"chain {c} trial {t}: slots ({ka}, {kb}) name other particles"
                n_acc_swaps += int(a)
            assert a == bool(acc[c, t]), f"chain {c} trial {t}"
        assert n_acc_swaps > 100
        opos, osp = orc.state()
        assert np.array_equal(osp, gsp[c])
        d = opos - gpos[c]
        assert np.max(np.abs(d - np.round(d / config0["box"]) * config0["box"])) < 1e-11  # same point, any periodic image


def test_stretched_bonds_consecutive_trials_on_bonded_sites():
    """Bonded partners interact through FENE up to r0 (1.425 .. 1.575 for Trimer), beyond the WCA cutoff that sizes the
    filter sphere of a trial.  Trials on the two ends of a stretched bond land in the same speculative round; the second
    must not keep an energy computed from the partner's old position (src/molecules.jl:160-176).  Injected proposals,
    hot system (nearly every move accepted), decisions and energy changes against the oracle."""
    par = M.flatten_model_matrix(M.Trimer())
    nmol, spacing, blen = 64, 3.2, 1.33
    grid = np.array([(x, y, z) for x in range(4) for y in range(4) for z in range(4)], dtype=float) * spacing + 0.7
    pos = np.zeros((3 * nmol, 3))
    pos[0::3] = grid
    pos[1::3] = grid + [blen, 0.0, 0.0]
    pos[2::3] = grid + [blen, blen, 0.0]
    sp = np.tile([1, 2, 3], nmol)
    bonds = []
    for m in range(nmol):
        bonds += [[3 * m + 1], [3 * m, 3 * m + 2], [3 * m + 1]]
    box = np.full(3, 4 * spacing)
    n = 3 * nmol
    rng = np.random.default_rng(3)
    nt = 3000
    tr = np.zeros((1, nt), dtype=TRIAL_DTYPE)
    # runs of trials that walk along one molecule: 0-1-2-1-0 ... bonded sites back to back
    mol = rng.integers(0, nmol, nt // 5 + 1)
    walk = np.array([0, 1, 2, 1, 0])
    tr["i"][0] = (3 * np.repeat(mol, 5)[:nt] + np.tile(walk, nt // 5 + 1)[:nt])
    tr["j"] = -1
    tr["delta"][0] = rng.normal(0, 0.004, (nt, 3))
    tr["u"][0] = rng.random(nt)
    T = 50.0
    orc = O.OracleSystem(pos, sp, box, T, M.MODEL_KG, par, O.LINKEDLIST, bonds=bonds)
    assert np.isfinite(orc.energy)
    with DeviceContext(1, n, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(bonds)
        ctx.upload(pos, sp, box, T)
        ctx.init_energy()
        assert rel(ctx.energy()[0], orc.energy) < RTOL_E
        g_acc, g_dE = ctx.replay(tr)
        z = np.zeros(nt, dtype=np.int32)
        o_acc, o_dE, _ = orc.replay(tr["kind"][0], tr["i"][0], z, z, z, tr["delta"][0], tr["u"][0], 0)
        assert o_acc.mean() > 0.5  # hot: the earlier trial of a round usually moved the partner
        assert np.array_equal(o_acc, g_acc[0])
        fin = np.isfinite(o_dE)
        assert np.max(np.abs(o_dE[fin] - g_dE[0][fin]) / np.maximum(1.0, np.abs(o_dE[fin]))) < 1e-10
        assert rel(ctx.energy()[0], orc.energy) < 1e-10


def test_trace_parity_small_molecules_flip_and_swap(molecule):
    """The four-warp speculative kernel for Molecules with a pool that holds all three move kinds: Displacement,
    MoleculeFlip (src/moves.jl:291-352) and DiscreteSwap (src/moves.jl:137-214) on 300 trimers -- every decision against
    the oracle, species lists and per-molecule composition consistent afterwards."""
    par = M.flatten_model_matrix(M.Trimer())
    n = 900
    pos = molecule["position"][:n].copy()
    sp = molecule["species"][:n]
    bonds = [[j for j in b if j <= n] for b in molecule["bonds"][:n]]
    box = molecule["box"]
    pool = [dict(kind="displacement", prob=0.5, sigma=0.06), dict(kind="flip", prob=0.3), dict(kind="swap", prob=0.2, species=(1, 3))]
    with DeviceContext(2, n, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(bonds))
        ctx.set_molecules(np.arange(0, n, 3), np.full(n // 3, 3))
        ctx.upload(np.stack([pos, pos]), np.stack([sp, sp]), box, [2.0, 6.0])
        ctx.init_energy()
        ctx.set_moves(pool)
        ctx.seed(123)
        orcs = [O.OracleSystem(pos - np.floor(pos / box) * box, sp, box, T, M.MODEL_KG, par, O.LINKEDLIST,
                               bonds=zero_based(bonds)) for T in (2.0, 6.0)]
        tr, acc = check_trace(ctx, orcs, {0: (0, 0), 1: (0, 0), 2: (1, 3)}, 4000)
        # every kind was drawn; flips are accepted now and then, swaps between molecules (they re-label bonds whose rest
        # lengths differ by 10 %) practically never -- their energy changes are what check_trace compared
        assert (tr["kind"] == 1).sum() > 500 and (tr["kind"] == 2).sum() > 800 and acc[tr["kind"] == 2].sum() > 0
        fl = tr["kind"] == 2
        assert np.all(tr["i"][fl] // 3 == tr["j"][fl] // 3) and np.all(tr["i"][fl] != tr["j"][fl])
        _, spf = ctx.download()
        assert np.array_equal(np.bincount(spf[0], minlength=4), np.bincount(sp, minlength=4))  # composition conserved
        ctx.run(2000)  # the species lists written back by the traced launch feed the next one
        e_run, e_tot = ctx.energy(), ctx.total_energy()
        assert np.max(np.abs(e_run - e_tot) / np.abs(e_tot)) < 1e-10


def test_flip_only_pool_leaves_species_lists_usable_for_later_swaps(molecule):
    """A pool without DiscreteSwap does not keep the species lists up to date while it runs (the reference's Molecules
    carry none, src/molecules.jl:24-41); the kernel rebuilds them from the species when it leaves, ids ascending per
    species as at upload.  Phase 1: Displacement + MoleculeFlip against the oracle.  Phase 2: a pool WITH DiscreteSwap --
    the slots its Philox stream draws must name the particles that an oracle built from the downloaded phase-1 state
    (lists in index order) resolves them to."""
    par = M.flatten_model_matrix(M.Trimer())
    n = 900
    pos = molecule["position"][:n].copy()
    sp = molecule["species"][:n]
    bonds = [[j for j in b if j <= n] for b in molecule["bonds"][:n]]
    box = molecule["box"]
    seed, T, n1, n2 = 77, 4.0, 3000, 1500
    with DeviceContext(1, n, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(bonds))
        ctx.set_molecules(np.arange(0, n, 3), np.full(n // 3, 3))
        ctx.upload(pos[None], sp[None], box, T)
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=0.5, sigma=0.06), dict(kind="flip", prob=0.5)])
        ctx.seed(seed)
        orc = O.OracleSystem(pos - np.floor(pos / box) * box, sp, box, T, M.MODEL_KG, par, O.LINKEDLIST, bonds=zero_based(bonds))
        tr, acc = check_trace(ctx, [orc], {0: (0, 0), 1: (0, 0)}, n1)
        assert acc[0][tr[0]["kind"] == 2].sum() > 50  # accepted flips: the lists of phase 2 depend on them
        pos1, sp1 = ctx.download()
        assert not np.array_equal(sp1[0], sp)
        ctx.set_moves([dict(kind="displacement", prob=0.5, sigma=0.06), dict(kind="swap", prob=0.5, species=(1, 3))])
        tr2, acc2, _ = ctx.run_traced(n2)
    counts = np.bincount(sp1[0], minlength=4)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    p1 = pos1[0] - np.floor(pos1[0] / box) * box
    orc2 = O.OracleSystem(p1, sp1[0], box, T, M.MODEL_KG, par, O.LINKEDLIST, bonds=zero_based(bonds))
    n_swaps = 0
    for t in range(n2):
        r = tr2[0, t]
        if r["kind"] == 0:
            a, _, _ = orc2.step_displacement(r["i"], r["delta"], r["u"], 0)
        else:
            blkA = O.philox4x32_10((n1 + t, 0, 0, 0), key)
            blkB = O.philox4x32_10((n1 + t, 0, 0, 1), key)
            ka = (int(blkA[1]) * int(counts[1])) >> 32
            kb = (int(blkB[0]) * int(counts[3])) >> 32
            a, i, j, _, _ = orc2.step_swap_draw(1, 3, ka, kb, r["u"], 0)
            assert (i, j) == (int(r["i"]), int(r["j"])), f"trial {t}: slots ({ka}, {kb}) name other particles"
            n_swaps += 1
        assert a == bool(acc2[0, t]), f"trial {t}"
    assert n_swaps > 500


def test_work_counters_count_what_the_sweep_kernel_evaluates():
    """pmc_work_counters (bench.py's roofline.frac_actual): candidates that reached the fp64 pass and trial evaluations,
    counted on the device.  Evaluations >= trials (speculative rounds repeat a few), survivors per evaluation close to
    the density x filter-sphere volume, nothing counted while switched off."""
    N, n_chains, n_trials = 1000, 8, 4000
    pos, sp, box = ka_config(N, 5)
    with DeviceContext(n_chains, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
        ctx.upload(np.stack([pos] * n_chains), np.stack([sp] * n_chains), box, 1.0)
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(4)
        ctx.run(n_trials)
        assert ctx.work_counters(2) == (0, 0)
        ctx.work_counters(1)
        ctx.run(n_trials)
        surv, evals = ctx.work_counters(0)
        assert n_chains * n_trials <= evals < 1.25 * n_chains * n_trials
        assert 60 < surv / evals < 120  # ~ 1.2 x 4/3 pi 2.6^3 = 88 for the A particles, fewer for B
        ctx.run(n_trials)
        assert ctx.work_counters(2) == (surv, evals)  # off again
