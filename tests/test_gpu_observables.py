"""On-device pair-distance histograms (the counts behind g(r)) against numpy on the same configurations."""
import numpy as np
import pytest

from particlesmc_b200 import _lib as L
from particlesmc_b200 import models as M
from particlesmc_b200.device import DeviceContext
from particlesmc_b200.observables import radial_distribution
from particlesmc_b200.synthetic import ka_lattice

pytestmark = pytest.mark.gpu


def numpy_counts(pos, sp, box, sa, sb, rmax, nbins):
    Lb = box[0]
    h = np.zeros(nbins, dtype=np.int64)
    n = len(pos)
    iu, ju = np.triu_indices(n, 1)
    for lo in range(0, len(iu), 2_000_000):
        i, j = iu[lo:lo + 2_000_000], ju[lo:lo + 2_000_000]
        si, sj = sp[i], sp[j]
        if sa == 0 and sb == 0:
            m = np.ones(len(i), bool)
        elif sa == 0 or sb == 0:
            s = sa or sb
            m = (si == s) | (sj == s)
        else:
            m = ((si == sa) & (sj == sb)) | ((si == sb) & (sj == sa))
        d = pos[i[m]] - pos[j[m]]
        d -= np.round(d / Lb) * Lb
        r = np.sqrt((d * d).sum(1))
        r = r[r < rmax]
        h += np.histogram(r, bins=nbins, range=(0, rmax))[0]
    return h


@pytest.mark.parametrize("pair", [(0, 0), (1, 1), (1, 2), (2, 2), (0, 2)])
def test_chain_pair_histogram_matches_numpy(pair):
    N, nch = 512, 3
    cfgs = []
    for k in range(nch):
        pos, sp, box = ka_lattice(N, 1.2, seed=k)
        pos = pos + np.random.default_rng(k).normal(0, 0.08, pos.shape)
        cfgs.append((pos - np.floor(pos / box) * box, sp, box))
    with DeviceContext(nch, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
        ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), cfgs[0][2], 1.0)
        h = ctx.pair_histogram(pair[0], pair[1], rmax=3.0, nbins=50)
        with pytest.raises(Exception, match="half the box"):
            ctx.pair_histogram(0, 0, rmax=5.0, nbins=10)
    ref = sum(numpy_counts(p, s, b, pair[0], pair[1], 3.0, 50) for p, s, b in cfgs)
    # a distance within 1e-12 of a bin edge may fall on either side (sqrt rounding): compare cumulative counts
    assert int(h.sum()) == int(ref.sum())
    assert np.max(np.abs(np.cumsum(h.astype(np.int64)) - np.cumsum(ref))) <= 2


def test_box_pair_histogram_and_gr_normalisation():
    N = 8000
    pos, sp, box = ka_lattice(N, 1.2, seed=3)
    pos = pos + np.random.default_rng(3).normal(0, 0.1, pos.shape)
    pos -= np.floor(pos / box) * box
    with DeviceContext(1, N, 3, 2, M.MODEL_LJ, mode=L.MODE_BOX) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
        ctx.upload(pos, sp, box, 1.0)
        h = ctx.pair_histogram(1, 1, rmax=2.5, nbins=50)
        nA = int((sp == 1).sum())
        r, g = radial_distribution(ctx, nA, nA, float(np.prod(box)), 1, 1, rmax=2.5, nbins=50)
    ref = numpy_counts(pos, sp, box, 1, 1, 2.5, 50)
    assert int(h.sum()) == int(ref.sum())
    assert np.max(np.abs(np.cumsum(h.astype(np.int64)) - np.cumsum(ref))) <= 2
    # normalisation: counts / (unordered pairs x shell volume / V); the jittered lattice has an empty core and
    # integrates to the ideal-gas pair count within a few per cent over [0, 2.5]
    edges = np.linspace(0, 2.5, 51)
    shell = 4.0 / 3.0 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)
    g_ref = ref / (nA * (nA - 1) / 2.0 * shell / float(np.prod(box)))
    assert np.allclose(g, g_ref, rtol=1e-12, atol=1e-12)
    assert g[:10].max() < 0.5 and abs((g * shell).sum() / shell.sum() - 1.0) < 0.1


def test_chain_correlation_matches_reference_restatement(molecule):
    """compute_chain_correlation (src/molecules.jl:224-246) on the reference's trimer fixture, before and after
    MoleculeFlip moves have shuffled the species inside the molecules, against the numpy restatement."""
    from oracle import oracle as O
    zero_based = lambda bonds: [[j - 1 for j in b] for b in bonds]
    par = M.flatten_model_matrix(M.Trimer())
    starts, lens = np.arange(0, 3000, 3), np.full(1000, 3)
    with DeviceContext(2, 3000, 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(par)
        ctx.set_bonds(zero_based(molecule["bonds"]))
        ctx.upload(np.stack([molecule["position"]] * 2), np.stack([molecule["species"]] * 2), molecule["box"],
                   [molecule["temperature"], 8.0])
        with pytest.raises(L.PMCError, match="pmc_set_molecules"):
            ctx.chain_correlation()
        ctx.set_molecules(starts, lens)
        cc0 = ctx.chain_correlation()
        ref0 = O.chain_correlation(molecule["species"], starts, lens)
        assert np.allclose(cc0, ref0, rtol=1e-13, atol=0)
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=0.5, sigma=0.05), dict(kind="flip", prob=0.5)])
        ctx.seed(5)
        ctx.run(4000)
        cc1 = ctx.chain_correlation()
        _, sp = ctx.download()
        ref1 = [O.chain_correlation(sp[c], starts, lens) for c in range(2)]
        assert np.allclose(cc1, ref1, rtol=1e-13, atol=0)
        assert not np.allclose(cc1, cc0)  # flips were accepted, the order parameter moved
        ctx.set_molecules(np.array([0, 3, 7]), np.array([3, 4, 2]))
        with pytest.raises(L.PMCError, match="same length"):
            ctx.chain_correlation()


def test_energy_histogram_counts_chains():
    from particlesmc_b200.observables import EnergyHistogram
    N, nch = 216, 64
    cfgs = []
    for k in range(nch):
        pos, sp, box = ka_lattice(N, 1.2, seed=k)
        pos = pos + np.random.default_rng(k).normal(0, 0.05, pos.shape)
        cfgs.append((pos - np.floor(pos / box) * box, sp, box))
    with DeviceContext(nch, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.KobAndersen()))
        ctx.upload(np.stack([c[0] for c in cfgs]), np.stack([c[1] for c in cfgs]), cfgs[0][2], 1.0)
        with pytest.raises(L.PMCError, match="pmc_init_energy"):
            ctx.energy_histogram(-8.0, -2.0, 10)
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=1.0, sigma=0.05)])
        ctx.seed(1)
        lo, hi = -10.0, 5.0  # the chains relax from the jittered lattice (e/N ~ +1) towards the liquid (~ -5)
        acc = EnergyHistogram(lo, hi, 40)
        total = np.zeros(40, dtype=np.int64)
        for _ in range(3):
            ctx.run(5 * N)
            e = ctx.energy() / N
            acc.add(ctx)
            total += np.histogram(e, bins=40, range=(lo, hi))[0]
        # a value within rounding of a bin edge may land on either side: compare cumulative counts
        assert int(acc.counts.sum()) == int(total.sum()) == 3 * nch
        assert np.max(np.abs(np.cumsum(acc.counts.astype(np.int64)) - np.cumsum(total))) <= 1
        assert abs(np.sum(acc.density()) * ((hi - lo) / 40) - 1.0) < 1e-12
        h_tot = ctx.energy_histogram(lo * N, hi * N, 40, per_particle=False)
        assert int(h_tot.sum()) == nch
