"""Statistical parity with the reference's own physics validation (examples/lj-mixture).

The reference publishes, for 23 state points of a binary Lennard-Jones mixture (N = 1000, rcut = 4 sigma_11,
unshifted, Displacement 0.9 sigma=0.05 + DiscreteSwap 0.1, 1000 sweeps from a lattice, energies sampled every
100 sweeps, second half averaged), the mean energy per particle with its standard error and the acceptance
rates of both moves (examples/lj-mixture/calculated-energies.csv, produced by run-validation.py:67-168 with the
reference itself).  Here all 23 state points run as 23 chains of ONE device context (per-chain box, temperature
and composition) through the same protocol and must agree within statistical error.  This is the only
reference-generated data that pins the acceptance rule and the swap move (the rule itself lives in Arianna.jl).
"""
import os

import numpy as np
import pytest

from particlesmc_b200 import models as M
from particlesmc_b200.device import DeviceContext

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lattice_config(n1, n2, L, rng):
    """create_config of run-validation.py:36-64: simple-cubic sites, species shuffled."""
    N = n1 + n2
    m = round(N ** (1 / 3))
    assert m ** 3 == N
    g = (np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) * (L / m) - L / 2
    sp = np.array([1] * n1 + [2] * n2, dtype=np.int64)
    rng.shuffle(sp)
    return g - np.floor(g / L) * L, sp


def test_lj_mixture_state_points_match_reference_table():
    tab = np.load(os.path.join(GOLDEN, "lj_mixture_table.npz"))["table"]
    N = 1000
    # run-validation.py:29-34, 98-115: epsilon_1/sigma_1 are written with 2 decimals, the others in full
    eps = [[1.0, 1.1523], [1.1523, 1.3702]]
    sig = [[1.0, 1.0339], [1.0339, 1.0640]]
    mm = [[M.LennardJones(eps[i][j], sig[i][j], rcut=4.0, shift_potential=False) for j in range(2)] for i in range(2)]
    par = M.flatten_model_matrix(mm)
    rng = np.random.default_rng(42)
    nS = len(tab)
    pos, sp, box = [], [], []
    for t, x, rho, *_ in tab:
        L = (N / rho) ** (1 / 3)
        n2 = round(N * x)
        p, s = lattice_config(N - n2, n2, L, rng)
        pos.append(p)
        sp.append(s)
        box.append([L] * 3)
    with DeviceContext(nS, N, 3, 2, M.MODEL_LJ) as ctx:
        ctx.set_model(par)
        ctx.upload(np.stack(pos), np.stack(sp), np.array(box), tab[:, 0])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=0.9, sigma=0.05), dict(kind="swap", prob=0.1, species=(1, 2))])
        ctx.seed(42)
        samples = [ctx.energy() / N]
        for _ in range(10):
            ctx.run(100 * N)
            samples.append(ctx.energy() / N)
        samples = np.array(samples)  # [11][nS], t = 0, 100, ..., 1000
        calls, acc = ctx.counters()
        drift = np.abs(ctx.energy() - ctx.total_energy()) / np.abs(ctx.total_energy())
    assert np.all(drift < 1e-10)
    second_half = samples[len(samples) // 2:]
    mean = second_half.mean(axis=0)
    err = second_half.std(axis=0) / np.sqrt(len(second_half))
    rate = acc / np.maximum(calls, 1)
    bad = []
    for k, (t, x, rho, e_ref, e_err, a_disp, a_swap) in enumerate(tab):
        tol_e = 5.0 * np.hypot(e_err, err[k]) + 0.02
        ok = abs(mean[k] - e_ref) < tol_e and abs(rate[k, 0] - a_disp) < 0.02 and abs(rate[k, 1] - a_swap) < 0.03
        if not ok:
            bad.append((t, x, rho, mean[k], e_ref, tol_e, rate[k, 0], a_disp, rate[k, 1], a_swap))
    assert not bad, bad
    # the move mix itself
    frac = calls[:, 1] / calls.sum(axis=1)
    assert np.all(np.abs(frac - 0.1) < 0.005)
