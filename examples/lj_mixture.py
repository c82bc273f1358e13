#!/usr/bin/env python
"""The workflow of the reference's examples/lj-mixture (run-validation.py:29-115) on the device path.

Writes a lattice configuration of the binary Lennard-Jones mixture to an XYZ file, loads it back with
``load_chains`` (nsim replicas), runs Displacement 0.9 + DiscreteSwap 0.1 for the second half of 1000 sweeps and
prints the mean energy per particle and the acceptance rates -- the quantities of the reference's
calculated-energies.csv.  Output files have the reference's layout (chains/<k>/energy.dat, trajectory.xyz,
lastframe.xyz, moves/<m>/acceptance.dat).

    python examples/lj_mixture.py [--rho 0.8] [--xA 0.5] [--T 1.2183] [--nsim 64] [--steps 1000] [--out /tmp/ljmix]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import particlesmc_b200 as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rho", type=float, default=0.8)
    ap.add_argument("--xA", type=float, default=0.5)
    ap.add_argument("--T", type=float, default=1.2183)
    ap.add_argument("--nsim", type=int, default=64)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--out", default="/tmp/ljmix")
    a = ap.parse_args()

    # the mixture of run-validation.py:29-34: eps = (1, 1.1523, 1.3702), sigma = (1, 1.0339, 1.0640), rcut = 4.0, unshifted
    eps = {(1, 1): 1.0, (1, 2): 1.1523, (2, 2): 1.3702}
    sig = {(1, 1): 1.0, (1, 2): 1.0339, (2, 2): 1.0640}
    table = {f"{i}-{j}": dict(name="LennardJones", epsilon=eps[(i, j)], sigma=sig[(i, j)], rcut=4.0, shift_potential=False)
             for (i, j) in eps}

    N, n = 1000, 10
    L = (N / a.rho) ** (1 / 3)
    grid = (np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3) + 0.5) * (L / n)
    species = np.where(np.random.default_rng(0).permutation(N) < int(round(a.xA * N)), 1, 2)
    os.makedirs(a.out, exist_ok=True)
    start = P.System(grid, species, a.rho, a.T, [[P.io._get_model(table, i, j) for j in (1, 2)] for i in (1, 2)], compute_energy=False)
    conf = os.path.join(a.out, "initial.xyz")
    with open(conf, "w") as f:
        P.store_trajectory(f, start, 0, P.XYZ())

    chains = P.load_chains(conf, args=dict(model=table, temperature=a.T, nsim=a.nsim), compute_energy=False)
    nA, nB = int(np.count_nonzero(species == 1)), int(np.count_nonzero(species == 2))
    pool = [P.Move(P.Displacement(0, np.zeros(3), 0.0), P.SimpleGaussian(), {"sigma": 0.05}, 0.9),
            P.Move(P.DiscreteSwap(0, 0, (1, 2), (nA, nB), 0.0), P.DoubleUniform(), [], 0.1)]
    sample = P.build_schedule(a.steps, a.steps // 2, max(1, a.steps // 100))
    algorithms = (
        dict(algorithm=P.Metropolis, pool=pool, seed=42, parallel=False, sweepstep=N),
        dict(algorithm=P.StoreCallbacks, callbacks=(P.energy,), scheduler=sample),
        dict(algorithm=P.StoreAcceptance, dependencies=(P.Metropolis,), scheduler=[a.steps]),
        dict(algorithm=P.StoreLastFrames, scheduler=[a.steps]),
    )
    sim = P.Simulation(chains, algorithms, a.steps, path=a.out)
    P.run(sim)
    e = np.stack([np.loadtxt(os.path.join(a.out, "chains", str(k + 1), "energy.dat"))[:, 1] for k in range(len(chains))])
    mean, err = e.mean(), e.mean(axis=1).std(ddof=1) / np.sqrt(len(chains))
    print(f"rho={a.rho} xA={a.xA} T={a.T}: energy/N = {mean:.4f} +- {err:.4f}; "
          f"acceptance displacement {pool[0].accepted_calls / pool[0].total_calls:.3f}, "
          f"swap {pool[1].accepted_calls / pool[1].total_calls:.3f}  ({len(chains)} chains, {a.steps} sweeps)")
    sim.close()


if __name__ == "__main__":
    main()
