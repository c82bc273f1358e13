#!/usr/bin/env python
"""Throughput of the BASELINE configs that are parity cases rather than the headline bench line:
  config 4: 2-D ternary mixture (the reference's own test/config_0 fixture, JBB) with Displacement 0.8 +
            DiscreteSwap (1,3) 0.1 + (2,3) 0.1 (test/gerhard_energy_distribution.jl:63-72), batched chains;
  config 5: 1000 trimers (test/molecule fixture, Trimer/GeneralKG, bonded FENE + non-bonded WCA), Displacement, and
            Displacement 0.8 + MoleculeFlip 0.2 (the pool of examples/ortho-terphenyl).
Device-resident, CUDA events inside the library (pmc_last_run_ms).  One JSON line per config.
    python bench/other_configs.py [--chains 1184] [--sweeps 20]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_config0, load_molecule  # noqa: E402
from particlesmc_b200 import models as M  # noqa: E402
from particlesmc_b200.device import DeviceContext  # noqa: E402


def run(name, ctx, N, chains, sweeps, pool):
    ctx.init_energy()
    ctx.set_moves(pool)
    ctx.seed(42)
    ctx.run(5 * N)
    ms = []
    for _ in range(3):
        ctx.run(sweeps * N)
        ms.append(ctx.last_run_ms())
    e_run, e_tot = ctx.energy(), ctx.total_energy()
    calls, acc = ctx.counters()
    print(json.dumps({"config": name, "metric": "attempted MC moves/sec", "value": chains * sweeps * N / (min(ms) * 1e-3),
                      "chains": chains, "N": N, "sweeps_per_launch": sweeps, "ms_per_launch": min(ms),
                      "acceptance_per_move": (acc.sum(0) / np.maximum(calls.sum(0), 1)).tolist(),
                      "energy_bookkeeping_rel_drift": float(np.max(np.abs(e_run - e_tot) / np.abs(e_tot)))}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1184)
    ap.add_argument("--sweeps", type=int, default=20)
    a = ap.parse_args()
    c = load_config0()
    with DeviceContext(a.chains, c["N"], 2, 3, M.MODEL_SMOOTHLJ) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.JBB()))
        ctx.upload(np.stack([c["position"]] * a.chains), np.stack([c["species"]] * a.chains), c["box"], c["temperature"])
        run("2-D ternary JBB N=1290 (test/config_0), Displacement 0.8 + DiscreteSwap (1,3) 0.1 + (2,3) 0.1", ctx, c["N"],
            a.chains, a.sweeps, [dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="swap", prob=0.1, species=(1, 3)),
                                 dict(kind="swap", prob=0.1, species=(2, 3))])
    with DeviceContext(a.chains, c["N"], 2, 3, M.MODEL_SMOOTHLJ) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.JBB()))
        ctx.upload(np.stack([c["position"]] * a.chains), np.stack([c["species"]] * a.chains), c["box"], c["temperature"])
        run("2-D ternary JBB N=1290 (test/config_0), Displacement only", ctx, c["N"], a.chains, a.sweeps,
            [dict(kind="displacement", prob=1.0, sigma=0.05)])
    m = load_molecule()
    nch = max(1, a.chains // 4)
    with DeviceContext(nch, m["N"], 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.Trimer()))
        ctx.set_bonds([[j - 1 for j in b] for b in m["bonds"]])
        ctx.upload(np.stack([m["position"]] * nch), np.stack([m["species"]] * nch), m["box"], m["temperature"])
        run("1000 trimers N=3000 (test/molecule), Trimer/GeneralKG, Displacement", ctx, m["N"], nch, max(1, a.sweeps // 4),
            [dict(kind="displacement", prob=1.0, sigma=0.05)])
    # the pool examples/ortho-terphenyl actually runs (params-template.toml:59-68): Displacement 0.8 + MoleculeFlip 0.2
    with DeviceContext(nch, m["N"], 3, 3, M.MODEL_KG, molecules=True) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.Trimer()))
        ctx.set_bonds([[j - 1 for j in b] for b in m["bonds"]])
        ctx.set_molecules(np.arange(0, m["N"], 3), np.full(m["N"] // 3, 3))
        ctx.upload(np.stack([m["position"]] * nch), np.stack([m["species"]] * nch), m["box"], m["temperature"])
        run("1000 trimers N=3000 (test/molecule), Trimer/GeneralKG, Displacement 0.8 + MoleculeFlip 0.2", ctx, m["N"], nch,
            max(1, a.sweeps // 4), [dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="flip", prob=0.2)])


if __name__ == "__main__":
    main()
