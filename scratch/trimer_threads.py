import sys, json, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import load_molecule
from particlesmc_b200 import models as M
from particlesmc_b200.device import DeviceContext
m = load_molecule()
for threads, nch in ((128, 296), (256, 296), (256, 148), (192, 296)):
    with DeviceContext(nch, m["N"], 3, 3, M.MODEL_KG, molecules=True, threads=threads) as ctx:
        ctx.set_model(M.flatten_model_matrix(M.Trimer()))
        ctx.set_bonds([[j - 1 for j in b] for b in m["bonds"]])
        ctx.set_molecules(np.arange(0, 3000, 3), np.full(1000, 3))
        ctx.upload(np.stack([m["position"]] * nch), np.stack([m["species"]] * nch), m["box"], m["temperature"])
        ctx.init_energy()
        ctx.set_moves([dict(kind="displacement", prob=0.8, sigma=0.05), dict(kind="flip", prob=0.2)])
        ctx.seed(1)
        ctx.run(2 * 3000)
        ms = []
        for _ in range(2):
            ctx.run(5 * 3000); ms.append(ctx.last_run_ms())
        print(threads, nch, nch * 5 * 3000 / (min(ms) * 1e-3))
