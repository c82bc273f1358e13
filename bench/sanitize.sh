#!/bin/bash
# compute-sanitizer passes over small cases of the two sweep kernels (memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards between the warps of a CTA).  Slow (10-100x); not part of the test suite.
# Usage: bench/sanitize.sh [chains] [mol] [box]  (default: all three).
# Round 2 result on a B200: memcheck 0 errors everywhere; racecheck 0 hazards for k_chain_sweep_spec (Atoms with swaps,
# Molecules with flips and swaps); for
# k_box_sweep_all it reports the commit stores of an accepted trial against the position loads of the next trial --
# by design EVERY thread of the cell's CTA performs the same commit stores (same address, same value), so that its own
# program order makes the new position visible to it without a second block barrier per trial (box.cu, trial loop).
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from particlesmc_b200 import _lib as L, models as M
from particlesmc_b200.device import DeviceContext
from particlesmc_b200.synthetic import ka_lattice
par = M.flatten_model_matrix(M.KobAndersen())
which = sys.argv[1]
if which == 'chains':
    pos, sp, box = ka_lattice(1000, 1.2, seed=1)
    with DeviceContext(3, 1000, 3, 2, M.MODEL_LJ) as c:
        c.set_model(par); c.upload(np.stack([pos]*3), np.stack([sp]*3), box, 1.0); c.init_energy()
        c.set_moves([dict(kind='displacement', prob=0.8, sigma=0.05), dict(kind='swap', prob=0.2, species=(1, 2))]); c.seed(3); c.run(600)
        c.set_moves([dict(kind='displacement', prob=1.0, sigma=0.05)]); c.run(600)
        print('chains', c.energy()[:2], c.total_energy()[:2])
elif which == 'mol':
    sys.path.insert(0, 'tests')
    from conftest import load_molecule
    m = load_molecule(); n = 900
    bonds = [[j - 1 for j in b if j <= n] for b in m['bonds'][:n]]
    with DeviceContext(2, n, 3, 3, M.MODEL_KG, molecules=True) as c:
        c.set_model(M.flatten_model_matrix(M.Trimer())); c.set_bonds(bonds); c.set_molecules(np.arange(0, n, 3), np.full(n // 3, 3))
        c.upload(np.stack([m['position'][:n]]*2), np.stack([m['species'][:n]]*2), m['box'], 4.0); c.init_energy()
        c.set_moves([dict(kind='displacement', prob=0.5, sigma=0.06), dict(kind='flip', prob=0.5)]); c.seed(3); c.run(800)
        c.set_moves([dict(kind='displacement', prob=0.4, sigma=0.06), dict(kind='flip', prob=0.3), dict(kind='swap', prob=0.3, species=(1, 3))]); c.run(800)
        print('mol', c.energy()[:2], c.total_energy()[:2])
else:
    pos, sp, box = ka_lattice(8192, 1.2, seed=1)
    with DeviceContext(1, 8192, 3, 2, M.MODEL_LJ, mode=L.MODE_BOX) as c:
        c.set_model(par); c.upload(pos, sp, box, 1.0); c.init_energy()
        c.set_moves([dict(kind='displacement', prob=1.0, sigma=0.05)]); c.seed(3); c.run(2 * 8192)
        print('box', c.energy(), c.total_energy())
PY
for tool in memcheck racecheck; do for w in ${@:-chains mol box}; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py $w > gpurun_out/san_${tool}_$w.log 2>&1
  echo "$tool $w: $(grep -c 'ERROR SUMMARY' gpurun_out/san_${tool}_$w.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/san_${tool}_$w.log | tail -1)"; grep -E "^(chains|mol|box) " gpurun_out/san_${tool}_$w.log | head -2
done; done
