import numpy as np, sys
sys.path.insert(0,'.')
from particlesmc_b200 import _lib as L, models as M
from particlesmc_b200.device import DeviceContext
from particlesmc_b200.synthetic import ka_lattice
N=1<<20
pos,sp,box=ka_lattice(N,1.2,seed=0)
par=M.flatten_model_matrix(M.KobAndersen())
ctx=DeviceContext(1,N,3,2,M.MODEL_LJ,mode=L.MODE_BOX)
ctx.set_model(par); ctx.upload(pos,sp,box,1.0); ctx.init_energy()
ctx.set_moves([dict(kind="displacement",prob=1.0,sigma=0.05)]); ctx.seed(42)
print("E0", repr(ctx.energy()[0]))
for k in range(3):
    ctx.run(2*N)
    print("run", repr(ctx.energy()[0]), "tot", repr(ctx.total_energy()[0]), ctx.counters())
