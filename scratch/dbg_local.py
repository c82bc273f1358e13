import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from conftest import load_config0
import particlesmc_b200 as P
from particlesmc_b200.systems import make_context
c=load_config0()
s=P.System(c["position"],c["species"],c["density"],c["temperature"],P.JBB(),list_type=P.LinkedList)
print("E/N", P.energy(s))
with make_context([s]) as ctx:
    e=ctx.local_energy(0); print("local first", e[:4], e.sum()/2)
    ctx.init_energy(); print("init", ctx.energy())
    e=ctx.local_energy(0); print("local second", e[:4], e.sum()/2)
    print("tot", ctx.total_energy())
