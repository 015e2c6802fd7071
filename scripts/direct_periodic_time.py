import os, sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import nns_b200
from nns_b200.ensemble import DirectEnsemble
D, Nm = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
nx = ny = 256; dx = dy = 2./(nx-1)
walls = lambda cls: [cls(0.0,'left',dx,dy), cls(0.0,'right',dx,dy)]
ens = DirectEnsemble(1, nx, ny, u_bc=walls(D), v_bc=walls(D), p_bc=walls(Nm), nit=50, dt=1e-4, rho=1, nu=0.1, periodic_x=True, force_x=1.0)
ens.run(20); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); ens.run(1000); e1.record(); torch.cuda.synchronize()
print("periodic channel 256^2 ms/step %.4f" % (e0.elapsed_time(e1)/1000), float(ens.u.mean()))
