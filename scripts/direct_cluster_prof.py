"""One launch of the direct_fd cluster kernel (256 x 256, nit = 50, 40 steps) for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import DirectEnsemble, cavity_bcs
nx = ny = 256
u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
ens = DirectEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=1e-4, rho=1, nu=0.1)
ens.run(5)
torch.cuda.synchronize()
ens.run(40)
torch.cuda.synchronize()
print("done", float(ens.u.abs().max()))
