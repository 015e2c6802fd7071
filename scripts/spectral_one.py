import os, sys
sys.path.insert(0, "/root/repo")
os.environ["NNS_SPECTRAL_NOGRAPH"] = "1"
import numpy as np, torch, nns_b200
from nns_b200.ensemble import SpectralEnsemble
D = nns_b200.DirichletBoundaryCondition
N = 127; dx = 2. / (N - 1.)
u_bc = [D(0, 'left', dx, dx), D(1, 'right', dx, dx), D(0, 'top', dx, dx), D(0, 'bottom', dx, dx)]
v_bc = [D(0, s, dx, dx) for s in ('left', 'right', 'top', 'bottom')]
ens = SpectralEnsemble(1, N, N, u_bc=u_bc, v_bc=v_bc, dt=1e-3, rho=1)
ens.run(6); torch.cuda.synchronize()
