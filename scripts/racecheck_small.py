import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import ChorinEnsemble, cavity_bcs, cavity_bc_values, cavity_ensemble_params
NX=NY=128; dx=dy=2./(NX-1)
B=int(sys.argv[1]) if len(sys.argv)>1 else 300
lid, nu = cavity_ensemble_params(B, seed=11)
u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
ens = ChorinEnsemble(B, NX, NY, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=int(os.environ.get("NIT","50")), dt=2e-4, rho=1, nu=nu, beta=1.25, method="explicit", bc_values=cavity_bc_values(lid))
ens.init_variables()
ens.step()
torch.cuda.synchronize()
print("done", float(ens.p.abs().max()))
