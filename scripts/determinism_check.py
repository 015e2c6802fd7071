import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import ChorinEnsemble, cavity_bcs, cavity_bc_values, cavity_ensemble_params
NX=NY=128; dx=dy=2./(NX-1)
B=int(sys.argv[1]) if len(sys.argv)>1 else 333
lid, nu = cavity_ensemble_params(B, seed=11)
u_bc, v_bc, p_bc = cavity_bcs(dx, dy)

runs=[]
for rep in range(6):
    ens = ChorinEnsemble(B, NX, NY, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=2e-4, rho=1, nu=nu, beta=1.25, method="explicit", bc_values=cavity_bc_values(lid))
    ens.init_variables()
    for _ in range(int(os.environ.get("NSTEPS","3"))): ens.step()
    runs.append((ens.u.clone(), ens.v.clone(), ens.p.clone(), ens.sweeps.clone()))
ref=runs[0]
for r in range(1,6):
    for n,x,y in zip("uvp", runs[r], ref):
        bad=(x!=y).flatten(1).any(1).nonzero().flatten().tolist()
        if bad:
            b=bad[0]; d=(x[b]-y[b]).abs(); idx=(d>0).nonzero()
            print("run",r,n,"members",bad[:12],"first member: ncells",len(idx),"rows",int(idx[:,0].min()),int(idx[:,0].max()),"cols",int(idx[:,1].min()),int(idx[:,1].max()),"max %.2e"%float(d.max()))
    if not torch.equal(runs[r][3], ref[3]): print("run", r, "sweeps differ")
print("done")
