import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import nns_b200
from nns_b200.ensemble import DirectEnsemble, cavity_bcs
for shape in ((40,36),(256,256)):
    nx, ny = shape
    u_bc, v_bc, p_bc = cavity_bcs(2./(nx-1), 2./(ny-1))
    for mode in ("chip","cluster","stream"):
        os.environ["NNS_DIRECT_MODE"]=mode
        ens = DirectEnsemble(2, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=1e-4, rho=1, nu=0.1)
        try:
            ens.run(3); torch.cuda.synchronize()
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            l0=ens.launches
            e0.record(); ens.run(20); e1.record(); torch.cuda.synchronize()
            print(shape, mode, "launches per run(20)", ens.launches-l0, "ms/step %.4f"%(e0.elapsed_time(e1)/20), "umax %.4f"%float(ens.u.abs().max()))
        except Exception as e:
            print(shape, mode, "ERR", e)
