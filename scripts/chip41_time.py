"""chorin_fd 41 x 41 cavity (BASELINE config 1) on the chip path: ms/step against the sweep count."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import ChorinEnsemble, cavity_bcs
nx = ny = 41
u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
for nit in (2, 10, 26, 50):
    ens = ChorinEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=1e-3, rho=1, nu=0.1, beta=1.25, method="explicit")
    ens.init_variables(); ens.run(50); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ens.run(500); e1.record(); torch.cuda.synchronize()
    print("nit", nit, "us/step %.2f" % (e0.elapsed_time(e1) / 500 * 1e3), "launches per run(500)", "sweeps", int(ens.sweeps.flatten()[-1]) if hasattr(ens, "sweeps") and ens.sweeps is not None else "?")
