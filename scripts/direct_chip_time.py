"""direct_fd on the chip path: the reference module's own demo size (50 x 50), us/step against the sweep count."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import DirectEnsemble, cavity_bcs
for nx in (50, 96):
    ny = nx
    u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
    for nit in (2, 50):
        ens = DirectEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=1e-4, rho=1, nu=0.1)
        ens.run(20); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.run(500); e1.record(); torch.cuda.synchronize()
        print("direct_fd chip %dx%d nit %d: %.2f us/step" % (nx, ny, nit, e0.elapsed_time(e1) / 500 * 1e3))
