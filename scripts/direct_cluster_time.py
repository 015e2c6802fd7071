"""direct_fd 256 x 256 (BASELINE config 2a) on the cluster path: ms/step for several CTA sizes (NNS_DIRECT_THREADS) and sweep counts."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200.ensemble import DirectEnsemble, cavity_bcs
nx = ny = 256
u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
for thr in sys.argv[1:] or ["512"]:
    os.environ["NNS_DIRECT_THREADS"] = thr
    for nit in (2, 26, 50):
        ens = DirectEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=1e-4, rho=1, nu=0.1)
        ens.run(20); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ens.run(1000); e1.record(); torch.cuda.synchronize()
        print("threads", thr, "nit", nit, "ms/step %.4f" % (e0.elapsed_time(e1) / 1000), "finite", bool(torch.isfinite(ens.u).all()))
