"""Quick GPU check of the wave kernel against the legacy member-at-a-time kernel (same arithmetic per cell:
fields must agree to rounding, sweep counts exactly).  Usage: python scripts/gpu_wave_check.py [B] [steps] [tol]
Run under `timeout`: a protocol error between the two roles would hang the kernel."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 333
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0

from nns_b200.ensemble import ChorinEnsemble, cavity_bcs, cavity_bc_values, cavity_ensemble_params  # noqa: E402

NX = NY = 128
dx = dy = 2. / (NX - 1)
lid, nu = cavity_ensemble_params(B, seed=5)
u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
out = {}
for mode in ("wave", "legacy"):  # legacy is the default
    os.environ["NNS_STREAM_MODE"] = mode
    ens = ChorinEnsemble(B, NX, NY, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=2e-4, rho=1, nu=nu, beta=1.25,
                         method='explicit', bc_values=cavity_bc_values(lid), tol=tol)
    ens.init_variables()
    sw = []
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        ens.step()
        sw.append(ens.sweeps.clone())
    torch.cuda.synchronize()
    print(mode, "done in %.3f s" % (time.time() - t0), "launches", ens.launches, flush=True)
    out[mode] = (ens.u.clone(), ens.v.clone(), ens.p.clone(), torch.stack(sw))
a, b = out["wave"], out["legacy"]
print("sweeps equal:", bool(torch.equal(a[3], b[3])), "min", int(b[3].min()), "max", int(b[3].max()))
for name, x, y in zip("uvp", a[:3], b[:3]):
    d = float((x - y).norm() / y.norm())
    bad = (x != y).flatten(1).any(1).nonzero().flatten()[:8].tolist()
    print(name, "rel l2 diff %.3e" % d, "finite", bool(torch.isfinite(x).all()), "first differing members", bad)
