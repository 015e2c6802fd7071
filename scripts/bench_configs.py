"""Timing of the single-simulation configs of BASELINE.json (configs[0..2]) through the drop-in classes
(host buffers in, full trajectories out: the reference's own API), next to the CPU oracle (C port of the reference,
one thread, bounded sample).  The headline config (ensembles) is bench.py; these are parity-test cases, reported here
for completeness.  Usage: python scripts/bench_configs.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nns_b200  # noqa: E402,F401
from nns_b200.chorin_fd.simulate import NavierStokesSystem as Chorin  # noqa: E402
from nns_b200.direct_fd.simulate import NavierStokesSystem as Direct  # noqa: E402
from nns_b200.ensemble import cavity_bcs  # noqa: E402
from oracle import fd as ofd  # noqa: E402


def timed(f, reps=3):
    f()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t0)
    return best


out = []
# config 1: chorin_fd cavity 41 x 41, nt = 500, nit = 50, explicit
nx = ny = 41
u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
z = np.zeros((nx, ny))
kw = dict(nt=500, nit=50, dt=1e-3, rho=1, nu=0.1)
t = timed(lambda: Chorin(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, nx=nx, ny=ny, beta=1.25, method='explicit', **kw).simulate())
tc = timed(lambda: ofd.chorin_simulate(z, z, z, u_bc, v_bc, p_bc, beta=1.25, method='explicit', **kw), reps=1)
out.append({"config": "chorin_fd cavity 41x41 nt=500 nit=50 (simulate(), host arrays in, trajectories out)", "seconds": t,
            "ms_per_step": 1e3 * t / 500, "cell_updates_per_s": nx * ny * 500 / t, "cpu_port_seconds": tc,
            "cpu_port_cell_updates_per_s": nx * ny * 500 / tc, "note": "one 13 KiB problem on one SM: latency-bound"})
# config 2a: direct_fd cavity 256 x 256, nt = 2000, nit = 50, dt = 1e-4
nx = ny = 256
u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
z = np.zeros((nx, ny))
kw = dict(nt=2000, nit=50, dt=1e-4, rho=1, nu=0.1)
t = timed(lambda: Direct(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, nx=nx, ny=ny, **kw).simulate(), reps=2)
kc = dict(kw, nt=50)
tc = timed(lambda: ofd.direct_simulate(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, **kc), reps=1)
out.append({"config": "direct_fd cavity 256x256 nt=2000 nit=50 dt=1e-4 (simulate(), trajectories out: 3.1 GB to the host)",
            "seconds": t, "ms_per_step": 1e3 * t / 2000, "cell_updates_per_s": nx * ny * 2000 / t,
            "cpu_port_cell_updates_per_s": nx * ny * 50 / tc, "cpu_port_sample": "50 steps, 1 thread"})
# config 3: chorin_spectral N = 127 (odd N: real spectrum), timing only (the reference scheme overflows)
try:
    from nns_b200.chorin_spectral.simulate import NavierStokesSystem as Spectral
    from nns_b200.boundary import DirichletBoundaryCondition as D
    N = 127
    dxs = 2. / (N - 1)
    ub = [D(0, 'left', dxs, dxs), D(1, 'right', dxs, dxs), D(0, 'top', dxs, dxs), D(0, 'bottom', dxs, dxs)]
    vb = [D(0, 'left', dxs, dxs), D(0, 'right', dxs, dxs), D(0, 'top', dxs, dxs), D(0, 'bottom', dxs, dxs)]
    z = np.zeros((N, N))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = Spectral(z.copy(), z.copy(), z.copy(), ub, vb, nt=1000, nit=50, nx=N, ny=N, dt=1e-3, rho=1, nu=1)
        t0 = time.perf_counter()
        try:
            s.simulate()
            ok = "finite"
        except Exception as e:      # the reference raises on overflow (warnings are errors); so does the drop-in
            ok = "raised %s" % type(e).__name__
        t = time.perf_counter() - t0
    out.append({"config": "chorin_spectral N=127 nt=1000 (operator setup on the host included)", "seconds": t, "outcome": ok})
except Exception as e:
    out.append({"config": "chorin_spectral", "error": repr(e)})
for o in out:
    print(json.dumps(o))
