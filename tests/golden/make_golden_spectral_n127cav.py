"""Fixture for the N = 127 cavity start of chorin_spectral (BASELINE config 3 size), recorded from the REFERENCE's own
classes (imported from /root/reference in the build container).  Kept apart from spectral.npz, whose generator skips this
case for size.  Usage: PYTHONPATH=/root/reference python tests/golden/make_golden_spectral_n127cav.py"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")
import src.boundary as rb  # noqa: E402
import src.chorin_spectral.simulate as rs  # noqa: E402

N = 127
D = rb.DirichletBoundaryCondition
dx = dy = 2. / (N - 1.)
u_bc = [D(0, 'left', dx, dy), D(1, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
v_bc = [D(0, 'left', dx, dy), D(0, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
z = np.zeros((N, N))
ref = rs.NavierStokesSystem(z, z, z, u_bc, v_bc, nt=1, nit=50, nx=N, ny=N, dt=1e-3, rho=1, nu=0.1, beta=1.25)
u0, v0, p0 = ref._init_variables()
ui, vi = ref._predictor_step(u0, v0, u0.copy(), v0.copy())
u2, v2, p2 = ref._correction_step(ui, vi, p0)
np.savez_compressed(os.path.join(HERE, "spectral_n127cav.npz"), N127_cav_ui=ui, N127_cav_vi=vi, N127_cav_Q=p2[1:-1, 1:-1],
                    N127_cav_u2=u2, N127_cav_v2=v2)
print("wrote spectral_n127cav.npz", float(np.abs(p2).max()))
