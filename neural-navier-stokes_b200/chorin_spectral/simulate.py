"""Chorin projection, Chebyshev pseudo-spectral -- drop-in for the reference's
``src/chorin_spectral/simulate.py`` (class name, constructor signature without ``p_bc``, ``simulate`` /
``step`` / ``_init_variables``, return shapes), executed by ``libnns_b200.so`` (``nns_spectral_*``).

Reference behaviour kept (file:line of the reference):
  * ``dx = 2/nx``, ``dy = 2/ny``; ``nu``, ``beta``, ``nit`` stored but unused        chorin_spectral:41-52
  * the operators are built once in the constructor (host, LAPACK)                    :59-199
  * Neumann BCs raise ``NotImplementedError``                                          :221
  * ``step(un, vn, un1, vn1, p)`` returns new ``(u, v, p)`` arrays                       :54-57
  * a complex eigen-decomposition (even N >= 64) makes the first step raise
    ``ComplexWarning`` (the reference promotes warnings to errors, :3, :379); overflow raises too.
The per-step work -- 28 dense (N-2)^3 fp64 products, the AB2 right-hand side, the eigen-space
divisions and the boundary closure -- runs on the GPU.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .operators import SpectralOperators

try:                                    # numpy >= 1.25
    from numpy.exceptions import ComplexWarning
except Exception:                       # pragma: no cover
    ComplexWarning = np.ComplexWarning


class NavierStokesSystem():
    def __init__(self, u_ic, v_ic, p_ic, u_bc, v_bc, nt=200, nit=50,
                 nx=50, ny=50, dt=0.001, rho=1, nu=1, beta=1.25):
        self.u_ic, self.v_ic, self.p_ic = u_ic, v_ic, p_ic
        self.u_bc, self.v_bc = u_bc, v_bc
        self.nt, self.nit, self.dt, self.nx, self.ny = nt, nit, dt, nx, ny
        self.dx, self.dy = 2. / self.nx, 2. / self.ny
        self.rho, self.nu, self.beta = rho, nu, beta
        self._pseudospectral_setup()
        self._handle = None

    def _pseudospectral_setup(self):
        self.ops = SpectralOperators(self.nx, self.ny, self.u_bc, self.v_bc)
        o = self.ops        # attribute names of the reference, for callers that inspect them
        self.Dx, self.Dy, self.Dx_sqr, self.Dy_sqr = o.Dx, o.Dy, o.Dx_sqr, o.Dy_sqr
        self.DPx, self.DPy, self.DxDPx, self.DyDPy = o.DPx, o.DPy, o.DxDPx, o.DyDPy
        self.u_Dx_lambda, self.u_Dx_P = o.helm['u']['lx'], o.helm['u']['P']
        self.DxDPx_lambda, self.DxDPx_P_inv = o.pres['lx'], o.pres['Pinv']

    def _h(self):
        # dt and rho are read at every step by the reference (the operators are fixed at construction, :41-52)
        if self._handle is not None and self._handle_key != (self.dt, self.rho):
            self._handle.close()
            self._handle = None
        if self._handle is None:
            self._handle_key = (self.dt, self.rho)
            if not self.ops.is_real():
                raise ComplexWarning("Casting complex values to real discards the imaginary part")
            h = _lib.Handle(_lib.SOLVER_CHORIN_SPECTRAL, self.nx, self.ny, self.nit, self.dt, self.rho, self.nu,
                            beta=self.beta, batch=1)
            arrs = self.ops.abi_arrays()
            ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
            _lib.check(_lib.lib().nns_spectral_set_operators(h.h, ptrs, len(arrs)))
            self._handle = h
        return self._handle

    def _check(self, a, name):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (self.nx, self.ny):
            raise ValueError("%s has shape %r, expected %r" % (name, a.shape, (self.nx, self.ny)))
        return a

    def _run(self, u, v, u1, v1, p, nsteps, trajectory):
        h = self._h()
        tu = tv = tp = None
        if trajectory:
            tu, tv, tp = (np.empty((nsteps, self.nx, self.ny)) for _ in range(3))
        _lib.check(_lib.lib().nns_spectral_run_host(h.h, u.ctypes.data, v.ctypes.data, u1.ctypes.data, v1.ctypes.data,
                                                    p.ctypes.data, nsteps, _lib.host_ptr(tu), _lib.host_ptr(tv),
                                                    _lib.host_ptr(tp)))
        return tu, tv, tp

    def step(self, un, vn, un1, vn1, p):
        u, v, u1, v1, pw = (self._check(a, n).copy() for a, n in
                            ((un, 'un'), (vn, 'vn'), (un1, 'un1'), (vn1, 'vn1'), (p, 'p')))
        self._run(u, v, u1, v1, pw, 1, False)
        return u, v, pw

    def _init_variables(self):
        u, v, p = (np.array(a, dtype=np.float64, order='C', copy=True) for a in (self.u_ic, self.v_ic, self.p_ic))
        for bc in self.u_bc:
            u = bc.apply(u)
        for bc in self.v_bc:
            v = bc.apply(v)
        return u, v, p

    def simulate(self):
        u, v, p = self._init_variables()
        u, v, p = self._check(u, 'u_ic'), self._check(v, 'v_ic'), self._check(p, 'p_ic')
        u1, v1 = u.copy(), v.copy()
        if self.nt <= 0:
            z = np.empty((0, self.nx, self.ny))
            return z, z.copy(), z.copy()
        return self._run(u, v, u1, v1, p, self.nt, True)


if __name__ == "__main__":
    # the reference module's own demo (chorin_spectral/simulate.py:580-621)
    from ..boundary import DirichletBoundaryCondition

    nt, nit, nx, ny, dt, rho, nu, beta = 200, 200, 51, 51, 0.001, 1, 0.1, 1.25
    dx, dy = 2. / (nx - 1.), 2. / (ny - 1.)
    u_bc = [DirichletBoundaryCondition(0, 'left', dx, dy), DirichletBoundaryCondition(1, 'right', dx, dy),
            DirichletBoundaryCondition(0, 'top', dx, dy), DirichletBoundaryCondition(0, 'bottom', dx, dy)]
    v_bc = [DirichletBoundaryCondition(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
    system = NavierStokesSystem(np.zeros((nx, ny)), np.zeros((nx, ny)), np.zeros((nx, ny)), u_bc, v_bc, nt=nt, nit=nit,
                                nx=nx, ny=ny, dt=dt, rho=rho, nu=nu, beta=beta)
    u_data, v_data, p_data = system.simulate()
    np.savez('./data.npz', u=u_data, v=v_data, p=p_data)
