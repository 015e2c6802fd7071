"""Direct finite-difference solver -- drop-in for the reference's ``src/direct_fd/simulate.py``
(Jacobi pressure Poisson with BCs re-applied every sweep, upwind/central explicit update),
executed by ``libnns_b200.so`` (``nns_direct_fd_run_host``).

Reference behaviour kept (file:line of the reference):
  * constructor signature / defaults (``nu=0.1``, no beta/method)      direct_fd/simulate.py:46-54
  * ``simulate()`` does NOT copy the ICs and applies no BC before step 0: the caller's
    ``u_ic, v_ic, p_ic`` arrays are advanced in place                  :129-144
  * ``step(u, v, p)`` updates its arguments in place and returns them   :90-127
  * exactly ``nit`` Jacobi sweeps, no residual test                     :76-86
"""
import numpy as np

from .. import _lib


class NavierStokesSystem():
    def __init__(self, u_ic, v_ic, p_ic, u_bc, v_bc, p_bc,
                 nt=200, nit=50, nx=50, ny=50, dt=0.001, rho=1, nu=0.1, periodic_x=False, force_x=0.0):
        # periodic_x / force_x: EXTENSION (channel flow of BASELINE.json config 2; the reference has neither): axis 1
        # is periodic and force_x * dt is added to u every step; the BC lists then only name 'left' / 'right'
        self.periodic_x, self.force_x = bool(periodic_x), float(force_x)
        self.u_ic, self.v_ic, self.p_ic = u_ic, v_ic, p_ic
        self.u_bc, self.v_bc, self.p_bc = u_bc, v_bc, p_bc
        self.nt, self.dt, self.nx, self.ny = nt, dt, nx, ny
        self.dx, self.dy = 2. / (self.nx - 1), 2. / (self.ny - 1)
        self.nit, self.rho, self.nu = nit, rho, nu
        self._handle = None

    def _key(self):
        bcs = tuple((bc.type, bc.boundary, float(bc.value)) for lst in (self.u_bc, self.v_bc, self.p_bc) for bc in lst)
        return (self.nx, self.ny, self.nit, self.dt, self.rho, self.nu, bcs, self.periodic_x, self.force_x)

    def _h(self):
        # the reference reads its attributes at every step: rebuild the device handle when they have changed
        if self._handle is not None and self._handle_key != self._key():
            self._handle.close()
            self._handle = None
        if self._handle is None:
            self._handle_key = self._key()
            self._handle = _lib.Handle(_lib.SOLVER_DIRECT_FD, self.nx, self.ny, self.nit, self.dt, self.rho,
                                       self.nu, batch=1, u_bc=self.u_bc, v_bc=self.v_bc, p_bc=self.p_bc,
                                       periodic_x=self.periodic_x, force_x=self.force_x)
        return self._handle

    def _work(self, a, name):
        w = np.ascontiguousarray(a, dtype=np.float64)
        if w.shape != (self.nx, self.ny):
            raise ValueError("%s has shape %r, expected %r" % (name, w.shape, (self.nx, self.ny)))
        return w

    def _run(self, u, v, p, nsteps, trajectory):
        wu, wv, wp = self._work(u, 'u'), self._work(v, 'v'), self._work(p, 'p')
        tu = tv = tp = None
        if trajectory:
            tu, tv, tp = (np.empty((nsteps, self.nx, self.ny)) for _ in range(3))
        _lib.check(_lib.lib().nns_direct_fd_run_host(self._h().h, wu.ctypes.data, wv.ctypes.data, wp.ctypes.data,
                                                     nsteps, _lib.host_ptr(tu), _lib.host_ptr(tv),
                                                     _lib.host_ptr(tp)))
        for w, a in ((wu, u), (wv, v), (wp, p)):     # in-place contract even if a copy was needed
            if w is not a:
                a[...] = w
        return tu, tv, tp

    def step(self, u, v, p):
        self._run(u, v, p, 1, False)
        return u, v, p

    def simulate(self):
        u, v, p = self.u_ic, self.v_ic, self.p_ic
        if self.nt <= 0:
            z = np.empty((0, self.nx, self.ny))
            return z, z.copy(), z.copy()
        return self._run(u, v, p, self.nt, True)


if __name__ == "__main__":
    # the reference module's own demo (direct_fd/simulate.py:147-194)
    from ..boundary import DirichletBoundaryCondition, NeumannBoundaryCondition

    nt, nit, nx, ny, dt, rho, nu = 200, 50, 50, 50, 0.001, 1, 0.1
    dx, dy = 2. / (nx - 1.), 2. / (ny - 1.)
    u_bc = [DirichletBoundaryCondition(0, 'left', dx, dy), DirichletBoundaryCondition(1, 'right', dx, dy),
            DirichletBoundaryCondition(0, 'top', dx, dy), DirichletBoundaryCondition(0, 'bottom', dx, dy)]
    v_bc = [DirichletBoundaryCondition(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
    p_bc = [DirichletBoundaryCondition(0, 'top', dx, dy), NeumannBoundaryCondition(0, 'bottom', dx, dy),
            NeumannBoundaryCondition(0, 'left', dx, dy), NeumannBoundaryCondition(0, 'right', dx, dy)]
    system = NavierStokesSystem(np.zeros((nx, ny)), np.zeros((nx, ny)), np.zeros((nx, ny)), u_bc, v_bc, p_bc,
                                nt=nt, nit=nit, nx=nx, ny=ny, dt=dt, rho=rho, nu=nu)
    u_data, v_data, p_data = system.simulate()
    np.savez('./data.npz', u=u_data, v=v_data, p=p_data)
