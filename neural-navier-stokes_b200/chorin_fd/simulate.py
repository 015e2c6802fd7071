"""Chorin projection, finite differences -- drop-in for the reference's
``src/chorin_fd/simulate.py`` (class name, constructor signature, ``simulate`` / ``step`` /
``_init_variables``, return shapes and in-place conventions), executed by the sm_100a kernels of
``libnns_b200.so`` through the C ABI (``nns_chorin_fd_run_host``).

Reference behaviour kept (file:line of the reference):
  * ``dx = 2/(nx-1)``, ``dy = 2/(ny-1)``                               chorin_fd/simulate.py:58
  * ``assert method in ['semi_implicit', 'explicit']``                 :60
  * ``simulate()`` copies the ICs, applies u/v/p BCs, sets u^{-1}=u^0, returns three
    ``(nt, nx, ny)`` float64 arrays                                    :236-271
  * ``step(un, vn, un1, vn1, p)`` returns new arrays for u, v and updates ``p`` IN PLACE :212-234
  * at most ``nit-1`` SOR sweeps, exit when ``max|p - pPrev| <= 5e-6``  :183-200
  * non-finite results raise (the reference turns warnings into errors) :3
"""
import numpy as np

from .. import _lib


class NavierStokesSystem():
    """Wrapper class around a 2D incompressible Navier-Stokes system (chorin_fd).

    Args: u_ic, v_ic, p_ic (np.ndarray (nx, ny)); u_bc, v_bc, p_bc (lists of BoundaryCondition
    objects, applied in list order); nt, nit, nx, ny (int); dt, rho, nu, beta (float);
    method ('explicit' | 'semi_implicit').
    """

    def __init__(self, u_ic, v_ic, p_ic, u_bc, v_bc, p_bc,
                 nt=200, nit=50, nx=50, ny=50, dt=0.001,
                 rho=1, nu=1, beta=1.25, method='semi_implicit'):
        self.u_ic, self.v_ic, self.p_ic = u_ic, v_ic, p_ic
        self.u_bc, self.v_bc, self.p_bc = u_bc, v_bc, p_bc
        self.nt, self.nit, self.dt, self.nx, self.ny = nt, nit, dt, nx, ny
        self.dx, self.dy = 2. / (self.nx - 1), 2. / (self.ny - 1)
        self.rho, self.nu, self.beta = rho, nu, beta
        assert method in ['semi_implicit', 'explicit']
        self.method = method
        self._handle = None
        self.last_sweeps = None      # int32 (nsteps,) SOR sweeps executed per step of the last call

    # -- device handle (created on first use; raises without the CUDA library / a GPU) -----
    def _key(self):
        bcs = tuple((bc.type, bc.boundary, float(bc.value)) for lst in (self.u_bc, self.v_bc, self.p_bc) for bc in lst)
        return (self.nx, self.ny, self.nit, self.dt, self.rho, self.nu, self.beta, self.method, bcs)

    def _h(self):
        # the reference reads its attributes at every step: rebuild the device handle when they have changed
        if self._handle is not None and self._handle_key != self._key():
            self._handle.close()
            self._handle = None
        if self._handle is None:
            self._handle_key = self._key()
            if self.method not in _lib.METHODS:
                raise Exception('method not recognized: {}'.format(self.method))
            self._handle = _lib.Handle(_lib.SOLVER_CHORIN_FD, self.nx, self.ny, self.nit, self.dt, self.rho,
                                       self.nu, beta=self.beta, method=self.method, batch=1,
                                       u_bc=self.u_bc, v_bc=self.v_bc, p_bc=self.p_bc)
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
        sw = np.zeros((nsteps, 1), dtype=np.int32)
        _lib.check(_lib.lib().nns_chorin_fd_run_host(
            h.h, u.ctypes.data, v.ctypes.data, u1.ctypes.data, v1.ctypes.data, p.ctypes.data, nsteps,
            _lib.host_ptr(tu), _lib.host_ptr(tv), _lib.host_ptr(tp), sw.ctypes.data))
        self.last_sweeps = sw[:, 0]
        return tu, tv, tp

    def step(self, un, vn, un1, vn1, p):
        """One time step.  Returns (u^{n+1}, v^{n+1}, p); ``p`` is the caller's array, updated in
        place, exactly as the reference's ``_get_pressure`` mutates its argument."""
        u, v, u1, v1 = (self._check(a, n).copy() for a, n in ((un, 'un'), (vn, 'vn'), (un1, 'un1'), (vn1, 'vn1')))
        pw = self._check(p, 'p')
        if pw is not p:
            pw = pw.copy()
        self._run(u, v, u1, v1, pw, 1, False)
        if pw is not p:
            p[...] = pw
        return u, v, p

    def _init_variables(self):
        u, v, p = self.u_ic, self.v_ic, self.p_ic
        u, v, p = (np.array(a, dtype=np.float64, order='C', copy=True) for a in (u, v, p))
        for bc in self.u_bc:
            u = bc.apply(u)
        for bc in self.v_bc:
            v = bc.apply(v)
        for bc in self.p_bc:
            p = bc.apply(p)
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
    # the reference module's own demo (chorin_fd/simulate.py:274-324): lid-driven cavity
    from ..boundary import DirichletBoundaryCondition, NeumannBoundaryCondition

    nt, nit, nx, ny, dt, rho, nu, beta = 200, 200, 51, 51, 0.001, 1, 0.1, 1.25
    method = 'semi_implicit'
    dx, dy = 2. / (nx - 1.), 2. / (ny - 1.)
    zeros = np.zeros((nx, ny))
    u_bc = [DirichletBoundaryCondition(0, 'left', dx, dy), DirichletBoundaryCondition(1, 'right', dx, dy),
            DirichletBoundaryCondition(0, 'top', dx, dy), DirichletBoundaryCondition(0, 'bottom', dx, dy)]
    v_bc = [DirichletBoundaryCondition(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
    p_bc = [DirichletBoundaryCondition(0, 'top', dx, dy), NeumannBoundaryCondition(0, 'bottom', dx, dy),
            NeumannBoundaryCondition(0, 'left', dx, dy), NeumannBoundaryCondition(0, 'right', dx, dy)]
    system = NavierStokesSystem(zeros, zeros.copy(), zeros.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit, nx=nx,
                                ny=ny, dt=dt, rho=rho, nu=nu, beta=beta, method=method)
    u_data, v_data, p_data = system.simulate()
    np.savez('./data_{}.npz'.format(method), u=u_data, v=v_data, p=p_data)
