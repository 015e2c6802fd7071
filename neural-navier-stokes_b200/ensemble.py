"""Batched ensembles of independent simulations, device-resident (new capability: the reference
runs one simulation per process; BASELINE config 4 generates neural_spectral training
trajectories from thousands of cavities with random lid velocity / Reynolds number).

State lives in torch CUDA float64 tensors ``[B, nx, ny]`` (torch is only the allocator / stream
provider); every step is one launch of the fused chorin_fd kernel through the C ABI
(``nns_chorin_fd_step`` / ``nns_chorin_fd_run``).  Members never communicate, so sharding an
ensemble over GPUs is a plain split of the member range (``shard_members``).
"""
import numpy as np
import torch

from . import _lib


def shard_members(batch, rank, world):
    """Contiguous member block [lo, hi) of `rank` out of `world` (remainder to the low ranks)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Ensemble:
    solver = None

    def __init__(self, batch, nx, ny, *, u_bc, v_bc, p_bc, nit=50, dt=0.001, rho=1, nu=0.1, beta=1.25,
                 method='explicit', bc_values=None, device=None, check_finite=False, tol=0.0, periodic_x=False,
                 force_x=0.0):
        if not torch.cuda.is_available():
            raise RuntimeError("nns_b200 ensembles need a CUDA device (no CPU fallback)")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.batch, self.nx, self.ny = int(batch), int(nx), int(ny)
        nu_arr = np.broadcast_to(np.asarray(nu, dtype=np.float64), (self.batch,)) if np.ndim(nu) else None
        with torch.cuda.device(self.device):
            self.handle = _lib.Handle(self.solver, nx, ny, nit, dt, rho, float(np.ravel(nu)[0]), beta=beta,
                                      method=method, batch=self.batch, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc,
                                      nu_per_member=nu_arr, bc_value_per_member=bc_values, tol=tol,
                                      device=self.device.index, check_finite=check_finite, periodic_x=periodic_x,
                                      force_x=force_x)
        self._L = _lib.lib()

    def _zeros(self):
        return torch.zeros((self.batch, self.nx, self.ny), dtype=torch.float64, device=self.device)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _to_dev(dst, src):
        if isinstance(src, torch.Tensor):
            dst.copy_(src)
        else:
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64)))

    def apply_bc(self, field, t):
        _lib.check(self._L.nns_apply_bc(self.handle.h, field, t.data_ptr(), self._stream()))

    @property
    def launches(self):
        return self.handle.launches


class ChorinEnsemble(_Ensemble):
    """B independent chorin_fd simulations (semantics of src/chorin_fd/simulate.py per member).

    nu: scalar or (B,) array; bc_values: optional (B, n_bcs) array overriding the BC values per
    member, columns in the order u_bc + v_bc + p_bc.
    """
    solver = _lib.SOLVER_CHORIN_FD

    def __init__(self, batch, nx, ny, **kw):
        super().__init__(batch, nx, ny, **kw)
        self.u, self.v, self.p = self._zeros(), self._zeros(), self._zeros()
        self.u1, self.v1 = self._zeros(), self._zeros()
        self._un, self._vn = self._zeros(), self._zeros()
        self.sweeps = torch.zeros((self.batch,), dtype=torch.int32, device=self.device)

    def set_state(self, u, v, p, u1=None, v1=None):
        self._to_dev(self.u, u)
        self._to_dev(self.v, v)
        self._to_dev(self.p, p)
        self._to_dev(self.u1, u if u1 is None else u1)
        self._to_dev(self.v1, v if v1 is None else v1)

    def init_variables(self):
        """_init_variables (chorin_fd/simulate.py:236-249) on the device, then u^{-1} := u^0 (:256)."""
        self.apply_bc(_lib.FIELD_U, self.u)
        self.apply_bc(_lib.FIELD_V, self.v)
        self.apply_bc(_lib.FIELD_P, self.p)
        self.u1.copy_(self.u)
        self.v1.copy_(self.v)

    def step(self):
        """One fused step for all members (one kernel launch); rotates (u1, u) <- (u, u_new)."""
        _lib.check(self._L.nns_chorin_fd_step(self.handle.h, self.u.data_ptr(), self.v.data_ptr(),
                                              self.u1.data_ptr(), self.v1.data_ptr(), self.p.data_ptr(),
                                              self._un.data_ptr(), self._vn.data_ptr(), self.sweeps.data_ptr(),
                                              self._stream()))
        self.u1, self.u, self._un = self.u, self._un, self.u1
        self.v1, self.v, self._vn = self.v, self._vn, self.v1

    def run(self, nsteps, trajectory=False, sweeps=False):
        """nsteps steps in one launch.  Returns (traj_u, traj_v, traj_p) as [B, nsteps, nx, ny]
        tensors when trajectory=True, and the [nsteps, B] sweep counts when sweeps=True."""
        tu = tv = tp = sw = None
        if trajectory:
            tu, tv, tp = (torch.empty((self.batch, nsteps, self.nx, self.ny), dtype=torch.float64,
                                      device=self.device) for _ in range(3))
        if sweeps:
            sw = torch.zeros((nsteps, self.batch), dtype=torch.int32, device=self.device)
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        _lib.check(self._L.nns_chorin_fd_run(self.handle.h, self.u.data_ptr(), self.v.data_ptr(),
                                             self.u1.data_ptr(), self.v1.data_ptr(), self.p.data_ptr(), nsteps,
                                             ptr(tu), ptr(tv), ptr(tp), ptr(sw), self._stream()))
        out = []
        if trajectory:
            out += [tu, tv, tp]
        if sweeps:
            out.append(sw)
        return tuple(out) if out else None

    # stage entry points (unit parity)
    def predictor(self, u, v, u1, v1):
        ui, vi = torch.empty_like(u), torch.empty_like(v)
        _lib.check(self._L.nns_chorin_fd_predictor(self.handle.h, u.data_ptr(), v.data_ptr(), u1.data_ptr(),
                                                   v1.data_ptr(), ui.data_ptr(), vi.data_ptr(), self._stream()))
        return ui, vi

    def pressure(self, ui, vi, p):
        sw = torch.zeros((self.batch,), dtype=torch.int32, device=self.device)
        _lib.check(self._L.nns_chorin_fd_pressure(self.handle.h, ui.data_ptr(), vi.data_ptr(), p.data_ptr(),
                                                  sw.data_ptr(), self._stream()))
        return p, sw

    def correct(self, ui, vi, p):
        uo, vo = torch.empty_like(ui), torch.empty_like(vi)
        _lib.check(self._L.nns_chorin_fd_correct(self.handle.h, ui.data_ptr(), vi.data_ptr(), p.data_ptr(),
                                                 uo.data_ptr(), vo.data_ptr(), self._stream()))
        return uo, vo, p


class DirectEnsemble(_Ensemble):
    """B independent direct_fd simulations (semantics of src/direct_fd/simulate.py per member)."""
    solver = _lib.SOLVER_DIRECT_FD

    def __init__(self, batch, nx, ny, **kw):
        kw.pop('beta', None)
        kw.pop('method', None)
        super().__init__(batch, nx, ny, **kw)
        self.u, self.v, self.p = self._zeros(), self._zeros(), self._zeros()

    def set_state(self, u, v, p):
        self._to_dev(self.u, u)
        self._to_dev(self.v, v)
        self._to_dev(self.p, p)

    def run(self, nsteps, trajectory=False):
        tu = tv = tp = None
        if trajectory:
            tu, tv, tp = (torch.empty((self.batch, nsteps, self.nx, self.ny), dtype=torch.float64,
                                      device=self.device) for _ in range(3))
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        _lib.check(self._L.nns_direct_fd_run(self.handle.h, self.u.data_ptr(), self.v.data_ptr(),
                                             self.p.data_ptr(), nsteps, ptr(tu), ptr(tv), ptr(tp), self._stream()))
        return (tu, tv, tp) if trajectory else None

    def step(self):
        self.run(1)


class SpectralEnsemble:
    """B independent chorin_spectral simulations with the same operators (semantics of
    src/chorin_spectral/simulate.py per member); device-resident state, ``run`` replays the step chain from a CUDA graph."""

    def __init__(self, batch, nx, ny, *, u_bc, v_bc, dt=0.001, rho=1, nu=1, beta=1.25, nit=50, device=None,
                 check_finite=False):
        import ctypes as C
        from .chorin_spectral.operators import SpectralOperators
        if not torch.cuda.is_available():
            raise RuntimeError("nns_b200 ensembles need a CUDA device (no CPU fallback)")
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.batch, self.nx, self.ny = int(batch), int(nx), int(ny)
        self.ops = SpectralOperators(self.nx, self.ny, u_bc, v_bc)
        if not self.ops.is_real():
            raise ValueError("the operators of this size have a complex spectrum (the reference fails too: even N >= 64)")
        with torch.cuda.device(self.device):
            self.handle = _lib.Handle(_lib.SOLVER_CHORIN_SPECTRAL, nx, ny, nit, dt, rho, nu, beta=beta, batch=self.batch,
                                      device=self.device.index, check_finite=check_finite)
            arrs = self.ops.abi_arrays()
            ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
            _lib.check(_lib.lib().nns_spectral_set_operators(self.handle.h, ptrs, len(arrs)))
        self._L = _lib.lib()
        z = lambda: torch.zeros((self.batch, self.nx, self.ny), dtype=torch.float64, device=self.device)  # noqa: E731
        self.u, self.v, self.u1, self.v1, self.p = z(), z(), z(), z(), z()

    def set_state(self, u, v, p, u1=None, v1=None):
        for dst, src in ((self.u, u), (self.v, v), (self.p, p), (self.u1, u if u1 is None else u1),
                         (self.v1, v if v1 is None else v1)):
            _Ensemble._to_dev(dst, src)

    def run(self, nsteps, trajectory=False):
        tu = tv = tp = None
        if trajectory:
            tu, tv, tp = (torch.empty((self.batch, nsteps, self.nx, self.ny), dtype=torch.float64,
                                      device=self.device) for _ in range(3))
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._L.nns_spectral_run(self.handle.h, self.u.data_ptr(), self.v.data_ptr(), self.u1.data_ptr(),
                                            self.v1.data_ptr(), self.p.data_ptr(), nsteps, ptr(tu), ptr(tv), ptr(tp), st))
        return (tu, tv, tp) if trajectory else None

    @property
    def launches(self):
        return self.handle.launches


def cavity_bcs(dx, dy, lid=1.0):
    """The lid-driven-cavity BC lists of the reference demos (chorin_fd/simulate.py:296-315)."""
    from .boundary import DirichletBoundaryCondition as D, NeumannBoundaryCondition as N
    u_bc = [D(0, 'left', dx, dy), D(lid, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
    v_bc = [D(0, 'left', dx, dy), D(0, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
    p_bc = [D(0, 'top', dx, dy), N(0, 'bottom', dx, dy), N(0, 'left', dx, dy), N(0, 'right', dx, dy)]
    return u_bc, v_bc, p_bc


def cavity_ensemble_params(batch, seed=0, lo=0, hi=None):
    """BASELINE config 4 draw: rng=default_rng(seed); lid ~ U[0.5,1.5], Re ~ U[10,100], nu = 1/Re.
    Always draws `batch` values so that member b gets the same parameters on any sharding;
    returns the slice [lo, hi)."""
    rng = np.random.default_rng(seed)
    lid = rng.uniform(0.5, 1.5, size=batch)
    re = rng.uniform(10, 100, size=batch)
    hi = batch if hi is None else hi
    return lid[lo:hi].copy(), (1.0 / re)[lo:hi].copy()


def cavity_bc_values(lid):
    """(B, 12) per-member BC value table for cavity_bcs: only the u 'right' entry varies."""
    vals = np.zeros((len(lid), 12))
    vals[:, 1] = lid
    return vals
