"""One large chorin_fd grid split into ROW SLABS over the GPUs of a box (new capability: the reference runs
one process; BASELINE config 5).  One process per GPU (``torch.distributed``, any backend, is only used to hand
out the NCCL id and to gather results); the halo exchange -- single rows of ``p`` after every SOR tick,
``u, v`` once per step -- is NCCL send/recv inside ``libnns_b200.so`` (``nns_chorin_fd_slab_step``).

Rank g owns the global rows ``[row0, row0 + nrows)`` and stores each field as a CUDA float64 tensor
``[nrows + 2, ny]`` (one halo row above and below).  Results equal the single-GPU / reference results: the
lexicographic SOR order is kept across slab boundaries (tile hyperplane, see csrc/chorin_fd_slab.cu).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def partition(nx, world, rank, tile_rows=0):
    """(row0, nrows) owned by `rank` (host logic of nns_slab_partition)."""
    r0, nr = C.c_int32(0), C.c_int32(0)
    _lib.check(_lib.lib().nns_slab_partition(nx, world, rank, tile_rows, C.byref(r0), C.byref(nr)))
    return r0.value, nr.value


def plan(nx, ny, world, rank, tick, sweep, tile_rows=0):
    """dict(TR, TC, nI, nJ, I0, I1, Ilo, Ihi) of nns_slab_plan."""
    out = (C.c_int32 * 8)()
    _lib.check(_lib.lib().nns_slab_plan(nx, ny, world, rank, tile_rows, tick, sweep, out))
    return dict(zip(("TR", "TC", "nI", "nJ", "I0", "I1", "Ilo", "Ihi"), list(out)))


class SlabChorin:
    """chorin_fd (explicit) on this rank's slab of an (nx, ny) grid."""

    def __init__(self, nx, ny, *, u_bc, v_bc, p_bc, nit=50, dt=0.001, rho=1, nu=0.1, beta=1.25, rank=None, world=None,
                 device=None, check_finite=False, p2p=True):
        if not torch.cuda.is_available():
            raise RuntimeError("nns_b200 slabs need a CUDA device (no CPU fallback)")
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world, self.nx, self.ny = rank, world, nx, ny
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.handle = _lib.Handle(_lib.SOLVER_CHORIN_FD, nx, ny, nit, dt, rho, nu, beta=beta, method='explicit', batch=1,
                                  u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, device=self.device.index, check_finite=check_finite)
        self._L = _lib.lib()
        idbuf = np.zeros(128, dtype=np.uint8)
        if world > 1:
            if rank == 0:
                _lib.check(self._L.nns_nccl_unique_id(idbuf.ctypes.data))
            box = [idbuf.tobytes()]
            dist.broadcast_object_list(box, src=0)
            idbuf = np.frombuffer(box[0], dtype=np.uint8).copy()
        with torch.cuda.device(self.device):
            _lib.check(self._L.nns_slab_attach(self.handle.h, rank, world, idbuf.ctypes.data))
        self.p2p = False
        if world > 1 and p2p:
            # peer-memory mailboxes for the per-tick exchange of p (CUDA IPC handles travel through torch.distributed)
            mine = np.zeros(64, dtype=np.uint8)
            _lib.check(self._L.nns_slab_ipc_export(self.handle.h, mine.ctypes.data))
            allh = [None] * world
            dist.all_gather_object(allh, mine.tobytes())
            nb = [np.frombuffer(allh[r], dtype=np.uint8).copy() if 0 <= r < world else None for r in (rank - 1, rank + 1)]
            ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
            with torch.cuda.device(self.device):
                rc = self._L.nns_slab_ipc_connect(self.handle.h, ptr(nb[0]), ptr(nb[1]))
            ok = torch.tensor([1 if rc == 0 else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks use the same exchange
            if int(ok.item()) != 1:
                raise RuntimeError("CUDA IPC peer mapping failed on some rank (%s); construct SlabChorin(p2p=False) to "
                                   "use NCCL send/recv in the tick loop" % self._L.nns_last_error().decode())
            self.p2p = True
        self.row0, self.nrows = partition(nx, world, rank)
        z = lambda: torch.zeros((self.nrows + 2, ny), dtype=torch.float64, device=self.device)  # noqa: E731
        self.u, self.v, self.p, self.u1, self.v1, self._un, self._vn = z(), z(), z(), z(), z(), z(), z()
        self.last_sweeps = 0

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def owned(self, t):
        """View of the owned rows of a local field."""
        return t[1:self.nrows + 1]

    def set_state(self, u, v, p):
        """Copy this rank's rows out of GLOBAL (nx, ny) arrays (every rank passes the same arrays)."""
        sl = slice(self.row0, self.row0 + self.nrows)
        for dst, src in ((self.u, u), (self.v, v), (self.p, p)):
            self.owned(dst).copy_(torch.from_numpy(np.ascontiguousarray(src[sl], dtype=np.float64)))

    def exchange(self, t):
        _lib.check(self._L.nns_slab_exchange(self.handle.h, t.data_ptr(), self._stream()))

    def init_variables(self):
        """_init_variables (chorin_fd/simulate.py:236-249) on the slabs, then u^{-1} := u^0 (:256) and halos."""
        for f, t in ((_lib.FIELD_U, self.u), (_lib.FIELD_V, self.v), (_lib.FIELD_P, self.p)):
            _lib.check(self._L.nns_slab_apply_bc(self.handle.h, f, t.data_ptr(), self._stream()))
            self.exchange(t)
        self.u1.copy_(self.u)
        self.v1.copy_(self.v)

    def step(self):
        sw = C.c_int32(0)
        _lib.check(self._L.nns_chorin_fd_slab_step(self.handle.h, self.u.data_ptr(), self.v.data_ptr(),
                                                   self.u1.data_ptr(), self.v1.data_ptr(), self.p.data_ptr(),
                                                   self._un.data_ptr(), self._vn.data_ptr(), C.byref(sw), self._stream()))
        self.last_sweeps = sw.value
        self.u1, self.u, self._un = self.u, self._un, self.u1
        self.v1, self.v, self._vn = self.v, self._vn, self.v1

    def gather(self, t):
        """Global (nx, ny) numpy array on every rank (tests / small grids)."""
        mine = self.owned(t).cpu().numpy()
        if self.world == 1:
            return mine.copy()
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, (self.row0, mine))
        out = np.empty((self.nx, self.ny))
        for r0, a in parts:
            out[r0:r0 + a.shape[0]] = a
        return out

    def last_sor_timing(self):
        """(ms, ticks) of the SOR tick loop of the last step (device time)."""
        ms, tk = C.c_float(0), C.c_int32(0)
        _lib.check(self._L.nns_slab_last_timing(self.handle.h, C.byref(ms), C.byref(tk)))
        return ms.value, tk.value

    @property
    def launches(self):
        return self.handle.launches


class SlabDirect:
    """direct_fd (src/direct_fd/simulate.py) on this rank's row slab of an (nx, ny) grid: Jacobi sweeps with one halo-row
    exchange per sweep (NCCL over NVLink).  u, v, p are advanced in place, as the reference does."""

    def __init__(self, nx, ny, *, u_bc, v_bc, p_bc, nit=50, dt=0.001, rho=1, nu=0.1, rank=None, world=None, device=None,
                 check_finite=False):
        if not torch.cuda.is_available():
            raise RuntimeError("nns_b200 slabs need a CUDA device (no CPU fallback)")
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world, self.nx, self.ny = rank, world, nx, ny
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.handle = _lib.Handle(_lib.SOLVER_DIRECT_FD, nx, ny, nit, dt, rho, nu, batch=1, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc,
                                  device=self.device.index, check_finite=check_finite)
        self._L = _lib.lib()
        idbuf = np.zeros(128, dtype=np.uint8)
        if world > 1:
            if rank == 0:
                _lib.check(self._L.nns_nccl_unique_id(idbuf.ctypes.data))
            box = [idbuf.tobytes()]
            dist.broadcast_object_list(box, src=0)
            idbuf = np.frombuffer(box[0], dtype=np.uint8).copy()
        with torch.cuda.device(self.device):
            _lib.check(self._L.nns_slab_attach(self.handle.h, rank, world, idbuf.ctypes.data))
        self.row0, self.nrows = partition(nx, world, rank)
        z = lambda: torch.zeros((self.nrows + 2, ny), dtype=torch.float64, device=self.device)  # noqa: E731
        self.u, self.v, self.p = z(), z(), z()

    _stream = SlabChorin._stream
    owned = SlabChorin.owned
    set_state = SlabChorin.set_state
    exchange = SlabChorin.exchange
    gather = SlabChorin.gather

    def sync_halos(self):
        """Halo rows of u, v, p from the neighbouring ranks (once after set_state; the reference applies no BCs before
        step 0, direct_fd/simulate.py:132)."""
        for t in (self.u, self.v, self.p):
            self.exchange(t)

    def run(self, nsteps):
        _lib.check(self._L.nns_direct_fd_slab_run(self.handle.h, self.u.data_ptr(), self.v.data_ptr(), self.p.data_ptr(),
                                                  nsteps, self._stream()))

    @property
    def launches(self):
        return self.handle.launches
