"""ctypes binding of libnns_b200.so (C ABI: include/nns_b200.h).  Fails loudly: no fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NNS_B200_LIB") or os.path.join(_HERE, "libnns_b200.so")   # override: experiments only

SOLVER_CHORIN_FD, SOLVER_DIRECT_FD, SOLVER_CHORIN_SPECTRAL = 0, 1, 2
METHODS = {"explicit": 0, "semi_implicit": 1}
FIELD_U, FIELD_V, FIELD_P = 0, 1, 2
FLAG_CHECK_FINITE = 1
FLAG_PERIODIC_X = 4
SPECTRAL_N_OPERATORS = 29
MAX_BC = 8


class NnsBC(C.Structure):
    _fields_ = [("field", C.c_int32), ("side", C.c_int32), ("type", C.c_int32), ("reserved", C.c_int32),
                ("value", C.c_double)]


class NnsParams(C.Structure):
    _fields_ = [("solver", C.c_int32), ("method", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
                ("batch", C.c_int32), ("nit", C.c_int32), ("dt", C.c_double), ("rho", C.c_double),
                ("nu", C.c_double), ("beta", C.c_double), ("tol", C.c_double), ("device", C.c_int32),
                ("flags", C.c_int32), ("force_x", C.c_double)]


class NnsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnns_b200 error %d: %s" % (code, msg))
        self.code = code


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_PROTOS = {
    "nns_abi_version": (_i32, []),
    "nns_last_error": (C.c_char_p, []),
    "nns_create": (_i32, [C.POINTER(NnsParams), C.POINTER(NnsBC), _i32, _vp, _vp, C.POINTER(_vp)]),
    "nns_destroy": (_i32, [_vp]),
    "nns_launch_count": (_i64, [_vp]),
    "nns_nonfinite_count": (_i32, [_vp, C.POINTER(_i64)]),
    "nns_apply_bc": (_i32, [_vp, _i32, _vp, _vp]),
    "nns_chorin_fd_step": (_i32, [_vp] * 10),
    "nns_chorin_fd_step_host": (_i32, [_vp] * 9),
    "nns_chorin_fd_run": (_i32, [_vp] * 6 + [_i32] + [_vp] * 5),
    "nns_chorin_fd_run_host": (_i32, [_vp] * 6 + [_i32] + [_vp] * 4),
    "nns_chorin_fd_predictor": (_i32, [_vp] * 8),
    "nns_chorin_fd_pressure": (_i32, [_vp] * 6),
    "nns_chorin_fd_correct": (_i32, [_vp] * 7),
    "nns_direct_fd_run": (_i32, [_vp] * 4 + [_i32] + [_vp] * 4),
    "nns_direct_fd_run_host": (_i32, [_vp] * 4 + [_i32] + [_vp] * 3),
    "nns_slab_partition": (_i32, [_i32, _i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "nns_slab_plan": (_i32, [_i32] * 7 + [C.POINTER(_i32)]),
    "nns_slab_apply_bc": (_i32, [_vp, _i32, _vp, _vp]),
    "nns_nccl_unique_id": (_i32, [_vp]),
    "nns_slab_attach": (_i32, [_vp, _i32, _i32, _vp]),
    "nns_slab_exchange": (_i32, [_vp, _vp, _vp]),
    "nns_slab_ipc_export": (_i32, [_vp, _vp]),
    "nns_slab_ipc_connect": (_i32, [_vp, _vp, _vp]),
    "nns_chorin_fd_slab_step": (_i32, [_vp] * 10),
    "nns_direct_fd_slab_run": (_i32, [_vp] * 4 + [_i32, _vp]),
    "nns_slab_last_timing": (_i32, [_vp, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "nns_traj_coarsen": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "nns_traj_observations": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "nns_spectral_set_operators": (_i32, [_vp, C.POINTER(_vp), _i32]),
    "nns_spectral_predictor": (_i32, [_vp] * 8),
    "nns_spectral_correct": (_i32, [_vp] * 9),
    "nns_spectral_run": (_i32, [_vp] * 6 + [_i32] + [_vp] * 4),
    "nns_spectral_run_host": (_i32, [_vp] * 6 + [_i32] + [_vp] * 3),
}

_lib = None


def exported_symbols():
    return sorted(_PROTOS)


def lib():
    """Load the CUDA library; raise if it has not been built (there is no other code path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libnns_b200.so is not built (%s). Run `python neural-navier-stokes_b200/build.py` "
                "or __graft_entry__.build(); this package has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise NnsError(rc, lib().nns_last_error().decode("utf-8", "replace"))


def bc_table(u_bc, v_bc, p_bc):
    """Flatten reference-style BC lists (list order kept) into an nns_bc array."""
    entries = [(f, bc) for f, lst in ((FIELD_U, u_bc), (FIELD_V, v_bc), (FIELD_P, p_bc)) for bc in (lst or [])]
    arr = (NnsBC * max(1, len(entries)))()
    for k, (f, bc) in enumerate(entries):
        if bc.type not in ("dirichlet", "neumann"):
            raise Exception("Boundary type {} not supported".format(bc.type))
        side, typ = bc.abi_codes() if hasattr(bc, "abi_codes") else (
            ("left", "right", "bottom", "top").index(bc.boundary), 0 if bc.type == "dirichlet" else 1)
        arr[k].field, arr[k].side, arr[k].type, arr[k].value = f, side, typ, float(bc.value)
    return arr, len(entries)


class Handle:
    """Owns one nns_handle."""

    def __init__(self, solver, nx, ny, nit, dt, rho, nu, beta=1.25, method="explicit", batch=1,
                 u_bc=(), v_bc=(), p_bc=(), nu_per_member=None, bc_value_per_member=None, tol=0.0,
                 device=-1, check_finite=True, periodic_x=False, force_x=0.0):
        L = lib()
        if method not in METHODS:
            raise Exception("method not recognized: {}".format(method))
        self.params = NnsParams(solver, METHODS[method], nx, ny, batch, nit, float(dt), float(rho),
                                float(nu), float(beta), float(tol), device,
                                (FLAG_CHECK_FINITE if check_finite else 0) | (FLAG_PERIODIC_X if periodic_x else 0),
                                float(force_x))
        arr, n = bc_table(u_bc, v_bc, p_bc)
        self.n_bcs = n
        # the kernels apply Neumann conditions with the solver's spacing; the reference uses the BC object's own
        # dx / dy (src/boundary.py:75-84): a non-zero flux built with another spacing would silently differ
        sdx, sdy = (2. / nx, 2. / ny) if solver == SOLVER_CHORIN_SPECTRAL else (2. / (nx - 1), 2. / (ny - 1))
        for bc in list(u_bc or ()) + list(v_bc or ()) + list(p_bc or ()):
            if getattr(bc, "type", None) == "neumann" and float(bc.value) != 0.0:
                d, s = (bc.dx, sdx) if bc.boundary in ("left", "right") else (bc.dy, sdy)
                if abs(float(d) - s) > 1e-14 * abs(s):
                    raise ValueError("Neumann boundary condition on %r was built with spacing %r, the solver uses %r"
                                     % (bc.boundary, d, s))
        nu_ptr = bv_ptr = None
        if nu_per_member is not None:
            self._nu = np.ascontiguousarray(nu_per_member, dtype=np.float64)
            assert self._nu.shape == (batch,)
            nu_ptr = self._nu.ctypes.data
        if bc_value_per_member is not None:
            self._bv = np.ascontiguousarray(bc_value_per_member, dtype=np.float64)
            assert self._bv.shape == (batch, n)
            bv_ptr = self._bv.ctypes.data
        h = C.c_void_p()
        check(L.nns_create(C.byref(self.params), arr, n, nu_ptr, bv_ptr, C.byref(h)))
        self.h = h
        self.nx, self.ny, self.batch = nx, ny, batch

    def close(self):
        if getattr(self, "h", None):
            lib().nns_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().nns_launch_count(self.h))

    def nonfinite(self):
        c = C.c_int64(0)
        check(lib().nns_nonfinite_count(self.h, C.byref(c)))
        return c.value


def host_ptr(a):
    """Pointer of a C-contiguous float64 (or int32) numpy array, or None."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data
