"""ctypes binding of oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The C file restates src/chorin_fd/simulate.py, src/direct_fd/simulate.py and
src/boundary.py of the reference; every function there cites the lines it follows.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

SIDES = {"left": 0, "right": 1, "bottom": 2, "top": 3}
TYPES = {"dirichlet": 0, "neumann": 1}


class OrcBC(C.Structure):
    _fields_ = [("side", C.c_int), ("type", C.c_int), ("value", C.c_double)]


class OrcChorinParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nit", C.c_int), ("method", C.c_int),
                ("dt", C.c_double), ("rho", C.c_double), ("nu", C.c_double), ("beta", C.c_double),
                ("n_ubc", C.c_int), ("n_vbc", C.c_int), ("n_pbc", C.c_int)]


class OrcDirectParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nit", C.c_int),
                ("dt", C.c_double), ("rho", C.c_double), ("nu", C.c_double),
                ("n_ubc", C.c_int), ("n_vbc", C.c_int), ("n_pbc", C.c_int)]


def build(force=False):
    """Compile oracle.c -> oracle/_build/liboracle.so (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    base = [gcc, "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math"]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base + extra + ["-o", _SO, src, "-lm"], capture_output=True, text=True)
        if r.returncode == 0:
            return _SO
    raise RuntimeError("oracle build failed:\n" + r.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def bc_array(bcs):
    """[(side,type,value)] or reference-style BC objects -> ctypes array of OrcBC."""
    arr = (OrcBC * max(1, len(bcs)))()
    for k, bc in enumerate(bcs):
        if hasattr(bc, "boundary"):
            side, typ, val = SIDES[bc.boundary], TYPES[bc.type], float(bc.value)
        else:
            side, typ, val = bc
            side = SIDES.get(side, side)
            typ = TYPES.get(typ, typ)
        arr[k].side, arr[k].type, arr[k].value = int(side), int(typ), float(val)
    return arr


def bc_apply(A, bc, dx, dy):
    nx, ny = A.shape
    arr = bc_array([bc])
    lib().orc_bc_apply(_dp(A), nx, ny, arr, C.c_double(dx), C.c_double(dy))
    return A


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def chorin_params(nx, ny, nit, dt, rho, nu, beta, method, u_bc, v_bc, p_bc):
    return OrcChorinParams(nx, ny, nit, 0 if method == "explicit" else 1, float(dt), float(rho),
                           float(nu), float(beta), len(u_bc), len(v_bc), len(p_bc))


def chorin_step(un, vn, un1, vn1, p, u_bc, v_bc, p_bc, *, nit, dt, rho, nu, beta, method="explicit"):
    """One reference step (chorin_fd/simulate.py:212-234).  p is updated IN PLACE.
    Returns (u_new, v_new, p, sweeps)."""
    nx, ny = un.shape
    P = chorin_params(nx, ny, nit, dt, rho, nu, beta, method, u_bc, v_bc, p_bc)
    un, vn, un1, vn1 = map(_f64, (un, vn, un1, vn1))
    assert p.dtype == np.float64 and p.flags.c_contiguous
    uo, vo = np.empty_like(un), np.empty_like(vn)
    s = C.c_int(0)
    rc = lib().orc_chorin_step(C.byref(P), bc_array(u_bc), bc_array(v_bc), bc_array(p_bc),
                               _dp(un), _dp(vn), _dp(un1), _dp(vn1), _dp(p), _dp(uo), _dp(vo), C.byref(s))
    if rc:
        raise RuntimeError("orc_chorin_step rc=%d" % rc)
    return uo, vo, p, s.value


def chorin_simulate(u_ic, v_ic, p_ic, u_bc, v_bc, p_bc, *, nt, nit, dt, rho, nu, beta,
                    method="explicit", want_state=False):
    """Restates NavierStokesSystem.simulate() (chorin_fd/simulate.py:251-271).
    Returns (u, v, p) each (nt, nx, ny), plus sweeps (nt,) [and the final state]."""
    nx, ny = u_ic.shape
    P = chorin_params(nx, ny, nit, dt, rho, nu, beta, method, u_bc, v_bc, p_bc)
    tu, tv, tp = (np.empty((nt, nx, ny)) for _ in range(3))
    sw = np.zeros(nt, dtype=np.int32)
    fin = [np.empty((nx, ny)) for _ in range(5)]
    rc = lib().orc_chorin_simulate(C.byref(P), bc_array(u_bc), bc_array(v_bc), bc_array(p_bc),
                                   _dp(_f64(u_ic)), _dp(_f64(v_ic)), _dp(_f64(p_ic)), nt,
                                   _dp(tu), _dp(tv), _dp(tp), _ip(sw), *[_dp(f) for f in fin])
    if rc:
        raise RuntimeError("orc_chorin_simulate rc=%d" % rc)
    if want_state:
        return tu, tv, tp, sw, fin
    return tu, tv, tp, sw


def _bc_table(bc_lists, batch):
    """bc_lists: one shared BC list (entries are BC objects or (side,type,value) tuples),
    or a ``list`` of per-member ``list``s (len batch)."""
    if len(bc_lists) == 0 or not isinstance(bc_lists[0], list):
        bc_lists = [list(bc_lists)] * batch
    assert len(bc_lists) == batch
    n = len(bc_lists[0])
    arr = (OrcBC * max(1, n * batch))()
    for b in range(batch):
        one = bc_array(bc_lists[b])
        for k in range(n):
            arr[b * n + k] = one[k]
    return arr, n


def chorin_ensemble_run(u, v, u1, v1, p, u_bcs, v_bcs, p_bcs, *, nt, nit, dt, rho, nu, beta,
                        method="explicit", threads=None):
    """Advance [B,nx,ny] state IN PLACE by nt steps; nu scalar or (B,).  Returns (sweeps[nt,B], threads)."""
    B, nx, ny = u.shape
    ua, nu_ = _bc_table(u_bcs, B)
    va, nv_ = _bc_table(v_bcs, B)
    pa, np_ = _bc_table(p_bcs, B)
    nu_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(nu, dtype=np.float64), (B,)))
    P = OrcChorinParams(nx, ny, nit, 0 if method == "explicit" else 1, float(dt), float(rho),
                        float(nu_arr[0]), float(beta), nu_, nv_, np_)
    sw = np.zeros((nt, B), dtype=np.int32)
    for a in (u, v, u1, v1, p):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    if threads:
        set_threads(threads)
    used = lib().orc_chorin_ensemble_run(C.byref(P), _dp(nu_arr), ua, va, pa, B, nt,
                                         _dp(u), _dp(v), _dp(u1), _dp(v1), _dp(p), _ip(sw))
    return sw, used


def direct_params(nx, ny, nit, dt, rho, nu, u_bc, v_bc, p_bc):
    return OrcDirectParams(nx, ny, nit, float(dt), float(rho), float(nu), len(u_bc), len(v_bc), len(p_bc))


def direct_step(u, v, p, u_bc, v_bc, p_bc, *, nit, dt, rho, nu):
    """direct_fd/simulate.py:90-127; u, v, p updated IN PLACE."""
    nx, ny = u.shape
    P = direct_params(nx, ny, nit, dt, rho, nu, u_bc, v_bc, p_bc)
    for a in (u, v, p):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    lib().orc_direct_step(C.byref(P), bc_array(u_bc), bc_array(v_bc), bc_array(p_bc), _dp(u), _dp(v), _dp(p))
    return u, v, p


def direct_simulate(u_ic, v_ic, p_ic, u_bc, v_bc, p_bc, *, nt, nit, dt, rho, nu, trajectory=True):
    """direct_fd/simulate.py:129-144.  Works on copies (the reference mutates its ICs)."""
    nx, ny = u_ic.shape
    P = direct_params(nx, ny, nit, dt, rho, nu, u_bc, v_bc, p_bc)
    u, v, p = (_f64(a).copy() for a in (u_ic, v_ic, p_ic))
    tu = tv = tp = None
    if trajectory:
        tu, tv, tp = (np.empty((nt, nx, ny)) for _ in range(3))
    lib().orc_direct_simulate(C.byref(P), bc_array(u_bc), bc_array(v_bc), bc_array(p_bc),
                              _dp(u), _dp(v), _dp(p), nt, _dp(tu), _dp(tv), _dp(tp))
    if trajectory:
        return tu, tv, tp
    return u, v, p


def direct_ensemble_run(u, v, p, u_bcs, v_bcs, p_bcs, *, nt, nit, dt, rho, nu):
    B, nx, ny = u.shape
    ua, nu_ = _bc_table(u_bcs, B)
    va, nv_ = _bc_table(v_bcs, B)
    pa, np_ = _bc_table(p_bcs, B)
    nu_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(nu, dtype=np.float64), (B,)))
    P = OrcDirectParams(nx, ny, nit, float(dt), float(rho), float(nu_arr[0]), nu_, nv_, np_)
    return lib().orc_direct_ensemble_run(C.byref(P), _dp(nu_arr), ua, va, pa, B, nt, _dp(u), _dp(v), _dp(p))


def max_threads():
    return lib().orc_max_threads()


def host_cores():
    """Cores this process may use (cgroup / affinity aware), whatever OMP_NUM_THREADS says."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def set_threads(n=None):
    """Use n OpenMP threads (default: every core of the host) in the following calls; returns n."""
    n = int(n or host_cores())
    lib().orc_set_threads(n)
    return n


def direct_periodic_simulate(u0, v0, p0, u_bc, v_bc, p_bc, nt, nit, dt, rho, nu, force_x):
    """EXTENSION oracle (reference-unpinned: the reference has no periodic condition and no body force; BASELINE.json
    config 2 asks for a periodic-x channel): src/direct_fd/simulate.py:56-127 restated with numpy, where every column
    slice [1:-1] / [2:] / [0:-2] of the differenced axis 1 becomes the full axis / np.roll(-1) / np.roll(+1) and the
    u equation gains + force_x * dt (Barba's channel flow, step 12 of the "12 steps", in the reference's own operation
    order).  BC entries: (side, type, value) with side in ('left', 'right') only.  Returns (u, v, p) trajectories."""
    nx, ny = u0.shape
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u, v, p = (np.array(a, dtype=np.float64) for a in (u0, v0, p0))

    def apply(A, bcs):
        for bc in bcs:
            side, typ, val = (bc.boundary, bc.type, float(bc.value)) if hasattr(bc, "boundary") else bc
            assert side in ("left", "right"), "periodic x: only the walls (rows 0 and nx-1) carry conditions"
            if typ == "dirichlet":
                A[0 if side == "left" else -1, :] = val
            elif side == "left":
                A[0, :] = A[1, :] - dx * val
            else:
                A[-1, :] = A[-2, :] + dx * val
        return A

    E = lambda a: np.roll(a, -1, axis=1)      # noqa: E731  column j + 1
    W = lambda a: np.roll(a, 1, axis=1)       # noqa: E731  column j - 1
    tu, tv, tp = [], [], []
    for _ in range(nt):
        un, vn = u.copy(), v.copy()
        b = np.zeros_like(u)
        b[1:-1, :] = (rho * (1 / dt * ((E(u)[1:-1] - W(u)[1:-1]) / (2 * dx) + (v[2:, :] - v[0:-2, :]) / (2 * dy))) -
                      ((E(u)[1:-1] - W(u)[1:-1]) / (2 * dx))**2 -
                      2 * ((u[2:, :] - u[0:-2, :]) / (2 * dy) * (E(v)[1:-1] - W(v)[1:-1]) / (2 * dx)) -
                      ((v[2:, :] - v[0:-2, :]) / (2 * dy))**2)
        for _q in range(nit):
            pn = p.copy()
            p[1:-1, :] = (((E(pn)[1:-1] + W(pn)[1:-1]) * dy**2 + (pn[2:, :] + pn[0:-2, :]) * dx**2) /
                          (2 * (dx**2 + dy**2)) - dx**2 * dy**2 / (2 * (dx**2 + dy**2)) * b[1:-1, :])
            p = apply(p, p_bc)
        u[1:-1, :] = (un[1:-1] - un[1:-1] * dt / dx * (un[1:-1] - W(un)[1:-1]) -
                      vn[1:-1] * dt / dy * (un[1:-1] - un[0:-2]) -
                      dt / (2 * rho * dx) * (E(p)[1:-1] - W(p)[1:-1]) +
                      nu * (dt / dx**2 * (E(un)[1:-1] - 2 * un[1:-1] + W(un)[1:-1]) +
                            dt / dy**2 * (un[2:] - 2 * un[1:-1] + un[0:-2])) + force_x * dt)
        v[1:-1, :] = (vn[1:-1] - un[1:-1] * dt / dx * (vn[1:-1] - W(vn)[1:-1]) -
                      vn[1:-1] * dt / dy * (vn[1:-1] - vn[0:-2]) -
                      dt / (2 * rho * dy) * (p[2:] - p[0:-2]) +
                      nu * (dt / dx**2 * (E(vn)[1:-1] - 2 * vn[1:-1] + W(vn)[1:-1]) +
                            dt / dy**2 * (vn[2:] - 2 * vn[1:-1] + vn[0:-2])))
        u, v = apply(u, u_bc), apply(v, v_bc)
        tu.append(u.copy()); tv.append(v.copy()); tp.append(p.copy())
    return np.stack(tu), np.stack(tv), np.stack(tp)
