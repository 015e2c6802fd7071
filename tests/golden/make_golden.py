#!/usr/bin/env python
"""Generate the golden fixtures by RUNNING THE REFERENCE ITSELF (build container only).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py [--only NAME ...]

The reference (mhw32/neural-navier-stokes, /root/reference) ships no tests and no golden
vectors, so its own classes, imported unmodified, are the source of
truth.  For each case this script

  1. runs the reference class (src.chorin_fd / src.direct_fd / src.chorin_spectral),
  2. runs the oracle restatement (oracle/oracle.c, oracle/spectral.py) on the same input,
  3. records the max abs / rel-L2 difference between the two in MANIFEST.json (the pin),
  4. writes a small .npz with the inputs' description and selected reference outputs.

/root/reference does not exist on the GPU box, so nothing but this script reads it.
`semi_implicit` needs an ``np.array`` shim because the reference builds a ragged
``np.array([...])`` (src/chorin_fd/simulate.py:105-110) that numpy >= 1.24 rejects; the
shim lives here, the reference files are untouched.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import fd as ofd  # noqa: E402


def ref_modules():
    import tqdm as _tqdm
    import src.boundary as rb
    import src.chorin_fd.simulate as rc
    import src.direct_fd.simulate as rd
    quiet = lambda it=None, **k: it  # noqa: E731
    rc.tqdm = quiet
    rd.tqdm = quiet
    return rb, rc, rd


def cavity_bcs(rb, dx, dy, lid=1.0):
    D, N = rb.DirichletBoundaryCondition, rb.NeumannBoundaryCondition
    u_bc = [D(0, 'left', dx, dy), D(lid, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
    v_bc = [D(0, 'left', dx, dy), D(0, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
    p_bc = [D(0, 'top', dx, dy), N(0, 'bottom', dx, dy), N(0, 'left', dx, dy), N(0, 'right', dx, dy)]
    return u_bc, v_bc, p_bc


def bc_spec(bcs):
    return [[bc.boundary, bc.type, float(bc.value)] for bc in bcs]


def mixed_bcs(rb, dx, dy, seed):
    """Seeded BC lists mixing types, sides, values and ORDER (corners depend on order)."""
    rng = np.random.default_rng(seed)
    D, N = rb.DirichletBoundaryCondition, rb.NeumannBoundaryCondition
    sides = ['left', 'right', 'bottom', 'top']

    def one(scale, p_neu):
        order = list(rng.permutation(4))
        out = []
        for k in order:
            cls = N if rng.random() < p_neu else D
            out.append(cls(float(np.round(rng.normal() * scale, 3)), sides[k], dx, dy))
        if rng.random() < 0.5:      # a repeated side: the later entry wins on that edge
            out.append(D(float(np.round(rng.normal() * scale, 3)), sides[int(rng.integers(4))], dx, dy))
        return out
    return one(0.5, 0.3), one(0.5, 0.3), one(0.2, 0.6)


def smooth_ic(nx, ny, seed, amp=0.3):
    rng = np.random.default_rng(seed)
    x = np.linspace(-1, 1, nx)[:, None]
    y = np.linspace(-1, 1, ny)[None, :]
    out = []
    for _ in range(3):
        f = np.zeros((nx, ny))
        for _k in range(4):
            a, b, c, d = rng.normal(size=4)
            f += a * np.sin(np.pi * (b * x + c * y) + d)
        out.append(amp * f / 4)
    return out


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return float(d / n) if n > 0 else float(d)


def sha(*arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def pin(ref, orc):
    return {"max_abs": [float(np.max(np.abs(r - o))) for r, o in zip(ref, orc)],
            "rel_l2": [rel_l2(o, r) for r, o in zip(ref, orc)],
            "bit_exact": bool(all(np.array_equal(r, o) for r, o in zip(ref, orc)))}


# ------------------------------------------------------------------ cases ----------

def case_chorin_cavity41(man):
    """BASELINE config 1: 41x41 cavity, nt=500, nit=50, explicit (chorin_fd:292-315)."""
    rb, rc, _ = ref_modules()
    nx = ny = 41
    nt, nit, dt, rho, nu, beta = 500, 50, 1e-3, 1, 0.1, 1.25
    dx, dy = 2. / (nx - 1.), 2. / (ny - 1.)
    u_bc, v_bc, p_bc = cavity_bcs(rb, dx, dy)
    z = np.zeros((nx, ny))
    t0 = time.time()
    sysm = rc.NavierStokesSystem(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit, nx=nx, ny=ny,
                                 dt=dt, rho=rho, nu=nu, beta=beta, method='explicit')
    u, v, p = sysm.simulate()
    t_ref = time.time() - t0
    t0 = time.time()
    ou, ov, op, sw = ofd.chorin_simulate(z, z, z, u_bc, v_bc, p_bc, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu,
                                         beta=beta, method='explicit')
    t_orc = time.time() - t0
    frames = np.array([0, 1, 2, 9, 49, 99, 249, 399, 499])
    norms = np.stack([[np.linalg.norm(a[n].ravel()) for n in range(nt)] for a in (u, v, p)])
    np.savez_compressed(os.path.join(HERE, "chorin_cavity41.npz"), frames=frames, u=u[frames], v=v[frames],
                        p=p[frames], norms=norms, sweeps=sw,
                        params=json.dumps(dict(nx=nx, ny=ny, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu, beta=beta,
                                               method='explicit', u_bc=bc_spec(u_bc), v_bc=bc_spec(v_bc),
                                               p_bc=bc_spec(p_bc))))
    man["chorin_cavity41"] = dict(pin=pin((u, v, p), (ou, ov, op)), sha256_ref=sha(u, v, p),
                                  ref_seconds=t_ref, oracle_seconds=t_orc,
                                  sweeps_min=int(sw.min()), sweeps_max=int(sw.max()),
                                  sweeps_mean=float(sw.mean()),
                                  note="sweeps come from the oracle (bit-exact trajectories imply equal counts)")


def case_chorin_mixed(man):
    """Non-square grids, every BC type/side/order, smooth random ICs, explicit."""
    rb, rc, _ = ref_modules()
    out = {}
    pins = {}
    for k, (nx, ny, nt, nit, seed) in enumerate([(24, 17, 12, 50, 11), (9, 30, 8, 20, 12), (33, 33, 10, 7, 13),
                                                 (16, 16, 40, 50, 14)]):
        dx, dy = 2. / (nx - 1), 2. / (ny - 1)
        u_bc, v_bc, p_bc = mixed_bcs(rb, dx, dy, seed)
        if k == 3:      # a converging case so that the early-exit path (< nit-1 sweeps) is pinned
            u_bc, v_bc, p_bc = cavity_bcs(rb, dx, dy, lid=0.01)
        u0, v0, p0 = smooth_ic(nx, ny, seed, amp=0.2 if k < 3 else 1e-4)
        dt, rho, nu, beta = 5e-4, 1.3 if k == 1 else 1, 0.07 + 0.01 * k, 1.25 if k != 2 else 1.6
        sysm = rc.NavierStokesSystem(u0.copy(), v0.copy(), p0.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit, nx=nx,
                                     ny=ny, dt=dt, rho=rho, nu=nu, beta=beta, method='explicit')
        u, v, p = sysm.simulate()
        ou, ov, op, sw = ofd.chorin_simulate(u0, v0, p0, u_bc, v_bc, p_bc, nt=nt, nit=nit, dt=dt, rho=rho,
                                             nu=nu, beta=beta, method='explicit')
        name = "c%d" % k
        out[name + "_u0"], out[name + "_v0"], out[name + "_p0"] = u0, v0, p0
        out[name + "_u"], out[name + "_v"], out[name + "_p"] = u, v, p
        out[name + "_sweeps"] = sw
        out[name + "_params"] = json.dumps(dict(nx=nx, ny=ny, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu, beta=beta,
                                                method='explicit', u_bc=bc_spec(u_bc), v_bc=bc_spec(v_bc),
                                                p_bc=bc_spec(p_bc)))
        pins[name] = dict(pin=pin((u, v, p), (ou, ov, op)), sweeps=[int(s) for s in sw])
    np.savez_compressed(os.path.join(HERE, "chorin_mixed.npz"), **out)
    man["chorin_mixed"] = pins


class _NpShim:
    """numpy proxy whose array() falls back to dtype=object for the ragged list at
    src/chorin_fd/simulate.py:105-110 (harness-side; the reference file is untouched)."""

    def __init__(self, real):
        self._np = real

    def __getattr__(self, k):
        return getattr(self._np, k)

    def array(self, obj, *a, **k):
        try:
            return self._np.array(obj, *a, **k)
        except ValueError:
            return self._np.array(obj, dtype=object)


def case_chorin_semi(man):
    """Reference default method `semi_implicit` (needs nx == ny and the np.array shim)."""
    rb, rc, _ = ref_modules()
    rc.np = _NpShim(np)
    out, pins = {}, {}
    try:
        for k, (n, nt, nit, seed) in enumerate([(21, 10, 50, 21), (41, 6, 50, 22)]):
            dx = dy = 2. / (n - 1)
            if k == 0:
                u_bc, v_bc, p_bc = cavity_bcs(rb, dx, dy)
                u0 = v0 = p0 = np.zeros((n, n))
            else:
                u_bc, v_bc, p_bc = mixed_bcs(rb, dx, dy, seed)
                u0, v0, p0 = smooth_ic(n, n, seed, amp=0.2)
            dt, rho, nu, beta = 1e-3, 1, 0.1, 1.25
            sysm = rc.NavierStokesSystem(u0.copy(), v0.copy(), p0.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit,
                                         nx=n, ny=n, dt=dt, rho=rho, nu=nu, beta=beta, method='semi_implicit')
            u, v, p = sysm.simulate()
            ou, ov, op, sw = ofd.chorin_simulate(u0, v0, p0, u_bc, v_bc, p_bc, nt=nt, nit=nit, dt=dt, rho=rho,
                                                 nu=nu, beta=beta, method='semi_implicit')
            name = "s%d" % k
            out[name + "_u0"], out[name + "_v0"], out[name + "_p0"] = u0, v0, p0
            out[name + "_u"], out[name + "_v"], out[name + "_p"] = u, v, p
            out[name + "_sweeps"] = sw
            out[name + "_params"] = json.dumps(dict(nx=n, ny=n, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu, beta=beta,
                                                    method='semi_implicit', u_bc=bc_spec(u_bc),
                                                    v_bc=bc_spec(v_bc), p_bc=bc_spec(p_bc)))
            pins[name] = dict(pin=pin((u, v, p), (ou, ov, op)), sweeps=[int(s) for s in sw])
    finally:
        rc.np = np
    np.savez_compressed(os.path.join(HERE, "chorin_semi.npz"), **out)
    man["chorin_semi"] = pins


def case_chorin_128(man):
    """BASELINE config 4 members: 128x128 cavities, dt=2e-4, a few (lid, Re) draws x 6 steps."""
    rb, rc, _ = ref_modules()
    nx = ny = 128
    nt, nit, dt, rho, beta = 6, 50, 2e-4, 1, 1.25
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    rng = np.random.default_rng(0)
    lids = rng.uniform(0.5, 1.5, size=4096)
    res = rng.uniform(10, 100, size=4096)
    members = [0, 1, 4095]
    out, pins = {"members": np.array(members), "lid": lids[members], "nu": 1.0 / res[members]}, {}
    z = np.zeros((nx, ny))
    for b in members:
        u_bc, v_bc, p_bc = cavity_bcs(rb, dx, dy, lid=float(lids[b]))
        nu = float(1.0 / res[b])
        t0 = time.time()
        sysm = rc.NavierStokesSystem(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit, nx=nx, ny=ny,
                                     dt=dt, rho=rho, nu=nu, beta=beta, method='explicit')
        u, v, p = sysm.simulate()
        t_ref = time.time() - t0
        ou, ov, op, sw = ofd.chorin_simulate(z, z, z, u_bc, v_bc, p_bc, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu,
                                             beta=beta, method='explicit')
        out["m%d_u" % b], out["m%d_v" % b], out["m%d_p" % b] = u[-1], v[-1], p[-1]
        out["m%d_sweeps" % b] = sw
        pins["m%d" % b] = dict(pin=pin((u, v, p), (ou, ov, op)), ref_seconds=t_ref, sweeps=[int(s) for s in sw])
    out["params"] = json.dumps(dict(nx=nx, ny=ny, nt=nt, nit=nit, dt=dt, rho=rho, beta=beta, method='explicit',
                                    rng="default_rng(0): lid=uniform(0.5,1.5,4096) then Re=uniform(10,100,4096)"))
    np.savez_compressed(os.path.join(HERE, "chorin_ens128.npz"), **out)
    man["chorin_ens128"] = pins


def case_direct(man):
    """direct_fd: module __main__ cavity 50x50 nt=200 (direct_fd:151-185), a mixed non-square
    case, and BASELINE config 2a (256x256 cavity, dt=1e-4) for 60 steps."""
    rb, _, rd = ref_modules()
    out, pins = {}, {}
    cases = [("d0", 50, 50, 200, 50, 1e-3, 1, 0.1, None), ("d1", 20, 31, 15, 9, 5e-4, 1.2, 0.05, 31),
             ("d2", 256, 256, 60, 50, 1e-4, 1, 0.1, None)]
    for name, nx, ny, nt, nit, dt, rho, nu, seed in cases:
        dx, dy = 2. / (nx - 1), 2. / (ny - 1)
        if seed is None:
            u_bc, v_bc, p_bc = cavity_bcs(rb, dx, dy)
            u0 = v0 = p0 = np.zeros((nx, ny))
        else:
            u_bc, v_bc, p_bc = mixed_bcs(rb, dx, dy, seed)
            u0, v0, p0 = smooth_ic(nx, ny, seed, amp=0.2)
        t0 = time.time()
        sysm = rd.NavierStokesSystem(u0.copy(), v0.copy(), p0.copy(), u_bc, v_bc, p_bc, nt=nt, nit=nit, nx=nx,
                                     ny=ny, dt=dt, rho=rho, nu=nu)
        u, v, p = sysm.simulate()
        t_ref = time.time() - t0
        ou, ov, op = ofd.direct_simulate(u0, v0, p0, u_bc, v_bc, p_bc, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu)
        frames = np.unique(np.array([0, 1, nt // 2, nt - 1]))
        if seed is not None:
            out[name + "_u0"], out[name + "_v0"], out[name + "_p0"] = u0, v0, p0
        out[name + "_frames"] = frames
        out[name + "_u"], out[name + "_v"], out[name + "_p"] = u[frames], v[frames], p[frames]
        out[name + "_norms"] = np.stack([[np.linalg.norm(a[n].ravel()) for n in range(nt)] for a in (u, v, p)])
        out[name + "_params"] = json.dumps(dict(nx=nx, ny=ny, nt=nt, nit=nit, dt=dt, rho=rho, nu=nu,
                                                u_bc=bc_spec(u_bc), v_bc=bc_spec(v_bc), p_bc=bc_spec(p_bc)))
        pins[name] = dict(pin=pin((u, v, p), (ou, ov, op)), ref_seconds=t_ref, sha256_ref=sha(u, v, p))
    np.savez_compressed(os.path.join(HERE, "direct_fd.npz"), **out)
    man["direct_fd"] = pins


def case_bc(man):
    """boundary.py apply() on a random field, all sides/types, sequential order."""
    rb, _, _ = ref_modules()
    rng = np.random.default_rng(5)
    nx, ny = 7, 5
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    A0 = rng.normal(size=(nx, ny))
    specs = [('neumann', 'left', 0.3), ('dirichlet', 'top', -1.0), ('neumann', 'top', 0.25),
             ('neumann', 'right', -0.7), ('dirichlet', 'bottom', 2.0), ('neumann', 'bottom', 0.1),
             ('dirichlet', 'left', 0.5), ('dirichlet', 'right', 4.0)]
    A = A0.copy()
    B = A0.copy()
    seq = []
    for typ, side, val in specs:
        cls = rb.DirichletBoundaryCondition if typ == 'dirichlet' else rb.NeumannBoundaryCondition
        bc = cls(val, side, dx, dy)
        A = bc.apply(A)
        ofd.bc_apply(B, bc, dx, dy)
        seq.append(A.copy())
    np.savez_compressed(os.path.join(HERE, "boundary.npz"), A0=A0, seq=np.stack(seq), specs=json.dumps(specs))
    man["boundary"] = dict(pin=pin((A,), (B,)))


CASES = dict(boundary=case_bc, chorin_mixed=case_chorin_mixed, chorin_semi=case_chorin_semi,
             direct_fd=case_direct, chorin_ens128=case_chorin_128, chorin_cavity41=case_chorin_cavity41)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    ofd.build(force=True)
    mpath = os.path.join(HERE, "MANIFEST.json")
    man = {}
    if os.path.exists(mpath):
        with open(mpath) as f:
            man = json.load(f)
    try:
        from tests.golden import make_golden_spectral  # noqa: F401  (added with the spectral path)
        CASES["spectral"] = make_golden_spectral.case_spectral
    except Exception:
        pass
    for name, fn in CASES.items():
        if args.only and name not in args.only:
            continue
        t0 = time.time()
        fn(man)
        print("%-18s done in %.1fs" % (name, time.time() - t0), flush=True)
        man["_env"] = dict(numpy=np.__version__, python=sys.version.split()[0],
                           reference="/root/reference (mhw32/neural-navier-stokes, unmodified)")
        with open(mpath, "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
