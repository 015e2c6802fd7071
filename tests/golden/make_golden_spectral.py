"""Spectral golden fixtures: run the REFERENCE's chorin_spectral class (build container only)
and pin oracle/spectral.py per operator.  Called from make_golden.py (`--only spectral`)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def smooth_state(N, seed, amp=1.0):
    """Seeded O(1) fields: sums of sin/cos of the Gauss-Lobatto mesh."""
    rng = np.random.default_rng(seed)
    x = np.cos(np.pi * np.arange(N) / (N - 1.0))
    X, Y = np.meshgrid(x, x, indexing="ij")
    out = []
    for _ in range(5):
        f = np.zeros((N, N))
        for _k in range(3):
            a, b, c, d = rng.normal(size=4)
            f += a * np.sin(b * X + c * Y + d)
        out.append(amp * f / 3)
    return out


def rel(a, b):
    n = np.linalg.norm(np.asarray(b).ravel())
    d = np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())
    return float(d / n) if n > 0 else float(d)


def case_spectral(man):
    import src.boundary as rb
    import src.chorin_spectral.simulate as rs
    from oracle import spectral as osp
    D = rb.DirichletBoundaryCondition
    out, pins = {}, {}
    for N in (21, 51, 127):
        dx = dy = 2. / (N - 1.)
        u_bc = [D(0, 'left', dx, dy), D(1, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
        v_bc = [D(0, 'left', dx, dy), D(0, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
        dt, rho = 1e-3, 1
        z = np.zeros((N, N))
        ref = rs.NavierStokesSystem(z, z, z, u_bc, v_bc, nt=1, nit=50, nx=N, ny=N, dt=dt, rho=rho, nu=0.1, beta=1.25)
        S = osp.Setup(N, N, u_bc, v_bc)
        assert S.is_real()
        pin = {"setup_max_abs": {
            "Dx": float(np.max(np.abs(S.Dx - ref.Dx))), "Dx_sqr": float(np.max(np.abs(S.Dx_sqr - ref.Dx_sqr))),
            "DPx": float(np.max(np.abs(S.DPx - ref.DPx))), "DxDPx": float(np.max(np.abs(S.DxDPx - ref.DxDPx))),
            "u_lambda_x": float(np.max(np.abs(S.helm['u']['lx'] - ref.u_Dx_lambda))),
            "u_P": float(np.max(np.abs(S.helm['u']['P'] - ref.u_Dx_P))),
            "p_lambda_x": float(np.max(np.abs(S.pres['lx'] - ref.DxDPx_lambda))),
            "p_Pinv": float(np.max(np.abs(S.pres['Pinv'] - ref.DxDPx_P_inv)))}}
        states = {}
        u0, v0, p0 = ref._init_variables()                 # cavity start: zeros + BCs
        states["cav"] = (u0, v0, u0.copy(), v0.copy(), p0)
        states["rnd"] = tuple(smooth_state(N, 100 + N))
        for name, (un, vn, un1, vn1, p) in states.items():
            ui, vi = ref._predictor_step(un, vn, un1, vn1)
            u2, v2, p2 = ref._correction_step(ui, vi, p)
            Q = p2[1:-1, 1:-1]
            oui, ovi = osp.predictor(S, dt, un, vn, un1, vn1)
            ou2, ov2, op2, oQ = osp.correction(S, dt, rho, ui, vi, p)
            pin[name] = {"ui": rel(oui, ui), "vi": rel(ovi, vi), "Q": rel(oQ, Q), "u_corrected": rel(ou2, u2),
                         "v_corrected": rel(ov2, v2), "absQ_max": float(np.max(np.abs(Q)))}
            if N <= 51 or name == "rnd":
                key = "N%d_%s_" % (N, name)
                if name == "rnd":
                    for k, a in zip(("un", "vn", "un1", "vn1", "p"), (un, vn, un1, vn1, p)):
                        out[key + k] = a
                out[key + "ui"], out[key + "vi"], out[key + "Q"] = ui, vi, Q
                out[key + "u2"], out[key + "v2"] = u2, v2
        # a few setup entries (portable spot checks, not whole matrices)
        out["N%d_Dx_row1" % N] = ref.Dx[1].copy()
        out["N%d_Dx_sqr_diag" % N] = np.diag(ref.Dx_sqr).copy()
        out["N%d_DxDPx_diag" % N] = np.diag(ref.DxDPx).copy()
        out["N%d_p_lambda_sorted" % N] = np.sort(ref.DxDPx_lambda)
        out["N%d_u_lambda_sorted" % N] = np.sort(ref.u_Dx_lambda)
        pins["N%d" % N] = pin
    out["params"] = json.dumps(dict(dt=1e-3, rho=1, lid=1.0, note="cavity BCs of chorin_spectral:600-612"))
    np.savez_compressed(os.path.join(HERE, "spectral.npz"), **out)
    man["spectral"] = pins
