"""numpy restatement of the reference's Chebyshev pseudo-spectral Chorin step
(src/chorin_spectral/simulate.py) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy is used because the path needs LAPACK (`np.linalg.eig`, `np.linalg.inv`, chorin_spectral:
174-183,196-199); no version of numpy/LAPACK is pinned by the reference, so eigenvector order and
scaling are whatever this numpy returns -- the same call the reference makes.

Parity status: pinned PER OPERATOR against the reference's own class run in the build container
(tests/golden/make_golden_spectral.py): setup matrices, `_predictor_step` outputs and the
pressure Q.  The post-correction u, v are NOT pinnable by any independent implementation: the
reference's own result changes by 2e-3..3e-1 relative when only its GEMM summation order changes
(SURVEY.md section 0.4), because Q ~ 1e16 is cancelled catastrophically (:379-380).

Quirks kept on purpose (they are the spec):
  * c-bar is 2 only for k == 0 (`k == N` never happens for k < N)            :391-393, :470-471
  * the sine formula divides by 2N with N = number of points, x_i uses N-1  :395-399, :472-473
  * D_sqr := D @ D.T with the diagonal replaced by minus the FULL row sum   :492-502
  * `nu` is never used; dx = 2/nx                                           :48, :277-282
"""
import numpy as np


def gauss_lobatto(N, k=1):
    """chorin_spectral/simulate.py:395-399"""
    i = np.arange(N)
    return np.cos(k * np.pi * i / float(N - 1))


def _cbar(k, N):
    return 2 if (k == 0 or k == N) else 1          # :391-393


def D_matrix(N):
    """:443-481 -- scalar loop on purpose (same libm calls as the reference)."""
    D = np.zeros((N, N))
    for i in range(N):
        for j in range(N):
            if i == j:
                continue
            diff = 2 * np.sin((j + i) * np.pi / (2. * N)) * np.sin((j - i) * np.pi / (2. * N))
            D[i, j] = _cbar(i, N) / _cbar(j, N) * (-1) ** (i + j) / diff
    for i in range(N):
        D[i, i] = -np.sum(D[i, :])
    return D


def D_sqr_matrix(N):
    """:483-504"""
    D = D_matrix(N)
    out = np.array(D @ D.T)
    for i in range(N):
        out[i, i] = -np.sum(out[i, :])
    return out


def D_pressure_matrix(N):
    """P_{N-2} derivative matrix on the interior Gauss-Lobatto points, :506-531"""
    D = np.zeros((N, N))
    x = gauss_lobatto(N)
    for i in range(1, N - 1):
        for j in range(1, N - 1):
            if i != j:
                D[i, j] = ((-1) ** (j + 1) * (1. - x[j] ** 2) / ((1. - x[i] ** 2) * (x[i] - x[j])))
            else:
                D[i, i] = 3 * x[i] / (2. * (1. - x[i] ** 2))
    return D[1:-1, 1:-1]


def process_bcs(bcs):
    """:201-230 -- Dirichlet only; returns dict of alpha/beta/g per side."""
    out = {}
    for bc in bcs:
        typ, side, val = (bc.type, bc.boundary, bc.value) if hasattr(bc, "type") else (bc[1], bc[0], bc[2])
        if typ == 'dirichlet':
            key = {'left': 'minus_x', 'right': 'plus_x', 'top': 'minus_y', 'bottom': 'plus_y'}.get(side)
            if key is None:
                raise Exception('Boundary side {} not supported'.format(side))
            out['alpha_' + key] = 1
            out['g_' + key] = val
        elif typ == 'neumann':
            raise NotImplementedError
        else:
            raise Exception('Boundary type {} not supported'.format(typ))
    for k in ('minus_x', 'plus_x', 'minus_y', 'plus_y'):
        out['beta_' + k] = 0
    return out


def boundary_constants(D, am, ap, bm, bp):
    """get_boundary_constants, :102-118"""
    c0_minus = -bp * D[0, -1]
    c0_plus = am + bm * D[-1, -1]
    cN_plus = -bm * D[-1, 0]
    cN_minus = ap + bp * D[0, 0]
    e = c0_plus * cN_minus - c0_minus * cN_plus
    b0 = -c0_plus * bp * D[0, 1:-1] - c0_minus * bm * D[-1, 1:-1]
    bN = -cN_minus * bm * D[-1, 1:-1] - cN_plus * bp * D[0, 1:-1]
    return dict(e=e, c0_minus=c0_minus, c0_plus=c0_plus, cN_minus=cN_minus, cN_plus=cN_plus, b0=b0, bN=bN)


class Setup:
    """_pseudospectral_setup, :59-199"""

    def __init__(self, nx, ny, u_bc, v_bc):
        self.nx, self.ny = nx, ny
        self.Dx, self.Dy = D_matrix(nx), D_matrix(ny)
        self.Dx_sqr, self.Dy_sqr = D_sqr_matrix(nx), D_sqr_matrix(ny)
        self.bc = {'u': process_bcs(u_bc), 'v': process_bcs(v_bc)}
        self.k = {}
        for f in ('u', 'v'):
            b = self.bc[f]
            self.k[f + 'x'] = boundary_constants(self.Dx, b['alpha_minus_x'], b['alpha_plus_x'],
                                                 b['beta_minus_x'], b['beta_plus_x'])
            self.k[f + 'y'] = boundary_constants(self.Dy, b['alpha_minus_y'], b['alpha_plus_y'],
                                                 b['beta_minus_y'], b['beta_plus_y'])
        self.helm = {}
        for f in ('u', 'v'):
            kx, ky = self.k[f + 'x'], self.k[f + 'y']
            Mx = self.Dx_sqr[1:-1, 1:-1] + 1. / kx['e'] * (kx['b0'] * self.Dx_sqr[1:-1, 0] +
                                                           kx['bN'] * self.Dx_sqr[1:-1, -1])      # :159-166
            My = self.Dy_sqr[1:-1, 1:-1] + 1. / ky['e'] * (ky['b0'] * self.Dy_sqr[1:-1, 0] +
                                                           ky['bN'] * self.Dy_sqr[1:-1, -1])
            lx, P = np.linalg.eig(Mx)
            ly, Q = np.linalg.eig(My)
            self.helm[f] = dict(lx=lx, P=P, ly=ly, Q=Q, Pinv=np.linalg.inv(P), Qinv=np.linalg.inv(Q))
        self.DPx, self.DPy = D_pressure_matrix(nx), D_pressure_matrix(ny)
        self.DxDPx = self.Dx[1:-1, 1:-1] @ self.DPx
        self.DyDPy = self.Dy[1:-1, 1:-1] @ self.DPy
        lx, P = np.linalg.eig(self.DxDPx)
        ly, Q = np.linalg.eig(self.DyDPy)
        self.pres = dict(lx=lx, P=P, ly=ly, Q=Q, Pinv=np.linalg.inv(P), Qinv=np.linalg.inv(Q))

    def is_real(self):
        arrs = [self.pres[k] for k in ('lx', 'ly', 'P', 'Q')]
        for f in ('u', 'v'):
            arrs += [self.helm[f][k] for k in ('lx', 'ly', 'P', 'Q')]
        return not any(np.iscomplexobj(a) for a in arrs)


def predictor(S, dt, un, vn, un1, vn1):
    """_predictor_step, :232-337.  Returns (ui, vi)."""
    Nx, Ny = S.nx, S.ny
    Dx, Dy = S.Dx[1:-1, 1:-1], S.Dy[1:-1, 1:-1]
    Dx2, Dy2 = S.Dx_sqr[1:-1, 1:-1], S.Dy_sqr[1:-1, 1:-1]
    _un, _un1, _vn, _vn1 = un[1:-1, 1:-1], un1[1:-1, 1:-1], vn[1:-1, 1:-1], vn1[1:-1, 1:-1]
    un_dx, un_dy = Dx @ _un, _un @ Dy.T
    un1_dx, un1_dy = Dx @ _un1, _un1 @ Dy.T
    vn_dx, vn_dy = Dx @ _vn, _vn @ Dy.T
    vn1_dx, vn1_dy = Dx @ _vn1, _vn1 @ Dy.T
    un_ddx, un_ddy = Dx2 @ _un, _un @ Dy2.T
    vn_ddx, vn_ddy = Dx2 @ _vn, _vn @ Dy2.T
    F = {'u': 2 * _un - 3 * dt * (_un * un_dx + _vn * un_dy) + dt * (_un1 * un1_dx + _vn1 * un1_dy) +
              dt * (un_ddx + un_ddy),
         'v': 2 * _vn - 3 * dt * (_un * vn_dx + _vn * vn_dy) + dt * (_un1 * vn1_dx + _vn1 * vn1_dy) +
              dt * (vn_ddx + vn_ddy)}
    out = {}
    for f in ('u', 'v'):
        h = S.helm[f]
        Ht = h['Pinv'] @ F[f]
        Hh = Ht @ h['Qinv'].T
        hat = Hh / (2. - dt * h['lx'][:, None].repeat(Nx - 2, axis=1) - dt * h['ly'][None, :].repeat(Ny - 2, axis=0))
        til = hat @ h['Q'].T
        sol = h['P'] @ til
        kx, ky, b = S.k[f + 'x'], S.k[f + 'y'], S.bc[f]
        x0 = 1. / kx['e'] * np.sum(kx['b0'][:, None] * sol, axis=0) + \
            1. / kx['e'] * (kx['c0_minus'] * b['g_minus_x'] + kx['c0_plus'] * b['g_plus_x'])
        xN = 1. / kx['e'] * np.sum(kx['bN'][:, None] * sol, axis=0)
        y0 = 1. / ky['e'] * np.sum(ky['b0'][None, :] * sol, axis=1) + \
            1. / ky['e'] * (ky['c0_minus'] * b['g_minus_y'] + ky['c0_plus'] * b['g_plus_y'])
        yN = 1. / ky['e'] * np.sum(ky['bN'][None, :] * sol, axis=1)
        A = np.zeros((Nx, Ny))
        A[1:-1, 1:-1] = sol
        A[0, 1:-1], A[-1, 1:-1], A[1:-1, 0], A[1:-1, -1] = x0, xN, y0, yN
        out[f] = A
    return out['u'], out['v']


def pressure_rhs_S(S):
    """:353-361"""
    Nx, Ny = S.nx, S.ny
    bu, bv = S.bc['u'], S.bc['v']
    u_tau = np.stack([np.ones(Ny - 2) * bu['g_minus_x'], np.ones(Ny - 2) * bu['g_plus_x']])
    v_tau = np.stack([np.ones(Nx - 2) * bv['g_minus_y'], np.ones(Nx - 2) * bv['g_plus_y']]).T
    Dx_bar = np.stack([S.Dx[1:-1, 0], S.Dx[1:-1, -1]]).T
    Dy_bar = np.stack([S.Dy[1:-1, 0], S.Dy[1:-1, -1]]).T
    return -(Dx_bar @ u_tau + v_tau @ Dy_bar.T)


def correction(S, dt, rho, ui, vi, p):
    """_correction_step, :339-383.  Returns (u, v, p, Q)."""
    Nx, Ny = S.nx, S.ny
    Sm = pressure_rhs_S(S)
    H = -rho / dt * (Sm - S.Dx[1:-1, 1:-1] @ ui[1:-1, 1:-1] - vi[1:-1, 1:-1] @ S.Dy[1:-1, 1:-1].T)
    pr = S.pres
    Ht = pr['Pinv'] @ H
    Hh = Ht @ pr['Qinv'].T
    Qh = Hh / (pr['lx'][:, None].repeat(Nx - 2, axis=1) + pr['ly'][None, :].repeat(Ny - 2, axis=0))
    Qt = Qh @ pr['Q'].T
    Q = pr['P'] @ Qt
    u, v, pn = ui.copy(), vi.copy(), p.copy()
    u[1:-1, 1:-1] = u[1:-1, 1:-1] - S.DxDPx @ Q * dt / rho
    v[1:-1, 1:-1] = v[1:-1, 1:-1] - Q @ S.DyDPy.T * dt / rho
    pn[1:-1, 1:-1] = Q
    return u, v, pn, Q


def step(S, dt, rho, un, vn, un1, vn1, p):
    ui, vi = predictor(S, dt, un, vn, un1, vn1)
    u, v, pn, _ = correction(S, dt, rho, ui, vi, p)
    return u, v, pn
