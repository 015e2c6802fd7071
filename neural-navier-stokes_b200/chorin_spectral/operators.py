"""One-time HOST setup of the Chebyshev collocation operators of the reference's
``src/chorin_spectral/simulate.py`` (``_pseudospectral_setup``, :59-199, and the matrix builders
:395-531), delivered in the order ``nns_spectral_set_operators`` expects (include/nns_b200.h).

Why the host: the reference obtains its Helmholtz / Uzawa eigenbases from LAPACK (``np.linalg.eig`` /
``inv``) and some of the resulting denominators are ~1e-11, so a 1-ulp change of a matrix entry moves
the pressure by ~1e-5.  The entries are therefore produced with the same scalar libm calls and the
same operation order as the reference, and the decompositions with the same LAPACK calls; the
device then only differs from the reference in the GEMM summation order.  The time step itself (28
dense products per step) runs in libnns_b200.so.

Reference quirks kept (they are the spec): c-bar is 2 only for index 0; the sine formula divides
by 2N with N the number of points while the mesh uses N-1; D_sqr is D @ D.T with the diagonal
replaced by minus the full row sum; ``nu`` is unused.
"""
import numpy as np

_SIDE_KEY = {'left': 'minus_x', 'right': 'plus_x', 'top': 'minus_y', 'bottom': 'plus_y'}   # :203-215


def lobatto_points(n):
    """Gauss-Lobatto mesh cos(pi i / (n-1))  (:395-399)."""
    return np.cos(np.pi * np.arange(n) / float(n - 1))


def first_derivative(n):
    """Collocation derivative matrix, sine form, negative-row-sum diagonal (:443-481)."""
    D = np.zeros((n, n))
    for r in range(n):
        cr = 2 if r == 0 else 1
        for c in range(n):
            if c == r:
                continue
            cc = 2 if c == 0 else 1
            denom = 2 * np.sin((c + r) * np.pi / (2. * n)) * np.sin((c - r) * np.pi / (2. * n))
            D[r, c] = cr / cc * (-1) ** (r + c) / denom
    for r in range(n):
        D[r, r] = -np.sum(D[r, :])
    return D


def second_derivative(n):
    """The reference's D_sqr: D @ D.T, then diag := -(row sum)  (:483-504)."""
    D = first_derivative(n)
    S = np.array(D @ D.T)
    for r in range(n):
        S[r, r] = -np.sum(S[r, :])
    return S


def pressure_derivative(n):
    """P_{N-2} derivative on the interior points (:506-531); returns the (n-2, n-2) block."""
    x = lobatto_points(n)
    D = np.zeros((n, n))
    for r in range(1, n - 1):
        for c in range(1, n - 1):
            if r == c:
                D[r, r] = 3 * x[r] / (2. * (1. - x[r] ** 2))
            else:
                D[r, c] = ((-1) ** (c + 1) * (1. - x[c] ** 2) / ((1. - x[r] ** 2) * (x[r] - x[c])))
    return D[1:-1, 1:-1]


def read_bcs(bcs):
    """alpha / beta / g per side from a list of BC objects; Dirichlet only (:201-230)."""
    out = {}
    for bc in bcs:
        if bc.type == 'dirichlet':
            if bc.boundary not in _SIDE_KEY:
                raise Exception('Boundary side {} not supported'.format(bc.boundary))
            out['alpha_' + _SIDE_KEY[bc.boundary]] = 1
            out['g_' + _SIDE_KEY[bc.boundary]] = bc.value
        elif bc.type == 'neumann':
            raise NotImplementedError       # chorin_spectral:221
        else:
            raise Exception('Boundary type {} not supported'.format(bc.type))
    for key in _SIDE_KEY.values():
        out['beta_' + key] = 0
    return out


def _edge_constants(D, am, ap, bm, bp):
    """e, c0-, c0+, cN-, cN+, b0, bN of one direction (:102-118)."""
    c0m = -bp * D[0, -1]
    c0p = am + bm * D[-1, -1]
    cNp = -bm * D[-1, 0]
    cNm = ap + bp * D[0, 0]
    e = c0p * cNm - c0m * cNp
    b0 = -c0p * bp * D[0, 1:-1] - c0m * bm * D[-1, 1:-1]
    bN = -cNm * bm * D[-1, 1:-1] - cNp * bp * D[0, 1:-1]
    return e, c0m, c0p, cNm, cNp, b0, bN


class SpectralOperators:
    """All matrices of one (nx, ny, u_bc, v_bc) configuration."""

    def __init__(self, nx, ny, u_bc, v_bc):
        self.nx, self.ny = nx, ny
        self.Dx, self.Dy = first_derivative(nx), first_derivative(ny)
        self.Dx_sqr, self.Dy_sqr = second_derivative(nx), second_derivative(ny)
        self.bc = {'u': read_bcs(u_bc), 'v': read_bcs(v_bc)}
        self.edge, self.helm = {}, {}
        for f in ('u', 'v'):
            b = self.bc[f]
            ex = _edge_constants(self.Dx, b['alpha_minus_x'], b['alpha_plus_x'], b['beta_minus_x'], b['beta_plus_x'])
            ey = _edge_constants(self.Dy, b['alpha_minus_y'], b['alpha_plus_y'], b['beta_minus_y'], b['beta_plus_y'])
            self.edge[f] = (ex, ey)
            Mx = self.Dx_sqr[1:-1, 1:-1] + 1. / ex[0] * (ex[5] * self.Dx_sqr[1:-1, 0] + ex[6] * self.Dx_sqr[1:-1, -1])
            My = self.Dy_sqr[1:-1, 1:-1] + 1. / ey[0] * (ey[5] * self.Dy_sqr[1:-1, 0] + ey[6] * self.Dy_sqr[1:-1, -1])
            lx, P = np.linalg.eig(Mx)
            ly, Q = np.linalg.eig(My)
            self.helm[f] = dict(lx=lx, ly=ly, P=P, Q=Q, Pinv=np.linalg.inv(P), Qinv=np.linalg.inv(Q))
        self.DPx, self.DPy = pressure_derivative(nx), pressure_derivative(ny)
        self.DxDPx = self.Dx[1:-1, 1:-1] @ self.DPx
        self.DyDPy = self.Dy[1:-1, 1:-1] @ self.DPy
        lx, P = np.linalg.eig(self.DxDPx)
        ly, Q = np.linalg.eig(self.DyDPy)
        self.pres = dict(lx=lx, ly=ly, P=P, Q=Q, Pinv=np.linalg.inv(P), Qinv=np.linalg.inv(Q))

    def is_real(self):
        """False when LAPACK returned complex pairs (even N >= 64): the reference then dies at its first
        step with ComplexWarning promoted to an error (chorin_spectral:3, :379)."""
        parts = [self.pres] + [self.helm[f] for f in ('u', 'v')]
        return not any(np.iscomplexobj(d[k]) for d in parts for k in ('lx', 'ly', 'P', 'Q'))

    def boundary_source(self):
        """S of the Uzawa right-hand side, from the Dirichlet data (:353-361); shape (nx-2, ny-2)."""
        bu, bv = self.bc['u'], self.bc['v']
        u_tau = np.stack([np.ones(self.ny - 2) * bu['g_minus_x'], np.ones(self.ny - 2) * bu['g_plus_x']])
        v_tau = np.stack([np.ones(self.nx - 2) * bv['g_minus_y'], np.ones(self.nx - 2) * bv['g_plus_y']]).T
        Dx_bar = np.stack([self.Dx[1:-1, 0], self.Dx[1:-1, -1]]).T
        Dy_bar = np.stack([self.Dy[1:-1, 0], self.Dy[1:-1, -1]]).T
        return -(Dx_bar @ u_tau + v_tau @ Dy_bar.T)

    def abi_arrays(self):
        """The 29 float64 arrays of nns_spectral_set_operators, in ABI order (include/nns_b200.h)."""
        if not self.is_real():
            raise ValueError("complex eigen-decomposition")
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        out = [self.Dx[1:-1, 1:-1], self.Dy[1:-1, 1:-1], self.Dx_sqr[1:-1, 1:-1], self.Dy_sqr[1:-1, 1:-1]]
        for f in ('u', 'v'):
            h = self.helm[f]
            out += [h['Pinv'], h['Qinv'], h['P'], h['Q']]
        for f in ('u', 'v'):
            out += [self.helm[f]['lx'], self.helm[f]['ly']]
        pr = self.pres
        out += [pr['Pinv'], pr['Qinv'], pr['P'], pr['Q'], pr['lx'], pr['ly'], self.DxDPx, self.DyDPy,
                self.boundary_source()]
        scal = []
        for f in ('u', 'v'):
            ex, ey = self.edge[f]
            b = self.bc[f]
            out.append(np.concatenate([ex[5], ex[6], ey[5], ey[6]]))
            # row 0 / column 0 constants (:322-334): 1/e * (c0- g- + c0+ g+)
            scal.append(np.array([1. / ex[0], 1. / ex[0] * (ex[1] * b['g_minus_x'] + ex[2] * b['g_plus_x']),
                                  1. / ey[0], 1. / ey[0] * (ey[1] * b['g_minus_y'] + ey[2] * b['g_plus_y'])]))
        out += scal
        return [c(a) for a in out]
