"""CPU: the chorin_spectral oracle (oracle/spectral.py) against the reference fixtures, and the product's
host-side operator setup against the oracle (bit for bit: both make the same libm / LAPACK calls)."""
import numpy as np
import pytest

from tests._util import load_golden, manifest, rel_l2

import nns_b200
from nns_b200.chorin_spectral.operators import SpectralOperators
from oracle import spectral as osp


def _bcs(N):
    D = nns_b200.DirichletBoundaryCondition
    dx = dy = 2. / (N - 1.)
    u_bc = [D(0, 'left', dx, dy), D(1, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
    v_bc = [D(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
    return u_bc, v_bc


def test_manifest_says_oracle_matches_reference_per_operator():
    """Recorded when the fixtures were generated from the reference's own class (make_golden_spectral.py)."""
    pins = manifest()["spectral"]
    for N in ("N21", "N51", "N127"):
        assert max(pins[N]["setup_max_abs"].values()) == 0.0
        for st in ("cav", "rnd"):
            assert pins[N][st]["ui"] <= 1e-12 and pins[N][st]["vi"] <= 1e-12 and pins[N][st]["Q"] <= 1e-12


@pytest.mark.parametrize("N", [21, 51])
def test_oracle_vs_reference_fixtures(N):
    g = load_golden("spectral")
    u_bc, v_bc = _bcs(N)
    S = osp.Setup(N, N, u_bc, v_bc)
    assert S.is_real()
    assert np.array_equal(S.Dx[1], g["N%d_Dx_row1" % N])
    assert np.array_equal(np.diag(S.Dx_sqr), g["N%d_Dx_sqr_diag" % N])
    assert np.array_equal(np.diag(S.DxDPx), g["N%d_DxDPx_diag" % N])
    assert np.allclose(np.sort(S.pres['lx']), g["N%d_p_lambda_sorted" % N], rtol=1e-9, atol=0)
    st = [g["N%d_rnd_%s" % (N, k)] for k in ("un", "vn", "un1", "vn1", "p")]
    ui, vi = osp.predictor(S, 1e-3, *st[:4])
    assert rel_l2(ui, g["N%d_rnd_ui" % N]) <= 1e-12 and rel_l2(vi, g["N%d_rnd_vi" % N]) <= 1e-12
    _, _, _, Q = osp.correction(S, 1e-3, 1, g["N%d_rnd_ui" % N], g["N%d_rnd_vi" % N], st[4])
    assert rel_l2(Q, g["N%d_rnd_Q" % N]) <= 1e-10


@pytest.mark.parametrize("N", [21, 51])
def test_product_operators_equal_oracle(N):
    u_bc, v_bc = _bcs(N)
    P, O = SpectralOperators(N, N, u_bc, v_bc), osp.Setup(N, N, u_bc, v_bc)
    for a, b in ((P.Dx, O.Dx), (P.Dx_sqr, O.Dx_sqr), (P.DPx, O.DPx), (P.DxDPx, O.DxDPx), (P.pres['P'], O.pres['P']),
                 (P.pres['lx'], O.pres['lx']), (P.helm['u']['Pinv'], O.helm['u']['Pinv']),
                 (P.helm['v']['ly'], O.helm['v']['ly']), (P.boundary_source(), osp.pressure_rhs_S(O))):
        assert np.array_equal(a, b)
    arrs = P.abi_arrays()
    assert len(arrs) == nns_b200._lib.SPECTRAL_N_OPERATORS
    n = N - 2
    assert arrs[0].shape == (n, n) and arrs[24].shape == (n, n) and arrs[25].shape == (4 * n,) and arrs[27].shape == (4,)
    assert all(a.dtype == np.float64 and a.flags.c_contiguous for a in arrs)


def test_reference_error_conventions():
    from nns_b200.chorin_spectral.simulate import NavierStokesSystem
    N = 12
    dx = dy = 2. / (N - 1.)
    u_bc, v_bc = _bcs(N)
    z = np.zeros((N, N))
    with pytest.raises(NotImplementedError):          # chorin_spectral:221
        NavierStokesSystem(z, z, z, [nns_b200.NeumannBoundaryCondition(0, 'left', dx, dy)] + u_bc[1:], v_bc, nx=N, ny=N)
    s = NavierStokesSystem(z, z, z, u_bc, v_bc, nt=2, nx=N, ny=N)
    assert (s.dx, s.dy) == (2. / N, 2. / N)           # chorin_spectral:48
    u, v, p = s._init_variables()
    assert u[-1, 1:-1].min() == 1.0 and u[-1, 0] == 0.0 and np.all(v == 0)


def test_even_n_complex_spectrum_is_detected():
    """The reference dies at its first step for even N >= 64 (complex eigenpairs)."""
    u_bc, v_bc = _bcs(64)
    assert not SpectralOperators(64, 64, u_bc, v_bc).is_real()
