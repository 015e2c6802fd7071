"""CPU: the C-ABI library loads and exports every symbol include/nns_b200.h declares; host-side
logic of the drop-in classes (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests._util import ROOT, has_gpu

import nns_b200
from nns_b200 import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nns_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nns_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_header_symbols():
    from nns_b200 import build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared_symbols()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(L, name), "missing export " + name
    assert sorted(_lib.exported_symbols()) == decl       # the ctypes prototypes cover the whole header
    assert _lib.lib().nns_abi_version() == 2


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.NnsBC) == 24
    assert ctypes.sizeof(_lib.NnsParams) == 6 * 4 + 5 * 8 + 2 * 4 + 8      # + force_x (ABI version 2)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(_lib.NnsError) as e:
        _lib.Handle(_lib.SOLVER_CHORIN_FD, 16, 16, 10, 1e-3, 1, 0.1)
    assert "no CUDA device" in str(e.value)
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    z = np.zeros((8, 8))
    s = NavierStokesSystem(z, z, z, [], [], [], nt=1, nx=8, ny=8, method='explicit')
    with pytest.raises(_lib.NnsError):
        s.simulate()


def test_boundary_api_and_asserts():
    D, N = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    bc = D(1, 'right', 0.05, 0.05)
    assert (bc.type, bc.boundary, bc.value, bc.dx, bc.dy) == ('dirichlet', 'right', 1, 0.05, 0.05)
    assert N(0, 'top', 0.1, 0.2).type == 'neumann'
    for bad in (lambda: D(0, 'front', 0.1, 0.1), lambda: D(0, 'left', 1, 0.1), lambda: N(0, 'left', 0.1, 2),
                lambda: N(0, 3, 0.1, 0.1)):
        with pytest.raises(AssertionError):       # boundary.py:16-19
            bad()
    with pytest.raises(NotImplementedError):
        nns_b200.BaseBoundaryCondition(0, 'left', 0.1, 0.1).apply(np.zeros((3, 3)))
    assert bc.abi_codes() == (1, 0) and N(0, 'top', 0.1, 0.2).abi_codes() == (3, 1)


def test_bc_table_keeps_list_order():
    D, N = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    u_bc = [D(0.5, 'top', .1, .1), N(2.0, 'left', .1, .1)]
    p_bc = [N(0.0, 'bottom', .1, .1)]
    arr, n = _lib.bc_table(u_bc, [], p_bc)
    assert n == 3
    assert [(arr[k].field, arr[k].side, arr[k].type, arr[k].value) for k in range(n)] == \
        [(0, 3, 0, 0.5), (0, 0, 1, 2.0), (2, 2, 1, 0.0)]


def test_constructor_contracts():
    from nns_b200.chorin_fd.simulate import NavierStokesSystem as C
    from nns_b200.direct_fd.simulate import NavierStokesSystem as Dd
    z = np.zeros((5, 7))
    with pytest.raises(AssertionError):            # chorin_fd/simulate.py:60
        C(z, z, z, [], [], [], nx=5, ny=7, method='implicit')
    s = C(z, z, z, [], [], [], nx=5, ny=7)
    assert s.method == 'semi_implicit' and s.nit == 50 and s.beta == 1.25 and s.nu == 1
    assert s.dx == 2. / 4 and s.dy == 2. / 6
    d = Dd(z, z, z, [], [], [], nx=5, ny=7)
    assert d.nu == 0.1 and d.nt == 200 and not hasattr(d, 'beta')
    u, v, p = s._init_variables()
    assert u is not z and np.array_equal(u, z)


def test_constants_names():
    from nns_b200 import constants
    assert constants.CHORIN_FD_DATA_FILE.endswith(os.path.join('data', 'chorin_fd', 'data_semi_implicit.npz'))
    assert constants.DIRECT_FD_DATA_FILE.endswith(os.path.join('data', 'direct_fd', 'data.npz'))


def test_member_sharding():
    from nns_b200.ensemble import cavity_ensemble_params, shard_members
    for B, W in ((4096, 8), (10, 4), (7, 8)):
        cover = []
        for r in range(W):
            lo, hi = shard_members(B, r, W)
            cover += list(range(lo, hi))
        assert cover == list(range(B))
    lid, nu = cavity_ensemble_params(4096)
    lid2, nu2 = cavity_ensemble_params(4096, lo=512, hi=1024)
    assert np.array_equal(lid[512:1024], lid2) and np.array_equal(nu[512:1024], nu2)
