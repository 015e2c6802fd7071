"""CPU: the oracle restatement (oracle/oracle.c) against the fixtures recorded from the
reference itself (tests/golden/make_golden.py).  This is what pins the oracle."""
import json

import numpy as np
import pytest

from tests._util import load_golden, make_bcs, manifest, rel_l2


def _params(g, key):
    return json.loads(str(g[key]))


def test_manifest_records_bit_exact_pin():
    m = manifest()
    assert m["boundary"]["pin"]["bit_exact"]
    assert m["chorin_cavity41"]["pin"]["bit_exact"]
    for k, v in m["chorin_mixed"].items():
        assert v["pin"]["bit_exact"], k
    for k, v in m["direct_fd"].items():
        assert v["pin"]["bit_exact"], k
    for k, v in m["chorin_semi"].items():          # LAPACK dgesv vs plain LU: rounding only
        assert max(v["pin"]["rel_l2"]) < 1e-13, k


def test_boundary_sequence(oracle_fd):
    g = load_golden("boundary")
    specs = json.loads(str(g["specs"]))
    A = g["A0"].copy()
    nx, ny = A.shape
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    for k, (typ, side, val) in enumerate(specs):
        oracle_fd.bc_apply(A, (side, typ, val), dx, dy)
        assert np.array_equal(A, g["seq"][k]), (k, typ, side)


def test_package_boundary_objects_match_reference_fixture():
    import nns_b200
    g = load_golden("boundary")
    specs = json.loads(str(g["specs"]))
    A = g["A0"].copy()
    nx, ny = A.shape
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    for k, (typ, side, val) in enumerate(specs):
        cls = nns_b200.DirichletBoundaryCondition if typ == "dirichlet" else nns_b200.NeumannBoundaryCondition
        out = cls(val, side, dx, dy).apply(A)
        assert out is A
        assert np.array_equal(A, g["seq"][k]), (k, typ, side)


@pytest.mark.parametrize("case", ["c0", "c1", "c2", "c3"])
def test_chorin_mixed_bit_exact(oracle_fd, case):
    g = load_golden("chorin_mixed")
    P = _params(g, case + "_params")
    u, v, p, sw = oracle_fd.chorin_simulate(
        g[case + "_u0"], g[case + "_v0"], g[case + "_p0"],
        [tuple(b) for b in P["u_bc"]], [tuple(b) for b in P["v_bc"]], [tuple(b) for b in P["p_bc"]],
        nt=P["nt"], nit=P["nit"], dt=P["dt"], rho=P["rho"], nu=P["nu"], beta=P["beta"], method=P["method"])
    assert np.array_equal(u, g[case + "_u"])
    assert np.array_equal(v, g[case + "_v"])
    assert np.array_equal(p, g[case + "_p"])
    assert np.array_equal(sw, g[case + "_sweeps"])
    assert sw.max() <= P["nit"] - 1


def test_chorin_mixed_has_early_exit_case():
    g = load_golden("chorin_mixed")
    sw = g["c3_sweeps"]
    assert sw.min() < 49 and sw.max() == 49      # both branches of the while loop are pinned


@pytest.mark.parametrize("case", ["s0", "s1"])
def test_chorin_semi_implicit(oracle_fd, case):
    g = load_golden("chorin_semi")
    P = _params(g, case + "_params")
    u, v, p, sw = oracle_fd.chorin_simulate(
        g[case + "_u0"], g[case + "_v0"], g[case + "_p0"],
        [tuple(b) for b in P["u_bc"]], [tuple(b) for b in P["v_bc"]], [tuple(b) for b in P["p_bc"]],
        nt=P["nt"], nit=P["nit"], dt=P["dt"], rho=P["rho"], nu=P["nu"], beta=P["beta"], method=P["method"])
    for a, name in ((u, "_u"), (v, "_v"), (p, "_p")):
        assert rel_l2(a, g[case + name]) < 1e-12
    assert np.array_equal(sw, g[case + "_sweeps"])


def test_chorin_cavity41_full_run(oracle_fd):
    """BASELINE config 1 (500 steps) -- known answers from SURVEY.md 8c and the fixture."""
    g = load_golden("chorin_cavity41")
    P = _params(g, "params")
    z = np.zeros((P["nx"], P["ny"]))
    u, v, p, sw = oracle_fd.chorin_simulate(
        z, z, z, [tuple(b) for b in P["u_bc"]], [tuple(b) for b in P["v_bc"]], [tuple(b) for b in P["p_bc"]],
        nt=P["nt"], nit=P["nit"], dt=P["dt"], rho=P["rho"], nu=P["nu"], beta=P["beta"], method=P["method"])
    fr = g["frames"]
    assert np.array_equal(u[fr], g["u"]) and np.array_equal(v[fr], g["v"]) and np.array_equal(p[fr], g["p"])
    norms = np.stack([[np.linalg.norm(a[n].ravel()) for n in range(P["nt"])] for a in (u, v, p)])
    assert np.array_equal(norms, g["norms"])
    # the surveyor's independent probe of the reference
    assert abs(norms[0][0] - 6.250864652313306) < 1e-12
    assert abs(norms[2][499] - 159.6216806443248) < 1e-9
    assert sw.min() == 36 and sw.max() == 49


def test_chorin_ens128_members(oracle_fd):
    g = load_golden("chorin_ens128")
    P = _params(g, "params")
    z = np.zeros((P["nx"], P["ny"]))
    dx = dy = 2. / (P["nx"] - 1)
    for k, b in enumerate(g["members"][:2]):
        lid, nu = float(g["lid"][k]), float(g["nu"][k])
        u_bc = [("left", "dirichlet", 0), ("right", "dirichlet", lid), ("top", "dirichlet", 0), ("bottom", "dirichlet", 0)]
        v_bc = [(s, "dirichlet", 0) for s in ("left", "right", "top", "bottom")]
        p_bc = [("top", "dirichlet", 0), ("bottom", "neumann", 0), ("left", "neumann", 0), ("right", "neumann", 0)]
        u, v, p, sw = oracle_fd.chorin_simulate(z, z, z, u_bc, v_bc, p_bc, nt=P["nt"], nit=P["nit"], dt=P["dt"],
                                                rho=P["rho"], nu=nu, beta=P["beta"], method="explicit")
        assert np.array_equal(u[-1], g["m%d_u" % b]) and np.array_equal(p[-1], g["m%d_p" % b])


@pytest.mark.parametrize("case", ["d0", "d1", "d2"])
def test_direct_fd_bit_exact(oracle_fd, case):
    g = load_golden("direct_fd")
    P = _params(g, case + "_params")
    nx, ny = P["nx"], P["ny"]
    if case + "_u0" in g.files:
        u0, v0, p0 = g[case + "_u0"], g[case + "_v0"], g[case + "_p0"]
    else:
        u0 = v0 = p0 = np.zeros((nx, ny))
    u, v, p = oracle_fd.direct_simulate(u0, v0, p0, [tuple(b) for b in P["u_bc"]], [tuple(b) for b in P["v_bc"]],
                                        [tuple(b) for b in P["p_bc"]], nt=P["nt"], nit=P["nit"], dt=P["dt"],
                                        rho=P["rho"], nu=P["nu"])
    fr = g[case + "_frames"]
    assert np.array_equal(u[fr], g[case + "_u"])
    assert np.array_equal(v[fr], g[case + "_v"])
    assert np.array_equal(p[fr], g[case + "_p"])
    if case == "d0":     # SURVEY.md 8c known answers for the module's own demo config
        n = g[case + "_norms"]
        assert abs(n[0][0] - 6.928203230275509) < 1e-12 and abs(n[2][199] - 28.81095091598851) < 1e-10


def test_oracle_ensemble_equals_members(oracle_fd):
    """Ensemble driver == per-member simulate (it is the CPU baseline the bench times)."""
    B, nx, ny, nt = 3, 12, 10, 4
    rng = np.random.default_rng(3)
    lids = rng.uniform(0.5, 1.5, B)
    nus = rng.uniform(0.01, 0.1, B)
    mk = lambda lid: ([("left", "dirichlet", 0), ("right", "dirichlet", lid), ("top", "dirichlet", 0),
                       ("bottom", "dirichlet", 0)],
                      [(s, "dirichlet", 0) for s in ("left", "right", "top", "bottom")],
                      [("top", "dirichlet", 0), ("bottom", "neumann", 0), ("left", "neumann", 0),
                       ("right", "neumann", 0)])
    u = np.zeros((B, nx, ny)); v = np.zeros_like(u); p = np.zeros_like(u)
    for b in range(B):
        ub, vb, pb = mk(lids[b])
        for bc in ub: oracle_fd.bc_apply(u[b], bc, 2. / (nx - 1), 2. / (ny - 1))
    u1, v1 = u.copy(), v.copy()
    sw, _ = oracle_fd.chorin_ensemble_run(u, v, u1, v1, p, [mk(l)[0] for l in lids], [mk(l)[1] for l in lids],
                                          [mk(l)[2] for l in lids], nt=nt, nit=20, dt=1e-3, rho=1, nu=nus, beta=1.25)
    for b in range(B):
        ub, vb, pb = mk(lids[b])
        z = np.zeros((nx, ny))
        tu, tv, tp, s1 = oracle_fd.chorin_simulate(z, z, z, ub, vb, pb, nt=nt, nit=20, dt=1e-3, rho=1,
                                                   nu=nus[b], beta=1.25)
        assert np.array_equal(tu[-1], u[b]) and np.array_equal(tp[-1], p[b]) and np.array_equal(tu[-2], u1[b])
        assert np.array_equal(s1, sw[:, b])
