"""GPU parity: chorin_fd CUDA path (through the C ABI) vs the oracle / the reference fixtures.
Tolerance: fp64 relative L2 <= 1e-10 on u, v, p (BASELINE.json north_star); sweep counts exact."""
import json

import numpy as np
import pytest

from tests._util import load_golden, make_bcs, rel_l2, smooth_ic

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True, params=["auto", "split"])
def chip_mode(request, monkeypatch):
    """Every test runs twice: with the planner's choice (register-block SOR where a template
    instance fits) and with the generic split-row shared-memory SOR forced."""
    if request.param == "auto":
        monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    else:
        monkeypatch.setenv("NNS_CHIP_MODE", request.param)
    return request.param


def _params(g, key):
    return json.loads(str(g[key]))


def _system(P, u0, v0, p0):
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    nx, ny = P["nx"], P["ny"]
    return NavierStokesSystem(u0, v0, p0, make_bcs(P["u_bc"], nx, ny), make_bcs(P["v_bc"], nx, ny),
                              make_bcs(P["p_bc"], nx, ny), nt=P["nt"], nit=P["nit"], nx=nx, ny=ny, dt=P["dt"],
                              rho=P["rho"], nu=P["nu"], beta=P["beta"], method=P["method"])


def _worst(a, b):
    return max(rel_l2(a[n], b[n]) for n in range(len(a)))


@pytest.mark.parametrize("case", ["c0", "c1", "c2", "c3"])
def test_mixed_bcs_nonsquare_vs_reference(case):
    g = load_golden("chorin_mixed")
    P = _params(g, case + "_params")
    u0, v0, p0 = g[case + "_u0"].copy(), g[case + "_v0"].copy(), g[case + "_p0"].copy()
    s = _system(P, u0, v0, p0)
    u, v, p = s.simulate()
    assert u.shape == (P["nt"], P["nx"], P["ny"]) and u.dtype == np.float64
    assert _worst(u, g[case + "_u"]) <= TOL
    assert _worst(v, g[case + "_v"]) <= TOL
    assert _worst(p, g[case + "_p"]) <= TOL
    assert np.array_equal(s.last_sweeps, g[case + "_sweeps"])      # incl. the early exits of c3
    assert np.array_equal(u0, g[case + "_u0"])                     # ICs are not mutated (chorin_fd:238)


def test_cavity41_config1_full_trajectory():
    g = load_golden("chorin_cavity41")
    P = _params(g, "params")
    z = np.zeros((P["nx"], P["ny"]))
    s = _system(P, z, z.copy(), z.copy())
    u, v, p = s.simulate()
    fr = g["frames"]
    for a, name in ((u, "u"), (v, "v"), (p, "p")):
        assert _worst(a[fr], g[name]) <= TOL, name
    norms = np.stack([[np.linalg.norm(a[n].ravel()) for n in range(P["nt"])] for a in (u, v, p)])
    assert np.max(np.abs(norms - g["norms"]) / np.maximum(g["norms"], 1e-300)) <= TOL
    assert np.array_equal(s.last_sweeps, g["sweeps"])
    assert s.last_sweeps.min() == 36


@pytest.mark.parametrize("case", ["s0", "s1"])
def test_semi_implicit_vs_reference(case):
    g = load_golden("chorin_semi")
    P = _params(g, case + "_params")
    s = _system(P, g[case + "_u0"].copy(), g[case + "_v0"].copy(), g[case + "_p0"].copy())
    u, v, p = s.simulate()
    assert _worst(u, g[case + "_u"]) <= TOL
    assert _worst(v, g[case + "_v"]) <= TOL
    assert _worst(p, g[case + "_p"]) <= TOL
    assert np.array_equal(s.last_sweeps, g[case + "_sweeps"])


def test_step_api_matches_oracle_and_mutates_p(oracle_fd):
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    nx, ny = 19, 26
    spec_u = [["left", "neumann", 0.2], ["top", "dirichlet", -0.3], ["right", "dirichlet", 1.0], ["bottom", "neumann", 0.0]]
    spec_v = [["bottom", "dirichlet", 0.1], ["left", "dirichlet", 0.0], ["top", "neumann", 0.4]]
    spec_p = [["top", "dirichlet", 0.0], ["bottom", "neumann", 0.1], ["left", "neumann", 0.0], ["right", "neumann", -0.2]]
    un, vn, p = smooth_ic(nx, ny, 7, amp=0.2)
    un1, vn1, _ = smooth_ic(nx, ny, 8, amp=0.2)
    kw = dict(nit=30, dt=4e-4, rho=1.1, nu=0.05, beta=1.3)
    s = NavierStokesSystem(un, vn, p, make_bcs(spec_u, nx, ny), make_bcs(spec_v, nx, ny), make_bcs(spec_p, nx, ny),
                           nx=nx, ny=ny, method='explicit', **kw)
    p_gpu, p_cpu = p.copy(), p.copy()
    u2, v2, p2 = s.step(un, vn, un1, vn1, p_gpu)
    ou, ov, op, sw = oracle_fd.chorin_step(un, vn, un1, vn1, p_cpu, [tuple(b) for b in spec_u],
                                           [tuple(b) for b in spec_v], [tuple(b) for b in spec_p],
                                           method='explicit', **kw)
    assert p2 is p_gpu                                   # same object, updated in place (chorin_fd:193,227)
    assert rel_l2(u2, ou) <= TOL and rel_l2(v2, ov) <= TOL and rel_l2(p_gpu, op) <= TOL
    assert int(s.last_sweeps[0]) == sw


def test_stage_entry_points_vs_oracle(oracle_fd):
    """predictor / pressure / correct separately, batch of 3 with per-member nu and BC values."""
    import torch
    from nns_b200.ensemble import ChorinEnsemble
    from oracle.fd import lib as olib, _dp, bc_array
    import ctypes as C
    B, nx, ny = 3, 22, 15
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    spec_u = [["left", "dirichlet", 0.0], ["right", "dirichlet", 1.0], ["top", "neumann", 0.1], ["bottom", "dirichlet", 0.0]]
    spec_v = [["left", "neumann", 0.0], ["right", "dirichlet", 0.0], ["top", "dirichlet", 0.2], ["bottom", "dirichlet", 0.0]]
    spec_p = [["top", "dirichlet", 0.0], ["bottom", "neumann", 0.0], ["left", "neumann", 0.3], ["right", "neumann", 0.0]]
    rng = np.random.default_rng(2)
    vals = np.tile(np.array([b[2] for b in spec_u + spec_v + spec_p], dtype=np.float64), (B, 1))
    vals[:, 1] = rng.uniform(0.5, 1.5, B)
    nus = rng.uniform(0.02, 0.1, B)
    kw = dict(nit=25, dt=5e-4, rho=1.0, beta=1.25)
    ens = ChorinEnsemble(B, nx, ny, u_bc=make_bcs(spec_u, nx, ny), v_bc=make_bcs(spec_v, nx, ny),
                         p_bc=make_bcs(spec_p, nx, ny), nu=nus, bc_values=vals, method='explicit', **kw)
    f = [np.stack([smooth_ic(nx, ny, 100 + 10 * b + k, amp=0.2)[0] for b in range(B)]) for k in range(5)]
    dev = [torch.from_numpy(a).cuda() for a in f]
    ui, vi = ens.predictor(*dev[:4])
    pd = dev[4].clone()
    ens.pressure(ui, vi, pd)
    p_after_sor = pd.clone()
    uo, vo, pf = ens.correct(ui, vi, pd)
    torch.cuda.synchronize()
    for b in range(B):
        bu = [(s, t, vals[b, k]) for k, (s, t, _) in enumerate(spec_u)]
        bv = [(s, t, vals[b, 4 + k]) for k, (s, t, _) in enumerate(spec_v)]
        bp = [(s, t, vals[b, 8 + k]) for k, (s, t, _) in enumerate(spec_p)]
        eui, evi = np.empty((nx, ny)), np.empty((nx, ny))
        olib().orc_chorin_explicit_predictor(_dp(f[0][b]), _dp(f[1][b]), _dp(f[2][b]), _dp(f[3][b]), _dp(eui),
                                             _dp(evi), nx, ny, C.c_double(kw["dt"]), C.c_double(dx),
                                             C.c_double(dy), C.c_double(nus[b]))
        for bc in bu: oracle_fd.bc_apply(eui, bc, dx, dy)
        for bc in bv: oracle_fd.bc_apply(evi, bc, dx, dy)
        assert rel_l2(ui[b].cpu().numpy(), eui) <= TOL and rel_l2(vi[b].cpu().numpy(), evi) <= TOL
        ep = f[4][b].copy()
        olib().orc_chorin_pressure(_dp(eui), _dp(evi), _dp(ep), nx, ny, kw["nit"], C.c_double(kw["dt"]),
                                   C.c_double(dx), C.c_double(dy), C.c_double(kw["rho"]), C.c_double(kw["beta"]))
        assert rel_l2(p_after_sor[b].cpu().numpy(), ep) <= TOL
        for bc in bp: oracle_fd.bc_apply(ep, bc, dx, dy)
        eu, ev = np.empty((nx, ny)), np.empty((nx, ny))
        olib().orc_chorin_correction(_dp(eui), _dp(evi), _dp(ep), _dp(eu), _dp(ev), nx, ny, C.c_double(kw["dt"]),
                                     C.c_double(dx), C.c_double(dy))
        assert rel_l2(uo[b].cpu().numpy(), eu) <= TOL and rel_l2(vo[b].cpu().numpy(), ev) <= TOL
        assert rel_l2(pf[b].cpu().numpy(), ep) <= TOL


def test_ensemble128_members_vs_reference():
    """BASELINE config 4 members (lid, Re drawn with default_rng(0)) vs the reference fixture,
    run inside a larger batch so that member placement in the grid does not matter."""
    import torch
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values, cavity_bcs, cavity_ensemble_params
    g = load_golden("chorin_ens128")
    P = _params(g, "params")
    nx = ny = P["nx"]
    lid_all, nu_all = cavity_ensemble_params(4096, seed=0)
    members = [int(m) for m in g["members"]]
    assert np.allclose(lid_all[members], g["lid"]) and np.allclose(nu_all[members], g["nu"])
    pick = members + [7, 300, 2048]                      # padding members: placement independence
    lid, nu = lid_all[pick], nu_all[pick]
    dx = dy = 2. / (nx - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = ChorinEnsemble(len(pick), nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=P["nit"], dt=P["dt"],
                         rho=P["rho"], nu=nu, beta=P["beta"], method='explicit', bc_values=cavity_bc_values(lid))
    ens.init_variables()
    sw = []
    for _ in range(P["nt"]):
        ens.step()
        sw.append(ens.sweeps.cpu().numpy().copy())
    sw = np.stack(sw)
    for k, b in enumerate(members):
        assert rel_l2(ens.u[k].cpu().numpy(), g["m%d_u" % b]) <= TOL
        assert rel_l2(ens.v[k].cpu().numpy(), g["m%d_v" % b]) <= TOL
        assert rel_l2(ens.p[k].cpu().numpy(), g["m%d_p" % b]) <= TOL
        assert np.array_equal(sw[:, k], g["m%d_sweeps" % b])


def test_run_equals_repeated_step_and_trajectory_layout():
    """nns_chorin_fd_run (in-kernel rotation + fix-up) == nsteps x nns_chorin_fd_step, bit for bit,
    for nsteps % 3 in {0, 1, 2}; trajectory is [B, nsteps, nx, ny] and its last frame is the state."""
    import torch
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values, cavity_bcs
    B, nx, ny = 5, 33, 40
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    lid = np.linspace(0.5, 1.5, B)
    mk = lambda: ChorinEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=20, dt=1e-3, nu=0.05,
                                bc_values=cavity_bc_values(lid))
    for nsteps in (3, 4, 5):
        a, b = mk(), mk()
        a.init_variables(); b.init_variables()
        for _ in range(nsteps):
            a.step()
        tu, tv, tp, sw = b.run(nsteps, trajectory=True, sweeps=True)
        for x, y in ((a.u, b.u), (a.v, b.v), (a.u1, b.u1), (a.v1, b.v1), (a.p, b.p)):
            assert torch.equal(x, y)
        assert tu.shape == (B, nsteps, nx, ny)
        assert torch.equal(tu[:, -1], b.u) and torch.equal(tv[:, -1], b.v) and torch.equal(tp[:, -1], b.p)
        assert torch.equal(tu[:, -2], b.u1)
        assert sw.shape == (nsteps, B) and int(sw.max()) <= 19


def test_member_result_independent_of_batch_and_order():
    """Full-size property (config 4 shape): a member computed inside a 296-member launch equals
    the same member computed alone -- bit for bit (no cross-member state, deterministic order)."""
    import torch
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values, cavity_bcs, cavity_ensemble_params
    nx = ny = 128
    B = 296
    lid, nu = cavity_ensemble_params(B, seed=1)
    dx = dy = 2. / (nx - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    big = ChorinEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=2e-4, nu=nu,
                         bc_values=cavity_bc_values(lid))
    big.init_variables()
    for _ in range(3):
        big.step()
    for b in (0, 151, 295):
        one = ChorinEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=2e-4, nu=nu[b:b + 1],
                             bc_values=cavity_bc_values(lid[b:b + 1]))
        one.init_variables()
        for _ in range(3):
            one.step()
        assert torch.equal(one.u[0], big.u[b]) and torch.equal(one.p[0], big.p[b])
    assert torch.isfinite(big.u).all() and int(big.sweeps.min()) == 49


def test_nit_edge_cases(oracle_fd):
    """nit = 1 (zero sweeps), nit = 2 (one sweep), nit = 80 (> 64 sweeps: flag array path)."""
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    nx, ny = 14, 11
    spec = dict(u=[["left", "dirichlet", 0.0], ["right", "dirichlet", 0.3], ["top", "dirichlet", 0.0], ["bottom", "dirichlet", 0.0]],
                v=[[s, "dirichlet", 0.0] for s in ("left", "right", "top", "bottom")],
                p=[["top", "dirichlet", 0.0], ["bottom", "neumann", 0.0], ["left", "neumann", 0.0], ["right", "neumann", 0.0]])
    z = np.zeros((nx, ny))
    for nit in (1, 2, 80, 200):
        s = NavierStokesSystem(z, z.copy(), z.copy(), make_bcs(spec["u"], nx, ny), make_bcs(spec["v"], nx, ny),
                               make_bcs(spec["p"], nx, ny), nt=25, nit=nit, nx=nx, ny=ny, dt=1e-3, rho=1, nu=0.1,
                               beta=1.25, method='explicit')
        u, v, p = s.simulate()
        ou, ov, op, sw = oracle_fd.chorin_simulate(z, z, z, [tuple(b) for b in spec["u"]], [tuple(b) for b in spec["v"]],
                                                   [tuple(b) for b in spec["p"]], nt=25, nit=nit, dt=1e-3, rho=1,
                                                   nu=0.1, beta=1.25)
        assert np.array_equal(s.last_sweeps, sw), nit
        assert _worst(u, ou) <= TOL and _worst(v, ov) <= TOL and _worst(p, op) <= TOL, nit


def test_nonfinite_raises_like_the_reference():
    from nns_b200 import _lib
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    nx = ny = 12
    z = np.zeros((nx, ny))
    big = 1e200 * np.random.default_rng(0).uniform(1.0, 2.0, size=(nx, ny))
    s = NavierStokesSystem(big, big.copy(), z, [], [], [], nt=3, nit=5, nx=nx, ny=ny, dt=1.0, rho=1, nu=1e150,
                           method='explicit')
    with pytest.raises(_lib.NnsError) as e:
        s.simulate()
    assert e.value.code == -4


def test_step_host_pipelined_equals_device_step():
    """nns_chorin_fd_step_host (chunked H2D || kernel || D2H pipeline over internal streams) ==
    the device-resident step, bit for bit, for a batch that is cut into 2 chunks."""
    import torch
    from nns_b200 import _lib
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values, cavity_bcs, cavity_ensemble_params
    B, nx, ny = 301, 41, 37
    lid, nu = cavity_ensemble_params(B, seed=3)
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = ChorinEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=30, dt=5e-4, nu=nu,
                         bc_values=cavity_bc_values(lid))
    ens.init_variables()
    ens.step(); ens.step()
    host = [t.cpu().contiguous().pin_memory() for t in (ens.u, ens.v, ens.u1, ens.v1, ens.p)]
    uo, vo = torch.empty_like(host[0]).pin_memory(), torch.empty_like(host[0]).pin_memory()
    sw = torch.zeros((B,), dtype=torch.int32).pin_memory()
    _lib.check(_lib.lib().nns_chorin_fd_step_host(ens.handle.h, *[h.data_ptr() for h in host], uo.data_ptr(),
                                                  vo.data_ptr(), sw.data_ptr()))
    ens.step()
    assert torch.equal(uo, ens.u.cpu()) and torch.equal(vo, ens.v.cpu()) and torch.equal(host[4], ens.p.cpu())
    assert torch.equal(sw, ens.sweeps.cpu())
