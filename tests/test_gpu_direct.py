"""GPU parity: direct_fd CUDA path vs the reference fixtures / the oracle (rel-L2 <= 1e-10)."""
import json

import numpy as np
import pytest

from tests._util import load_golden, make_bcs, rel_l2, smooth_ic

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _params(g, key):
    return json.loads(str(g[key]))


@pytest.mark.parametrize("case", ["d0", "d1", "d2"])
def test_direct_vs_reference(case):
    """d0: the module's own 50x50 demo (chip path); d1: non-square, mixed BC order; d2: BASELINE
    config 2a, 256x256 (stream path)."""
    from nns_b200.direct_fd.simulate import NavierStokesSystem
    g = load_golden("direct_fd")
    P = _params(g, case + "_params")
    nx, ny = P["nx"], P["ny"]
    if case + "_u0" in g.files:
        u0, v0, p0 = g[case + "_u0"].copy(), g[case + "_v0"].copy(), g[case + "_p0"].copy()
    else:
        u0, v0, p0 = np.zeros((nx, ny)), np.zeros((nx, ny)), np.zeros((nx, ny))
    s = NavierStokesSystem(u0, v0, p0, make_bcs(P["u_bc"], nx, ny), make_bcs(P["v_bc"], nx, ny),
                           make_bcs(P["p_bc"], nx, ny), nt=P["nt"], nit=P["nit"], nx=nx, ny=ny, dt=P["dt"],
                           rho=P["rho"], nu=P["nu"])
    u, v, p = s.simulate()
    fr = g[case + "_frames"]
    for a, name in ((u, "_u"), (v, "_v"), (p, "_p")):
        for k, n in enumerate(fr):
            assert rel_l2(a[n], g[case + name][k]) <= TOL, (name, n)
    norms = np.stack([[np.linalg.norm(a[n].ravel()) for n in range(P["nt"])] for a in (u, v, p)])
    ref = g[case + "_norms"]
    assert np.max(np.abs(norms - ref) / np.maximum(ref, 1e-30)) <= 1e-9
    # the reference aliases its ICs (direct_fd/simulate.py:132): caller arrays hold the final state
    assert np.array_equal(u0, u[-1]) and np.array_equal(p0, p[-1])


def test_direct_step_in_place(oracle_fd):
    from nns_b200.direct_fd.simulate import NavierStokesSystem
    nx, ny = 17, 23
    spec_u = [["left", "dirichlet", 0.0], ["right", "dirichlet", 1.0], ["top", "neumann", 0.1], ["bottom", "dirichlet", 0.0]]
    spec_v = [["top", "dirichlet", 0.2], ["left", "neumann", 0.0], ["right", "dirichlet", 0.0]]
    spec_p = [["top", "dirichlet", 0.0], ["bottom", "neumann", 0.0], ["left", "neumann", 0.3], ["right", "neumann", 0.0]]
    u, v, p = smooth_ic(nx, ny, 5, amp=0.2)
    ou, ov, op = u.copy(), v.copy(), p.copy()
    s = NavierStokesSystem(u, v, p, make_bcs(spec_u, nx, ny), make_bcs(spec_v, nx, ny), make_bcs(spec_p, nx, ny),
                           nit=13, nx=nx, ny=ny, dt=5e-4, rho=1.2, nu=0.05)
    r = s.step(u, v, p)
    assert r[0] is u and r[1] is v and r[2] is p
    oracle_fd.direct_step(ou, ov, op, [tuple(b) for b in spec_u], [tuple(b) for b in spec_v],
                          [tuple(b) for b in spec_p], nit=13, dt=5e-4, rho=1.2, nu=0.05)
    assert rel_l2(u, ou) <= TOL and rel_l2(v, ov) <= TOL and rel_l2(p, op) <= TOL


@pytest.mark.parametrize("shape", [(40, 36), (72, 60), (33, 130)])
def test_direct_chip_cluster_and_stream_paths_agree(oracle_fd, monkeypatch, shape):
    """The same inputs through the three code paths -- chip (member in one SM's shared memory), cluster (row bands
    over a thread-block cluster, halo rows through distributed shared memory) and stream (one launch per sweep) --
    for an odd nit, an odd number of steps, mixed BC order and a batch with per-member nu: each against the oracle
    (rel-L2 <= 1e-10) and against each other (same expression per cell up to the compiler's FMA contraction: rel-L2 <= 1e-13)."""
    import torch
    import nns_b200
    from nns_b200.ensemble import DirectEnsemble
    D, Nm = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    B = 3
    nx, ny = shape
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc = [D(0.0, 'left', dx, dy), D(1.0, 'right', dx, dy), Nm(0.1, 'top', dx, dy), D(0.0, 'bottom', dx, dy)]
    v_bc = [D(0.2, 'top', dx, dy), Nm(0.0, 'left', dx, dy), D(0.0, 'right', dx, dy), Nm(-0.1, 'bottom', dx, dy)]
    p_bc = [D(0.0, 'top', dx, dy), Nm(0.0, 'bottom', dx, dy), Nm(0.3, 'left', dx, dy), Nm(0.0, 'right', dx, dy)]
    nus = np.array([0.05, 0.1, 0.02])
    f = [np.stack([smooth_ic(nx, ny, 40 + 3 * b + k, amp=0.1)[0] for b in range(B)]) for k in range(3)]
    out = {}
    for mode in ("chip", "cluster", "cluster_regs", "stream"):
        monkeypatch.setenv("NNS_DIRECT_MODE", mode.split("_")[0])
        monkeypatch.setenv("NNS_DIRECT_NPT", "4" if mode == "cluster_regs" else "0")     # p / b of a thread's pairs in registers
        ens = DirectEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=7, dt=2e-4, rho=1.1, nu=nus)
        ens.set_state(*f)
        tu, tv, tp = ens.run(5, trajectory=True)
        torch.cuda.synchronize()
        out[mode] = (tu.clone(), tv.clone(), tp.clone(), ens.u.clone(), ens.v.clone(), ens.p.clone(), ens.launches)
    monkeypatch.delenv("NNS_DIRECT_MODE")
    monkeypatch.delenv("NNS_DIRECT_NPT")
    assert out["chip"][6] == 1 and out["cluster"][6] == 1 and out["cluster_regs"][6] == 1 and out["stream"][6] > 50     # the paths really ran
    for b in range(B):
        u, v, p = oracle_fd.direct_simulate(f[0][b], f[1][b], f[2][b], u_bc, v_bc, p_bc, nt=5, nit=7, dt=2e-4,
                                            rho=1.1, nu=nus[b])
        for mode in out:
            tu, tv, tp = out[mode][:3]
            assert rel_l2(tu[b].cpu().numpy(), u) <= TOL and rel_l2(tv[b].cpu().numpy(), v) <= TOL, mode
            assert rel_l2(tp[b].cpu().numpy(), p) <= TOL, mode
            assert rel_l2(out[mode][3][b].cpu().numpy(), u[-1]) <= TOL and rel_l2(out[mode][5][b].cpu().numpy(), p[-1]) <= TOL
    for k in range(6):
        for other in ("cluster", "cluster_regs", "stream"):
            d = float((out["chip"][k] - out[other][k]).norm() / out["chip"][k].norm())
            assert d <= 1e-13, (k, other, d)


@pytest.mark.parametrize("shape,mode", [((41, 48), "chip"), ((96, 128), "cluster"), ((96, 127), "cluster"), ((256, 256), "cluster")])
def test_direct_periodic_channel_extension(oracle_fd, monkeypatch, shape, mode):
    """BASELINE config 2b: channel flow, periodic in the differenced axis 1 with a body force (an EXTENSION: the reference
    has neither; reference-unpinned).  Oracle = numpy restatement of direct_fd:56-127 with wrapped column indices
    (oracle/fd.py direct_periodic_simulate); walls u = v = 0, dp/dn = 0; smooth start; chip and cluster paths."""
    import torch
    import nns_b200
    from nns_b200.ensemble import DirectEnsemble
    D, Nm = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    nx, ny = shape
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    walls = lambda cls, val: [cls(val, 'left', dx, dy), cls(val, 'right', dx, dy)]      # noqa: E731
    u_bc, v_bc, p_bc = walls(D, 0.0), walls(D, 0.0), walls(Nm, 0.0)
    x = np.linspace(0, 2 * np.pi, ny, endpoint=False)[None, :]
    y = np.linspace(-1, 1, nx)[:, None]
    u0 = 0.3 * (1 - y ** 2) * (1 + 0.2 * np.sin(2 * x))
    v0 = 0.05 * (1 - y ** 2) * np.cos(x)
    p0 = 0.1 * np.cos(3 * x) * y
    nt, nit, dt = (6, 11, 2e-4) if nx < 200 else (3, 50, 1e-4)
    monkeypatch.setenv("NNS_DIRECT_MODE", mode)
    ens = DirectEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=dt, rho=1.0, nu=0.1, periodic_x=True, force_x=1.0)
    ens.set_state(u0[None], v0[None], p0[None])
    for f in (0, 1, 2):
        ens.apply_bc(f, (ens.u, ens.v, ens.p)[f])
    tu, tv, tp = ens.run(nt, trajectory=True)
    torch.cuda.synchronize()
    monkeypatch.delenv("NNS_DIRECT_MODE")
    assert ens.launches <= 4
    a0 = [u0.copy(), v0.copy(), p0.copy()]
    for A, val in ((a0[0], 0.0), (a0[1], 0.0)):
        A[0, :] = val; A[-1, :] = val
    a0[2][0, :] = a0[2][1, :]; a0[2][-1, :] = a0[2][-2, :]
    ou, ov, op = oracle_fd.direct_periodic_simulate(a0[0], a0[1], a0[2], u_bc, v_bc, p_bc, nt, nit, dt, 1.0, 0.1, 1.0)
    for n in range(nt):
        assert rel_l2(tu[0, n].cpu().numpy(), ou[n]) <= TOL and rel_l2(tv[0, n].cpu().numpy(), ov[n]) <= TOL, n
        assert rel_l2(tp[0, n].cpu().numpy(), op[n]) <= TOL, n
    assert abs(float(tu[0, -1, 1:-1].mean()) - float(ou[-1][1:-1].mean())) < 1e-12      # the body force really drives u
