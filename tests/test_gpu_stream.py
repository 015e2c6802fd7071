"""GPU parity of the persistent warp-specialised chorin_fd kernel (chorin_fd_stream.cu, the path
BASELINE config 4 runs on: 128x128 members) against the oracle and against the one-CTA-per-member
kernel.  Tolerance: fp64 relative L2 <= 1e-10 on u, v, p; sweep counts exact."""
import numpy as np
import pytest

from tests._util import rel_l2, smooth_ic

pytestmark = pytest.mark.gpu
TOL = 1e-10
NX = NY = 128


def _bc_tuples(bcs):
    return [(b.boundary, b.type, float(b.value)) for b in bcs]


def _ens(B, u_bc, v_bc, p_bc, nu, lid=None, **kw):
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values
    args = dict(u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=2e-4, rho=1, nu=nu, beta=1.25, method='explicit')
    args.update(kw)
    if lid is not None:
        args["bc_values"] = cavity_bc_values(lid)
    return ChorinEnsemble(B, NX, NY, **args)


def test_stream_many_members_vs_oracle(oracle_fd, monkeypatch):
    """More members than SMs (several members per persistent CTA, ragged tail), cavity draw of
    config 4, 3 steps: sampled members vs the oracle."""
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    from nns_b200.ensemble import cavity_bcs, cavity_ensemble_params
    B = 333
    lid, nu = cavity_ensemble_params(B, seed=5)
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = _ens(B, u_bc, v_bc, p_bc, nu, lid)
    ens.init_variables()
    nsteps = 3
    sw = []
    for _ in range(nsteps):
        ens.step()
        sw.append(ens.sweeps.cpu().numpy().copy())
    z = np.zeros((NX, NY))
    for b in (0, 1, 147, 148, 149, 295, 296, 332):
        ub = _bc_tuples(u_bc)
        ub[1] = ("right", "dirichlet", float(lid[b]))
        ou, ov, op, osw = oracle_fd.chorin_simulate(z, z, z, ub, _bc_tuples(v_bc), _bc_tuples(p_bc), nt=nsteps, nit=50,
                                                    dt=2e-4, rho=1, nu=float(nu[b]), beta=1.25)
        assert rel_l2(ens.u[b].cpu().numpy(), ou[-1]) <= TOL
        assert rel_l2(ens.v[b].cpu().numpy(), ov[-1]) <= TOL
        assert rel_l2(ens.p[b].cpu().numpy(), op[-1]) <= TOL
        assert [int(s[b]) for s in sw] == list(osw)


def test_stream_mixed_bcs_smooth_ic_vs_oracle(oracle_fd, monkeypatch):
    """Neumann / Dirichlet mixes in list order (corners, Neumann reading freshly written lines) on
    u, v and p with a smooth non-zero start, via run() with trajectory."""
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    import nns_b200
    D, Nm = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    dx = dy = 2. / (NX - 1)
    u_bc = [Nm(0.3, 'left', dx, dy), D(0.7, 'right', dx, dy), Nm(-0.2, 'top', dx, dy), D(0.1, 'bottom', dx, dy)]
    v_bc = [D(0.0, 'top', dx, dy), Nm(0.5, 'bottom', dx, dy), D(-0.1, 'left', dx, dy), Nm(0.0, 'right', dx, dy)]
    p_bc = [Nm(0.0, 'left', dx, dy), D(0.2, 'top', dx, dy), Nm(0.1, 'right', dx, dy), Nm(0.0, 'bottom', dx, dy)]
    B = 5
    ics = [smooth_ic(NX, NY, 40 + b) for b in range(B)]
    ens = _ens(B, u_bc, v_bc, p_bc, 0.05, nit=30, dt=1e-4)
    ens.set_state(np.stack([c[0] for c in ics]), np.stack([c[1] for c in ics]), np.stack([c[2] for c in ics]))
    ens.init_variables()
    nsteps = 4
    tu, tv, tp, sw = ens.run(nsteps, trajectory=True, sweeps=True)
    for b in range(B):
        ou, ov, op, osw = oracle_fd.chorin_simulate(ics[b][0], ics[b][1], ics[b][2], _bc_tuples(u_bc), _bc_tuples(v_bc),
                                                    _bc_tuples(p_bc), nt=nsteps, nit=30, dt=1e-4, rho=1, nu=0.05, beta=1.25)
        for n in range(nsteps):
            assert rel_l2(tu[b, n].cpu().numpy(), ou[n]) <= TOL
            assert rel_l2(tv[b, n].cpu().numpy(), ov[n]) <= TOL
            assert rel_l2(tp[b, n].cpu().numpy(), op[n]) <= TOL
        assert list(sw[:, b].cpu().numpy()) == list(osw)


def test_stream_early_exit_vs_oracle(oracle_fd, monkeypatch):
    """Tiny lid velocities: the SOR loop of the reference exits after a few sweeps (different per
    member and step); the wavefront must reproduce the exact count and the capped result."""
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    from nns_b200.ensemble import cavity_bcs
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    lid = np.array([3e-7, 1e-6, 2e-6, 1e-5, 1e-4, 1.0])
    nu = np.full(len(lid), 0.1)
    ens = _ens(len(lid), u_bc, v_bc, p_bc, nu, lid)
    ens.init_variables()
    nsteps = 3
    sws = []
    for _ in range(nsteps):
        ens.step()
        sws.append(ens.sweeps.cpu().numpy().copy())
    sws = np.stack(sws)
    z = np.zeros((NX, NY))
    seen = set()
    for b in range(len(lid)):
        ub = _bc_tuples(u_bc)
        ub[1] = ("right", "dirichlet", float(lid[b]))
        ou, ov, op, osw = oracle_fd.chorin_simulate(z, z, z, ub, _bc_tuples(v_bc), _bc_tuples(p_bc), nt=nsteps, nit=50,
                                                    dt=2e-4, rho=1, nu=0.1, beta=1.25)
        assert list(sws[:, b]) == list(osw), (b, sws[:, b], osw)
        seen.update(int(s) for s in osw)
        assert rel_l2(ens.u[b].cpu().numpy(), ou[-1]) <= TOL
        assert rel_l2(ens.p[b].cpu().numpy(), op[-1]) <= TOL
    assert min(seen) < 49 and max(seen) == 49, seen      # the case really exercises both branches


def test_stream_equals_member_kernel(monkeypatch):
    """Persistent kernel vs the one-CTA-per-member register-block kernel: same arithmetic per cell,
    so the fields agree to rounding (<= 1e-13) and the sweep counts exactly -- including with a
    loose tolerance that makes members stop early at different sweeps."""
    import torch
    from nns_b200.ensemble import cavity_bcs, cavity_ensemble_params
    B = 160
    lid, nu = cavity_ensemble_params(B, seed=9)
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    out = {}
    for mode in ("stream", "reg76"):
        monkeypatch.setenv("NNS_CHIP_MODE", mode)
        ens = _ens(B, u_bc, v_bc, p_bc, nu, lid, tol=2e-2)
        ens.init_variables()
        sw = []
        for _ in range(4):
            ens.step()
            sw.append(ens.sweeps.clone())
        out[mode] = (ens.u.clone(), ens.v.clone(), ens.p.clone(), torch.stack(sw))
    a, b = out["stream"], out["reg76"]
    assert torch.equal(a[3], b[3])
    assert int(a[3].min()) < 49
    for x, y in zip(a[:3], b[:3]):
        assert float((x - y).norm() / y.norm()) <= 1e-13


def _run_once(monkeypatch, B, steps, seed=11):
    import torch
    from nns_b200.ensemble import cavity_bcs, cavity_ensemble_params
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    lid, nu = cavity_ensemble_params(B, seed=seed)
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = _ens(B, u_bc, v_bc, p_bc, nu, lid)
    ens.init_variables()
    sw = []
    for _ in range(steps):
        ens.step()
        sw.append(ens.sweeps.clone())
    return ens.u.clone(), ens.v.clone(), ens.p.clone(), torch.stack(sw)


def test_stream_kernel_is_bit_reproducible(monkeypatch):
    """Two runs of the same ensemble (several members per persistent CTA, ragged tail, 3 steps) agree bit for bit:
    the named-barrier / halo-slot hand-offs of the warp-specialised kernel leave no run-to-run freedom
    (scripts/determinism_check.py is the long version: 12+ runs x 4096 members)."""
    import torch
    a = _run_once(monkeypatch, 1200, 3)
    b = _run_once(monkeypatch, 1200, 3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("B", [300, 1200])
def test_step_host_chunked_pipeline_matches_device_step(monkeypatch, B):
    """nns_chorin_fd_step_host on the 128 x 128 stream path: the batch is cut into 2 (296 <= B < 1184) or 8
    (B >= 1184) member chunks whose launches run on 4 internal streams and may overlap; every chunk has its own
    scratch set, so the result equals the one-launch device-resident step bit for bit (3 steps, host roles rotated
    as bench.py does)."""
    import torch
    from nns_b200 import _lib
    from nns_b200.ensemble import cavity_bcs, cavity_ensemble_params
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    lid, nu = cavity_ensemble_params(B, seed=3)
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = _ens(B, u_bc, v_bc, p_bc, nu, lid)
    ens.init_variables()
    hs = [torch.zeros((B, NX, NY), dtype=torch.float64).pin_memory() for _ in range(7)]
    hsw = torch.zeros((B,), dtype=torch.int32).pin_memory()
    for k, src in enumerate((ens.u, ens.v, ens.u1, ens.v1, ens.p)):
        hs[k].copy_(src)
    torch.cuda.synchronize()
    L = _lib.lib()
    for _ in range(3):
        _lib.check(L.nns_chorin_fd_step_host(ens.handle.h, hs[0].data_ptr(), hs[1].data_ptr(), hs[2].data_ptr(),
                                             hs[3].data_ptr(), hs[4].data_ptr(), hs[5].data_ptr(), hs[6].data_ptr(),
                                             hsw.data_ptr()))
        hs[2], hs[0], hs[5] = hs[0], hs[5], hs[2]
        hs[3], hs[1], hs[6] = hs[1], hs[6], hs[3]
        ens.step()
        assert torch.equal(hs[0], ens.u.cpu()) and torch.equal(hs[1], ens.v.cpu()) and torch.equal(hs[4], ens.p.cpu())
        assert torch.equal(hsw, ens.sweeps.cpu())


def test_full_size_ensemble_4096_members(oracle_fd, monkeypatch):
    """BASELINE config 4 at full size (4096 members, the bench workload), two steps: sampled members against the
    oracle (fp64 rel-L2 <= 1e-10, sweep counts exact) and, as the size-independent property, every sampled
    member bit-identical to the same member stepped in a batch of its own (members never interact, whichever
    persistent CTA and position in its queue they land on)."""
    import torch
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    from nns_b200.ensemble import cavity_bcs, cavity_ensemble_params
    B = 4096
    lid, nu = cavity_ensemble_params(B, seed=0)
    dx = dy = 2. / (NX - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = _ens(B, u_bc, v_bc, p_bc, nu, lid)
    ens.init_variables()
    sw = []
    for _ in range(2):
        ens.step()
        sw.append(ens.sweeps.clone())
    assert bool(torch.isfinite(ens.p).all())
    sample = [0, 147, 148, 2047, 4000, 4095]
    small = _ens(len(sample), u_bc, v_bc, p_bc, nu[sample], lid[sample])
    small.init_variables()
    for _ in range(2):
        small.step()
    z = np.zeros((NX, NY))
    for k, b in enumerate(sample):
        assert torch.equal(ens.u[b], small.u[k]) and torch.equal(ens.v[b], small.v[k]) and torch.equal(ens.p[b], small.p[k])
        ub = _bc_tuples(u_bc)
        ub[1] = ("right", "dirichlet", float(lid[b]))
        ou, ov, op, osw = oracle_fd.chorin_simulate(z, z, z, ub, _bc_tuples(v_bc), _bc_tuples(p_bc), nt=2, nit=50,
                                                    dt=2e-4, rho=1, nu=float(nu[b]), beta=1.25)
        assert rel_l2(ens.u[b].cpu().numpy(), ou[-1]) <= TOL
        assert rel_l2(ens.v[b].cpu().numpy(), ov[-1]) <= TOL
        assert rel_l2(ens.p[b].cpu().numpy(), op[-1]) <= TOL
        assert [int(s[b]) for s in sw] == list(osw)
