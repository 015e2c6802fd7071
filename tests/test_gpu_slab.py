"""GPU parity of the tiled / row-slab chorin_fd path (csrc/chorin_fd_slab.cu) -- grids that do not fit one SM,
and one grid split over several GPUs -- against the oracle.  fp64 relative L2 <= 1e-10, sweep counts exact.
The multi-GPU test needs >= 2 visible GPUs (gpurun --gpus 2) and is skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from tests._util import rel_l2, smooth_ic

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _bcs(nx, ny, kind="cavity", lid=1.0):
    import nns_b200
    D, N = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    if kind == "cavity":
        u_bc = [D(0, 'left', dx, dy), D(lid, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
        v_bc = [D(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
        p_bc = [D(0, 'top', dx, dy), N(0, 'bottom', dx, dy), N(0, 'left', dx, dy), N(0, 'right', dx, dy)]
    else:
        u_bc = [N(0.3, 'left', dx, dy), D(0.7, 'right', dx, dy), N(-0.2, 'top', dx, dy), D(0.1, 'bottom', dx, dy)]
        v_bc = [D(0.0, 'top', dx, dy), N(0.5, 'bottom', dx, dy), D(-0.1, 'left', dx, dy), N(0.0, 'right', dx, dy)]
        p_bc = [N(0.0, 'left', dx, dy), D(0.2, 'top', dx, dy), N(0.1, 'right', dx, dy), N(0.0, 'bottom', dx, dy)]
    return u_bc, v_bc, p_bc


def _t(bcs):
    return [(b.boundary, b.type, float(b.value)) for b in bcs]


@pytest.mark.parametrize("nx,ny,kind,tr", [(150, 301, "cavity", 16), (200, 140, "mixed", 32), (137, 260, "mixed", 7)])
def test_tiled_simulate_vs_oracle(oracle_fd, monkeypatch, nx, ny, kind, tr):
    """NavierStokesSystem.simulate() on grids beyond the on-chip path (several tile rows and columns, ragged
    last tiles), zero and smooth starts, trajectory of every step."""
    monkeypatch.setenv("NNS_SLAB_TR", str(tr))
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    u_bc, v_bc, p_bc = _bcs(nx, ny, kind)
    ic = [np.zeros((nx, ny))] * 3 if kind == "cavity" else smooth_ic(nx, ny, 11)
    kw = dict(nt=4, nit=25, dt=1e-4, rho=1, nu=0.05)
    s = NavierStokesSystem(ic[0], ic[1], ic[2], u_bc, v_bc, p_bc, nx=nx, ny=ny, beta=1.25, method='explicit', **kw)
    u, v, p = s.simulate()
    ou, ov, op, sw = oracle_fd.chorin_simulate(ic[0], ic[1], ic[2], _t(u_bc), _t(v_bc), _t(p_bc), beta=1.25, **kw)
    for n in range(kw["nt"]):
        assert rel_l2(u[n], ou[n]) <= TOL and rel_l2(v[n], ov[n]) <= TOL and rel_l2(p[n], op[n]) <= TOL, n
    assert np.array_equal(s.last_sweeps, sw)


def test_tiled_early_exit_vs_oracle(oracle_fd, monkeypatch):
    monkeypatch.setenv("NNS_SLAB_TR", "16")
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    nx, ny = 160, 290
    z = np.zeros((nx, ny))
    seen = set()
    for lid in (2e-6, 1e-4):
        u_bc, v_bc, p_bc = _bcs(nx, ny, "cavity", lid)
        kw = dict(nt=3, nit=50, dt=2e-4, rho=1, nu=0.1)
        s = NavierStokesSystem(z, z.copy(), z.copy(), u_bc, v_bc, p_bc, nx=nx, ny=ny, beta=1.25, method='explicit', **kw)
        u, v, p = s.simulate()
        ou, ov, op, sw = oracle_fd.chorin_simulate(z, z, z, _t(u_bc), _t(v_bc), _t(p_bc), beta=1.25, **kw)
        assert np.array_equal(s.last_sweeps, sw), (lid, s.last_sweeps, sw)
        assert rel_l2(u[-1], ou[-1]) <= TOL and rel_l2(p[-1], op[-1]) <= TOL
        seen.update(int(x) for x in sw)
    assert min(seen) < 49 and max(seen) == 49


def test_slab_class_single_rank_equals_simulate(oracle_fd, monkeypatch):
    monkeypatch.setenv("NNS_SLAB_TR", "16")
    from nns_b200.slab import SlabChorin
    nx, ny = 130, 270
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 5)
    sl = SlabChorin(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=20, dt=1e-4, rho=1, nu=0.05, beta=1.25, rank=0, world=1)
    sl.set_state(*ic)
    sl.init_variables()
    sws = []
    for _ in range(3):
        sl.step()
        sws.append(sl.last_sweeps)
    ou, ov, op, sw = oracle_fd.chorin_simulate(ic[0], ic[1], ic[2], _t(u_bc), _t(v_bc), _t(p_bc), nt=3, nit=20, dt=1e-4,
                                               rho=1, nu=0.05, beta=1.25)
    assert rel_l2(sl.gather(sl.u), ou[-1]) <= TOL and rel_l2(sl.gather(sl.p), op[-1]) <= TOL
    assert sws == list(sw)


def _rank_main(rank, world, port, nx, ny, nsteps, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NNS_SLAB_TR="16")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nns_b200.slab import SlabChorin
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 21)
    sl = SlabChorin(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=30, dt=1e-4, rho=1, nu=0.05, beta=1.25)
    sl.set_state(*ic)
    sl.init_variables()
    sws = []
    for _ in range(nsteps):
        sl.step()
        sws.append(sl.last_sweeps)
    u, v, p = sl.gather(sl.u), sl.gather(sl.v), sl.gather(sl.p)
    if rank == 0:
        q.put((u, v, p, sws))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slabs_over_gpus_vs_oracle(oracle_fd, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    nx, ny, nsteps = 200, 290, 3
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, nx, ny, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    u, v, p, sws = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 21)
    ou, ov, op, sw = oracle_fd.chorin_simulate(ic[0], ic[1], ic[2], _t(u_bc), _t(v_bc), _t(p_bc), nt=nsteps, nit=30,
                                               dt=1e-4, rho=1, nu=0.05, beta=1.25)
    assert rel_l2(u, ou[-1]) <= TOL and rel_l2(v, ov[-1]) <= TOL and rel_l2(p, op[-1]) <= TOL
    assert sws == list(sw)


def test_ensemble_beyond_one_sm_vs_oracle(oracle_fd, monkeypatch):
    """ChorinEnsemble on a grid that fits neither the 128 x 128 stream kernel nor one SM (200 x 260): the members of
    the batch go through the tiled path one after the other, each with its own nu and lid value; smooth start, 2 steps
    with trajectories and sweep counts against the oracle (rel-L2 <= 1e-10, counts exact)."""
    monkeypatch.delenv("NNS_CHIP_MODE", raising=False)
    from nns_b200.ensemble import ChorinEnsemble, cavity_bc_values, cavity_bcs
    B, nx, ny = 3, 200, 260
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    lid, nu = np.array([0.7, 1.0, 1.3]), np.array([0.02, 0.05, 0.1])
    ics = [smooth_ic(nx, ny, 70 + b, amp=0.05) for b in range(B)]
    ens = ChorinEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=12, dt=2e-5, rho=1, nu=nu, beta=1.25,
                         method='explicit', bc_values=cavity_bc_values(lid))
    ens.set_state(np.stack([c[0] for c in ics]), np.stack([c[1] for c in ics]), np.stack([c[2] for c in ics]))
    ens.init_variables()
    tu, tv, tp, sw = ens.run(2, trajectory=True, sweeps=True)
    for b in range(B):
        ub = [(bc.boundary, bc.type, float(bc.value)) for bc in u_bc]
        ub[1] = ("right", "dirichlet", float(lid[b]))
        ou, ov, op, osw = oracle_fd.chorin_simulate(ics[b][0], ics[b][1], ics[b][2], ub, v_bc, p_bc, nt=2, nit=12, dt=2e-5,
                                                    rho=1, nu=float(nu[b]), beta=1.25)
        for n in range(2):
            assert rel_l2(tu[b, n].cpu().numpy(), ou[n]) <= 1e-10 and rel_l2(tv[b, n].cpu().numpy(), ov[n]) <= 1e-10
            assert rel_l2(tp[b, n].cpu().numpy(), op[n]) <= 1e-10
        assert list(sw[:, b].cpu().numpy()) == list(osw)


def test_semi_implicit_beyond_one_sm_vs_oracle(oracle_fd):
    """method='semi_implicit' (the reference's default) on a 200 x 200 grid, which does not fit one SM: predictor by
    column-parallel Thomas solves along axis 0 in HBM + the tiled SOR; drop-in class, 3 steps against the oracle."""
    from nns_b200.chorin_fd.simulate import NavierStokesSystem
    from nns_b200.ensemble import cavity_bcs
    nx = ny = 200
    dx = dy = 2. / (nx - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    u0, v0, p0 = smooth_ic(nx, ny, 9, amp=0.05)
    s = NavierStokesSystem(u0, v0, p0, u_bc, v_bc, p_bc, nt=3, nit=15, nx=nx, ny=ny, dt=5e-5, rho=1, nu=0.1, beta=1.25,
                           method='semi_implicit')
    u, v, p = s.simulate()
    ou, ov, op, osw = oracle_fd.chorin_simulate(u0, v0, p0, u_bc, v_bc, p_bc, nt=3, nit=15, dt=5e-5, rho=1, nu=0.1,
                                                beta=1.25, method='semi_implicit')
    for n in range(3):
        assert rel_l2(u[n], ou[n]) <= 1e-10 and rel_l2(v[n], ov[n]) <= 1e-10 and rel_l2(p[n], op[n]) <= 1e-10, n
    assert list(s.last_sweeps) == list(osw)


def _direct_rank_main(rank, world, port, nx, ny, nsteps, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    from nns_b200.slab import SlabDirect
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 33, amp=0.1)
    sl = SlabDirect(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=9, dt=1e-4, rho=1.1, nu=0.05)
    sl.set_state(*ic)
    sl.sync_halos()
    sl.run(nsteps)
    u, v, p = sl.gather(sl.u), sl.gather(sl.v), sl.gather(sl.p)
    if rank == 0:
        q.put((u, v, p))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_direct_fd_slabs_vs_oracle(oracle_fd, world):
    """direct_fd on row slabs (one halo row of p per Jacobi sweep over NCCL; u, v once per step), 130 x 200, mixed BC
    order, 3 steps (an odd number: the ping-pong copy-back), against the oracle; world = 1 runs the same kernels without
    a communicator."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    nx, ny, nsteps = 130, 200, 3
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 33, amp=0.1)
    if world == 1:
        from nns_b200.slab import SlabDirect
        sl = SlabDirect(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=9, dt=1e-4, rho=1.1, nu=0.05, rank=0, world=1)
        sl.set_state(*ic)
        sl.sync_halos()
        sl.run(nsteps)
        u, v, p = sl.gather(sl.u), sl.gather(sl.v), sl.gather(sl.p)
    else:
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_direct_rank_main, args=(r, world, port, nx, ny, nsteps, q)) for r in range(world)]
        for pr in procs:
            pr.start()
        u, v, p = q.get(timeout=600)
        for pr in procs:
            pr.join(timeout=120)
    ou, ov, op = oracle_fd.direct_simulate(ic[0].copy(), ic[1].copy(), ic[2].copy(), _t(u_bc), _t(v_bc), _t(p_bc), nt=nsteps,
                                           nit=9, dt=1e-4, rho=1.1, nu=0.05)
    assert rel_l2(u, ou[-1]) <= TOL and rel_l2(v, ov[-1]) <= TOL and rel_l2(p, op[-1]) <= TOL


def test_direct_fd_slab_fused_bcs_equal_list_walk(oracle_fd, monkeypatch):
    """The p BCs applied inside the Jacobi kernel (last entry per side + corner replay in registers) against the list walk
    with one launch per entry (NNS_DSLAB_BC=list): mixed BC order with non-zero Neumann values, same results to rounding."""
    from nns_b200.slab import SlabDirect
    nx, ny, nsteps = 70, 90, 3
    u_bc, v_bc, p_bc = _bcs(nx, ny, "mixed")
    ic = smooth_ic(nx, ny, 5, amp=0.1)
    out = {}
    for mode in ("fused", "fused_scalar", "list"):
        monkeypatch.setenv("NNS_DSLAB_BC", mode.split("_")[0])
        if mode == "fused_scalar":
            monkeypatch.setenv("NNS_DSLAB_SCALAR", "1")       # one cell per thread instead of two with 128-bit accesses
        else:
            monkeypatch.delenv("NNS_DSLAB_SCALAR", raising=False)
        sl = SlabDirect(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=7, dt=1e-4, rho=1.1, nu=0.05, rank=0, world=1)
        sl.set_state(*ic)
        sl.sync_halos()
        l0 = sl.launches
        sl.run(nsteps)
        out[mode] = (sl.gather(sl.u), sl.gather(sl.v), sl.gather(sl.p), sl.launches - l0)
    monkeypatch.delenv("NNS_DSLAB_BC")
    assert out["fused"][3] < out["list"][3]                   # the list walk really launches a kernel per entry
    for k in range(3):
        assert rel_l2(out["fused"][k], out["list"][k]) <= 1e-13
        assert rel_l2(out["fused"][k], out["fused_scalar"][k]) <= 1e-13
    ou, ov, op = oracle_fd.direct_simulate(ic[0].copy(), ic[1].copy(), ic[2].copy(), _t(u_bc), _t(v_bc), _t(p_bc), nt=nsteps,
                                           nit=7, dt=1e-4, rho=1.1, nu=0.05)
    assert rel_l2(out["fused"][2], op[-1]) <= TOL and rel_l2(out["fused"][0], ou[-1]) <= TOL
