"""CPU: host logic of the row-slab path (csrc/chorin_fd_slab.cu) -- partition and tick plan from the C ABI --
and a world_size-2 gloo emulation of the tile hyperplane with per-tick halo exchange, checked bit for bit
against the lexicographic SOR loop of the reference (src/chorin_fd/simulate.py:190-196)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from nns_b200 import slab


def test_partition_covers_rows_once_and_aligns_with_tiles():
    for nx, world, tr in ((16384, 8, 0), (16384, 3, 0), (50, 2, 8), (41, 4, 4), (35, 5, 7)):
        rows = []
        for r in range(world):
            r0, n = slab.partition(nx, world, r, tr)
            rows += list(range(r0, r0 + n))
            TR = slab.plan(nx, 64, world, r, 0, 0, tr)["TR"]
            if r > 0:
                assert (r0 - 1) % TR == 0           # slabs start on a tile row
        assert rows == list(range(nx))
    with pytest.raises(Exception):
        slab.partition(20, 8, 0, 16)                # fewer tile rows than ranks


def test_plan_runs_every_tile_sweep_exactly_once_in_dependency_order():
    nx, ny, world, tr, cap = 70, 300, 3, 8, 5
    g = slab.plan(nx, ny, world, 0, 0, 0, tr)
    nI, nJ = g["nI"], g["nJ"]
    tick_of = {}
    for T in range(nI + nJ + 2 * cap):
        for r in range(world):
            for s in range(cap):
                pl = slab.plan(nx, ny, world, r, T, s, tr)
                for I in range(pl["Ilo"], pl["Ihi"] + 1):
                    assert pl["I0"] <= I < pl["I1"]
                    J = T - 2 * s - I
                    assert 0 <= J < nJ and (I, J, s) not in tick_of
                    tick_of[(I, J, s)] = T
    assert len(tick_of) == nI * nJ * cap
    for (I, J, s), T in tick_of.items():           # lexicographic predecessors ran at an earlier tick
        for dep in ((I - 1, J, s), (I, J - 1, s), (I + 1, J, s - 1), (I, J + 1, s - 1)):
            if dep in tick_of:
                assert tick_of[dep] < T


def _sor_reference(p, C, nsweeps, dx, dy, beta):
    """The reference's loop (chorin_fd/simulate.py:190-196), same expression."""
    p = p.copy()
    nx, ny = p.shape
    for _ in range(nsweeps):
        for i in range(1, nx - 1):
            for j in range(1, ny - 1):
                p[i, j] = beta * (dy ** 2 * p[i + 1, j] + dy ** 2 * p[i - 1, j] + dx ** 2 * p[i, j + 1] +
                                  dx ** 2 * p[i, j - 1] - C[i, j]) / (2 * dx ** 2 + 2 * dy ** 2) + (1 - beta) * p[i, j]
    return p


def _slab_worker(rank, world, port, nx, ny, tr, cap, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    P0, Cg = rng.normal(size=(nx, ny)), rng.normal(size=(nx, ny))
    dx, dy, beta = 2. / (nx - 1), 2. / (ny - 1), 1.25
    r0, nr = slab.partition(nx, world, rank, tr)
    g = slab.plan(nx, ny, world, rank, 0, 0, tr)
    TR, TC, nI, nJ = g["TR"], g["TC"], g["nI"], g["nJ"]
    # local slab with halo rows: local row l <-> global row r0 - 1 + l
    lo = max(r0 - 1, 0)
    p = np.zeros((nr + 2, ny))
    p[(lo - (r0 - 1)):(lo - (r0 - 1)) + min(r0 + nr + 1, nx) - lo] = P0[lo:min(r0 + nr + 1, nx)]
    L = lambda i: i - (r0 - 1)  # noqa: E731

    def exchange():
        ops = []
        bufs = {}
        if rank > 0:
            bufs["up"] = torch.empty(ny, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, torch.from_numpy(p[1].copy()), rank - 1), dist.P2POp(dist.irecv, bufs["up"], rank - 1)]
        if rank < world - 1:
            bufs["dn"] = torch.empty(ny, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, torch.from_numpy(p[nr].copy()), rank + 1), dist.P2POp(dist.irecv, bufs["dn"], rank + 1)]
        for w in dist.batch_isend_irecv(ops) if ops else []:
            w.wait()
        if "up" in bufs:
            p[0] = bufs["up"].numpy()
        if "dn" in bufs:
            p[nr + 1] = bufs["dn"].numpy()

    for T in range(nI + nJ + 2 * cap):
        for s in range(cap):
            pl = slab.plan(nx, ny, world, rank, T, s, tr)
            for I in range(pl["Ilo"], pl["Ihi"] + 1):
                J = T - 2 * s - I
                for i in range(1 + I * TR, min(1 + (I + 1) * TR, nx - 1)):
                    for j in range(max(1, J * TC), min((J + 1) * TC, ny - 1)):
                        l = L(i)
                        p[l, j] = beta * (dy ** 2 * p[l + 1, j] + dy ** 2 * p[l - 1, j] + dx ** 2 * p[l, j + 1] +
                                          dx ** 2 * p[l, j - 1] - Cg[i, j]) / (2 * dx ** 2 + 2 * dy ** 2) + (1 - beta) * p[l, j]
        exchange()
    want = _sor_reference(P0, Cg, cap, dx, dy, beta)
    q.put((rank, bool(np.array_equal(p[1:nr + 1], want[r0:r0 + nr]))))
    dist.destroy_process_group()


def test_two_rank_gloo_tile_hyperplane_equals_lexicographic_sor():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    nx, ny, tr, cap = 22, 150, 4, 4          # 5 tile rows x 2 tile columns, 4 sweeps
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, nx, ny, tr, cap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
