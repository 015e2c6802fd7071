#!/usr/bin/env python
"""bench.py -- headline benchmark: chorin_fd step throughput (cell-updates/s) on the batched
ensemble of BASELINE.json configs[3]: 4096 lid-driven cavities at 128x128 fp64, nit=50,
dt=2e-4, lid ~ U[0.5,1.5], Re ~ U[10,100] (default_rng(0)); one ensemble of 4096 members PER GPU
(members never communicate => weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one fused time step (predictor + u/v BCs + <=49 exact-order SOR sweeps + p BCs +
projection) of every member: ONE kernel launch.  Prints one JSON line (see the task contract):
`value` = device-resident throughput, `e2e` = the same metric through the host-buffer C-ABI call
(nns_chorin_fd_step_host: 5 fields H2D, 3 fields D2H per step, pinned memory), `roofline` for the
fused kernel against the measured HBM peak, `cpu_baseline` = the oracle port on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chorin_step_cell_updates_per_sec"
UNIT = "cell-updates/s"
BYTES_PER_CELL_UPDATE = 64      # reads u,v,u1,v1,p + writes u,v,p, 8 B each

WORKLOADS = {
    # name: (nx, ny, members per GPU, nit, dt)
    "ensemble4096_cavity128": (128, 128, 4096, 50, 2e-4),
    "ensemble_cavity41": (41, 41, 8192, 50, 1e-3),
}
# one large grid split into row slabs over the GPUs (BASELINE configs[4]): name -> (nx, ny, nit, dt, nu)
SLAB_WORKLOADS = {
    "slab_cavity16384": (16384, 16384, 50, 1e-8, 0.1),
    "slab_cavity4096": (4096, 4096, 50, 1.5e-7, 0.1),
}
# direct_fd (Jacobi) on row slabs: name -> (nx, ny, nit, dt, nu)
DIRECT_SLAB_WORKLOADS = {
    "direct_slab8192": (8192, 8192, 50, 1e-8, 0.1),
}
BYTES_PER_CELL_SWEEP = 24       # read p, read C', write p in an un-blocked sweep


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.f = index, None, None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def cavity_oracle_sample(nx, ny, members, nit, dt, seed_offset=0):
    """Host state + BC lists for `members` cavities of the workload (same draw as the GPU arm)."""
    from nns_b200.ensemble import cavity_ensemble_params
    lid, nu = cavity_ensemble_params(max(4096, members + seed_offset), seed=0, lo=seed_offset,
                                     hi=seed_offset + members)
    u = np.zeros((members, nx, ny))
    u[:, -1, :] = lid[:, None]          # u 'right' Dirichlet = lid; later entries zero the corners
    u[:, :, -1] = 0.0
    u[:, :, 0] = 0.0
    v, p = np.zeros_like(u), np.zeros_like(u)
    mk = lambda l: [("left", "dirichlet", 0.0), ("right", "dirichlet", float(l)), ("top", "dirichlet", 0.0),
                    ("bottom", "dirichlet", 0.0)]
    u_bcs = [mk(l) for l in lid]
    v_bc = [(s, "dirichlet", 0.0) for s in ("left", "right", "top", "bottom")]
    p_bc = [("top", "dirichlet", 0.0), ("bottom", "neumann", 0.0), ("left", "neumann", 0.0),
            ("right", "neumann", 0.0)]
    return u, v, p, nu, u_bcs, v_bc, p_bc


def cpu_port_throughput(nx, ny, nit, dt, budget_s, steps_per_call=1, members=None):
    """Time the oracle port (oracle/oracle.c, OpenMP over members) on a bounded sample sized from
    a short calibration run so that it takes about `budget_s` seconds on this host."""
    from oracle import fd as ofd
    ofd.build()
    T = ofd.set_threads()           # every host core, whatever OMP_NUM_THREADS says (torchrun exports 1)

    def once(m):
        u, v, p, nu, u_bcs, v_bc, p_bc = cavity_oracle_sample(nx, ny, m, nit, dt)
        u1, v1 = u.copy(), v.copy()
        t0 = time.perf_counter()
        _, used = ofd.chorin_ensemble_run(u, v, u1, v1, p, u_bcs, v_bc, p_bc, nt=steps_per_call, nit=nit, dt=dt,
                                          rho=1, nu=nu, beta=1.25)
        return time.perf_counter() - t0, used, (u, v, u1, v1, p, nu, u_bcs, v_bc, p_bc)

    if members is None:
        m0 = max(8, T)
        el0, _, _ = once(m0)
        members = int(min(4096, max(m0, T * int(budget_s / max(el0, 1e-6) * m0 / T))))
    el, used, state = once(members)
    return members * nx * ny * steps_per_call / el, used, members, el, state


def workload_config(workload, nx, ny, B, world, nit, dt):
    """The `config` object of the JSON line: identical for the b200 and the reference arm."""
    return {"workload": workload, "nx": nx, "ny": ny, "members_per_gpu": B, "global_members": B * world,
            "nit": nit, "dt": dt, "beta": 1.25, "method": "explicit",
            "parallelism": "ensemble members sharded, no collective",
            "l2": "state per GPU %.2f GB >> 126 MB L2 (inputs larger than L2, no flush needed)"
                  % (5 * B * nx * ny * 8 / 1e9)}


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm on the host cores.  The reference itself is
    pure Python (2.7 us per cell-sweep) and is not on this box; its C restatement (oracle port,
    bit-exact vs the reference, all host threads) is what is timed."""
    if rank != 0:
        return
    nx, ny, B, nit, dt = WORKLOADS[args.workload]
    if args.members:
        B = args.members
    from oracle import fd as ofd
    ofd.build()
    T = ofd.set_threads()           # every host core, whatever OMP_NUM_THREADS says (torchrun exports 1)
    members = int(min(B, max(T, 8 * T)))
    _, _, _, _, st = cpu_port_throughput(nx, ny, nit, dt, 0, 1, members)      # allocs + first touch
    u, v, u1, v1, p, nu, u_bcs, v_bc, p_bc = st
    for _ in range(max(0, args.warmup - 1)):
        ofd.chorin_ensemble_run(u, v, u1, v1, p, u_bcs, v_bc, p_bc, nt=1, nit=nit, dt=dt, rho=1, nu=nu, beta=1.25)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ofd.chorin_ensemble_run(u, v, u1, v1, p, u_bcs, v_bc, p_bc, nt=1, nit=nit, dt=dt, rho=1, nu=nu, beta=1.25)
    el = time.perf_counter() - t0
    val = members * nx * ny * args.steps / el
    sample = "%d of %d members x %d steps (oracle/oracle.c, OpenMP)" % (members, B, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, nx, ny, B, world, nit, dt),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": T, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def slab_measure(workload, steps, warmup, rank, world, local, with_clocks=True):
    """One large chorin_fd grid on row slabs (strong scaling): a step = one time step of the whole grid.
    The process group must exist already for world > 1.  Returns the JSON object on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from nns_b200.ensemble import cavity_bcs
    from nns_b200.slab import SlabChorin
    nx, ny, nit, dt, nu = SLAB_WORKLOADS[workload]
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    sl = SlabChorin(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=dt, rho=1, nu=nu, beta=1.25)
    sl.init_variables()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        sl.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
    l0 = sl.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sor_ms, sweeps = [], []
    for _ in range(steps):
        sl.step()
        ms, ticks = sl.last_sor_timing()
        sor_ms.append(ms)
        sweeps.append(sl.last_sweeps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 and with_clocks else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    finite = bool(torch.isfinite(sl.u).all() and torch.isfinite(sl.p).all())
    mode = "single GPU" if world == 1 else ("peer-memory mailboxes" if sl.p2p else "nccl send/recv")
    out = None
    if rank == 0:
        peak, peak_src = peaks()
        cells = nx * ny
        kms = float(np.mean(sor_ms)) / ticks          # average sweep-kernel launch (one tick), exchange included for N > 1
        S = float(np.mean(sweeps))
        achieved = BYTES_PER_CELL_SWEEP * (cells / world) * S / ticks / (kms * 1e-3) / 1e9
        fused = BYTES_PER_CELL_UPDATE * cells / (ms_total / steps * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": cells * steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "nx": nx, "ny": ny, "nit": nit, "dt": dt, "nu": nu, "beta": 1.25,
                       "method": "explicit", "parallelism": "row slabs, halo rows of p per SOR tick (NCCL or peer mailboxes)",
                       "exchange_mode": mode,
                       "l2": "p + C' per GPU %.2f GB >> 126 MB L2" % (2 * cells * 8 / world / 1e9),
                       "sweeps_per_step": [int(min(sweeps)), int(max(sweeps))], "ticks_per_step": int(ticks),
                       "finite": finite},
            "e2e": None, "gpu_launches": int(sl.launches - l0), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": fused, "peak": peak, "unit": "GB/s", "frac": fused / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "slab step (all kernels of a time step)",
                         "bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
                         "sweep_kernel": {"achieved": achieved, "frac": achieved / peak, "kernel_ms": kms,
                                          "bytes_per_cell_sweep": BYTES_PER_CELL_SWEEP},
                         "note": "frac: fully fused model (64 B per cell-update); sweep_kernel: the un-blocked exact-order "
                                 "sweeps as they are implemented (24 B per cell and sweep, per-tick launch average)"}}
    del sl
    torch.cuda.empty_cache()
    return out


def run_direct_slab(args, rank, world, local):
    """One large direct_fd grid on row slabs (strong scaling): a step = RHS + nit Jacobi sweeps (one halo row of p per
    sweep over NCCL) + velocity update."""
    import torch
    import torch.distributed as dist
    from nns_b200.ensemble import cavity_bcs
    from nns_b200.slab import SlabDirect
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nit, dt, nu = DIRECT_SLAB_WORKLOADS[args.workload]
    u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
    sl = SlabDirect(nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=dt, rho=1, nu=nu)
    sl.sync_halos()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sl.run(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sl.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sl.run(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    finite = bool(torch.isfinite(sl.u).all() and torch.isfinite(sl.p).all())
    if rank == 0:
        peak, peak_src = peaks()
        cells = nx * ny
        fused = 48 * cells / (ms_total / args.steps * 1e-3) / 1e9
        unblocked = (24 * nit + 120) * (cells / world) / (ms_total / args.steps * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": cells * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "solver": "direct_fd", "nx": nx, "ny": ny, "nit": nit, "dt": dt, "nu": nu,
                       "parallelism": "row slabs, one halo row of p per Jacobi sweep (NCCL send/recv)",
                       "l2": "p + b per GPU %.2f GB >> 126 MB L2" % (2 * cells * 8 / world / 1e9), "finite": finite},
            "e2e": None, "gpu_launches": int(sl.launches - l0), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": fused, "peak": peak, "unit": "GB/s", "frac": fused / peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "direct_fd slab step", "bytes_per_cell_update": 48,
                         "unblocked": {"achieved_per_gpu": unblocked, "frac_per_gpu": unblocked / peak,
                                       "bytes_per_cell_update": 24 * nit + 120},
                         "note": "frac: fully fused model (48 B per cell-update); unblocked: the sweeps as implemented "
                                 "(24 B per cell and sweep + RHS and update passes)"}}))
    if world > 1:
        dist.destroy_process_group()


def run_slab(args, rank, world, local):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = slab_measure(args.workload, args.steps, args.warmup, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _events_ms(fn):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


DMMA_PEAK_TFLOPS = 37.09     # measured on B200: scripts/micro/dmma_bench.cu -> profiles/r2_micro_dmma_bench.txt


def extra_configs():
    """The single-simulation configs of BASELINE.json (configs[0..2]) device-resident and device-timed (CUDA events, no
    trajectory copy), each with its roofline fraction and a CPU baseline (oracle port, bounded sample).  N = 1 only."""
    import torch
    import nns_b200
    from nns_b200.ensemble import ChorinEnsemble, DirectEnsemble, SpectralEnsemble, cavity_bcs
    from oracle import fd as ofd
    ofd.build()
    peak, _ = peaks()
    out = {}
    # ---- config 1: chorin_fd cavity 41 x 41, nt = 500, nit = 50 (one 13 KiB problem: latency-bound, one SM)
    try:
        nx = ny = 41
        u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
        ens = ChorinEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=1e-3, rho=1, nu=0.1, beta=1.25,
                             method="explicit")
        ens.init_variables()
        ens.run(50)
        ms = _events_ms(lambda: ens.run(500)) / 500
        z = np.zeros((nx, ny))
        ofd.set_threads(1)
        t0 = time.perf_counter()
        ofd.chorin_simulate(z, z, z, u_bc, v_bc, p_bc, nt=500, nit=50, dt=1e-3, rho=1, nu=0.1, beta=1.25, method="explicit")
        tc = time.perf_counter() - t0
        ach = BYTES_PER_CELL_UPDATE * nx * ny / (ms * 1e-3) / 1e9
        out["chorin_fd_cavity41"] = {
            "config": "chorin_fd cavity 41x41, nit=50, dt=1e-3, explicit, steps 50..550 of the run (BASELINE configs[0])",
            "ms_per_step": ms, "value": nx * ny / (ms * 1e-3), "unit": UNIT, "steps": 500,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "note": "one 13 KiB problem on one SM: latency-bound, not a roofline config"},
            "cpu_baseline": {"value": nx * ny * 500 / tc, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "500 steps in %.2f s (oracle/oracle.c)" % tc}}
        del ens
    except Exception as exc:
        out["chorin_fd_cavity41"] = {"error": repr(exc)}
    # ---- config 2a: direct_fd cavity 256 x 256, nt = 2000, nit = 50, dt = 1e-4 (stable; the shipped 1e-3 is not)
    try:
        nx = ny = 256
        u_bc, v_bc, p_bc = cavity_bcs(2. / (nx - 1), 2. / (ny - 1))
        ens = DirectEnsemble(1, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=50, dt=1e-4, rho=1, nu=0.1)
        ens.run(20)
        l0 = ens.launches
        ms = _events_ms(lambda: ens.run(2000)) / 2000
        launches = ens.launches - l0
        z = np.zeros((nx, ny))
        t0 = time.perf_counter()
        ofd.direct_simulate(z.copy(), z.copy(), z.copy(), u_bc, v_bc, p_bc, nt=40, nit=50, dt=1e-4, rho=1, nu=0.1)
        tc = time.perf_counter() - t0
        ach = 48 * nx * ny / (ms * 1e-3) / 1e9
        out["direct_fd_cavity256"] = {
            "config": "direct_fd cavity 256x256, nt=2000, nit=50 Jacobi sweeps, dt=1e-4 (BASELINE configs[1], parity variant 2a)",
            "ms_per_step": ms, "value": nx * ny / (ms * 1e-3), "unit": UNIT, "steps": 2000, "gpu_launches": int(launches),
            "finite": bool(torch.isfinite(ens.u).all() and torch.isfinite(ens.p).all()),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "bytes_per_cell_update": 48,
                         "note": "one 1.5 MB problem on a 16-CTA thread-block cluster (p resident in shared memory, band edges by "
                                 "st.async through distributed shared memory): bound by the shared-memory bandwidth of 16 SMs and "
                                 "the per-sweep hand-off, not by HBM"},
            "cpu_baseline": {"value": nx * ny * 40 / tc, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "40 steps in %.2f s (oracle/oracle.c)" % tc}}
        del ens
    except Exception as exc:
        out["direct_fd_cavity256"] = {"error": repr(exc)}
    # ---- config 2b: the same grid as a channel, periodic in the differenced axis 1 with a body force (an extension: the
    # reference has neither; reference-unpinned, oracle = numpy restatement with wrapped indices)
    try:
        nx = ny = 256
        D, Nm = nns_b200.DirichletBoundaryCondition, nns_b200.NeumannBoundaryCondition
        dx = dy = 2. / (nx - 1)
        walls = lambda cls: [cls(0.0, 'left', dx, dy), cls(0.0, 'right', dx, dy)]     # noqa: E731
        ens = DirectEnsemble(1, nx, ny, u_bc=walls(D), v_bc=walls(D), p_bc=walls(Nm), nit=50, dt=1e-4, rho=1, nu=0.1,
                             periodic_x=True, force_x=1.0)
        ens.run(20)
        ms = _events_ms(lambda: ens.run(2000)) / 2000
        z = np.zeros((nx, ny))
        t0 = time.perf_counter()
        ofd.direct_periodic_simulate(z, z, z, walls(D), walls(D), walls(Nm), 3, 50, 1e-4, 1.0, 0.1, 1.0)
        tc = (time.perf_counter() - t0) / 3
        ach = 48 * nx * ny / (ms * 1e-3) / 1e9
        out["direct_fd_channel256_periodic"] = {
            "config": "direct_fd channel 256x256, periodic x-BC + body force, nt=2000, nit=50, dt=1e-4 (BASELINE configs[1] as named; "
                      "extension, reference-unpinned)",
            "ms_per_step": ms, "value": nx * ny / (ms * 1e-3), "unit": UNIT, "steps": 2000,
            "finite": bool(torch.isfinite(ens.u).all() and torch.isfinite(ens.p).all()), "u_mean": float(ens.u.mean()),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes_per_cell_update": 48},
            "cpu_baseline": {"value": nx * ny / tc, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "3 steps of the numpy restatement (oracle/fd.py direct_periodic_simulate), %.0f ms per step" % (tc * 1e3)}}
        del ens
    except Exception as exc:
        out["direct_fd_channel256_periodic"] = {"error": repr(exc)}
    # ---- config 3: chorin_spectral, N = 127 (odd N: real spectrum), 1000 steps; the reference scheme overflows within
    # ~10 steps: inf / NaN do not change fp64 GEMM timing -- timing only
    try:
        N = 127
        D = nns_b200.DirichletBoundaryCondition
        dxs = 2. / (N - 1.)
        u_bc = [D(0, 'left', dxs, dxs), D(1, 'right', dxs, dxs), D(0, 'top', dxs, dxs), D(0, 'bottom', dxs, dxs)]
        v_bc = [D(0, s, dxs, dxs) for s in ('left', 'right', 'top', 'bottom')]
        flops = 2.0 * 28 * (N - 2) ** 3
        res = {}
        for B, steps in ((1, 1000), (1024, 30)):
            ens = SpectralEnsemble(B, N, N, u_bc=u_bc, v_bc=v_bc, dt=1e-3, rho=1)
            ens.set_state(*[np.zeros((B, N, N))] * 3)
            ens.run(9)
            ms = _events_ms(lambda: ens.run(steps)) / steps
            tf = flops * B / (ms * 1e-3) / 1e12
            res[B] = {"members": B, "steps": steps, "ms_per_step": ms, "value": B * N * N / (ms * 1e-3), "unit": UNIT,
                      "roofline": {"bound": "tensor", "achieved": tf, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                                   "frac": tf / DMMA_PEAK_TFLOPS,
                                   "note": "fp64 mma.sync m8n8k4 (DMMA) peak measured by scripts/micro/dmma_bench.cu; "
                                           "28 products of 125^3 per member-step"}}
            del ens
        from oracle import spectral as osp
        S = osp.Setup(N, N, u_bc, v_bc)
        rng = np.random.default_rng(0)
        st = [1e-3 * rng.standard_normal((N, N)) for _ in range(5)]
        t0 = time.perf_counter()
        nrep = 10
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(nrep):
                ui, vi = osp.predictor(S, 1e-3, st[0], st[1], st[2], st[3])
                osp.correction(S, 1e-3, 1, ui, vi, st[4])
        tc = (time.perf_counter() - t0) / nrep
        out["chorin_spectral127"] = {
            "config": "chorin_spectral N=127 (Chebyshev collocation, 28 dense 125^3 fp64 products per step), dt=1e-3 (BASELINE configs[2]); "
                      "values overflow after a few steps as in the reference: timing only",
            "single": res[1], "ensemble1024": res[1024],
            "cpu_baseline": {"value": N * N / tc, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d steps of oracle/spectral.py (numpy / BLAS threads), %.1f ms per step" % (nrep, tc * 1e3)}}
    except Exception as exc:
        out["chorin_spectral127"] = {"error": repr(exc)}
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ensemble4096_cavity128",
                    choices=sorted(WORKLOADS) + sorted(SLAB_WORKLOADS) + sorted(DIRECT_SLAB_WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--members", type=int, default=None, help="override members per GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (extra.configs / extra.slab16384)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in DIRECT_SLAB_WORKLOADS:
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "slab workloads have no CPU arm; use the default workload"}))
            return
        run_direct_slab(args, rank, world, local)
        return
    if args.workload in SLAB_WORKLOADS:
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "slab workloads have no CPU arm; use the default workload"}))
            return
        run_slab(args, rank, world, local)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from nns_b200.ensemble import (ChorinEnsemble, cavity_bc_values, cavity_bcs, cavity_ensemble_params)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    nx, ny, B, nit, dt = WORKLOADS[args.workload]
    if args.members:
        B = args.members
    lid, nu = cavity_ensemble_params(B * world, seed=0, lo=rank * B, hi=(rank + 1) * B)
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = ChorinEnsemble(B, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=nit, dt=dt, rho=1, nu=nu, beta=1.25,
                         method="explicit", bc_values=cavity_bc_values(lid))
    ens.init_variables()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ens.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ens.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        ens.step()                       # one kernel launch on torch's current stream
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ens.launches - l0
    ms_total = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    sweeps_last = ens.sweeps.cpu().numpy()
    finite = bool(torch.isfinite(ens.u).all() and torch.isfinite(ens.p).all())
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cells = B * world * nx * ny
    value = cells * args.steps / (ms_total * 1e-3)

    # ---- end to end: host-buffer C-ABI step (H2D 5 fields, D2H 3 fields per step) ----------
    from nns_b200 import _lib
    e2e = None
    try:
        if args.e2e_steps <= 0:
            raise RuntimeError("skipped (--e2e-steps 0)")
        hs = [torch.zeros((B, nx, ny), dtype=torch.float64).pin_memory() for _ in range(7)]
        hsw = torch.zeros((B,), dtype=torch.int32).pin_memory()
        for k, src in enumerate((ens.u, ens.v, ens.u1, ens.v1, ens.p)):
            hs[k].copy_(src)
        torch.cuda.synchronize()
        L = _lib.lib()

        def host_step():
            _lib.check(L.nns_chorin_fd_step_host(ens.handle.h, hs[0].data_ptr(), hs[1].data_ptr(), hs[2].data_ptr(),
                                                 hs[3].data_ptr(), hs[4].data_ptr(), hs[5].data_ptr(),
                                                 hs[6].data_ptr(), hsw.data_ptr()))
            hs[2], hs[0], hs[5] = hs[0], hs[5], hs[2]          # rotate host roles: u1 <- u <- u_out
            hs[3], hs[1], hs[6] = hs[1], hs[6], hs[3]
        host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        fb = B * nx * ny * 8
        e2e = {"value": cells * args.e2e_steps / el, "unit": UNIT, "h2d_bytes_per_step": 5 * fb * world,
               "d2h_bytes_per_step": 3 * fb * world + 4 * B * world, "steps": args.e2e_steps,
               "api": "nns_chorin_fd_step_host (pinned host buffers, chunked copy/compute pipeline)"}
        del hs
    except Exception as exc:      # report, never fake
        e2e = {"value": None, "unit": UNIT, "error": str(exc)}

    # ---- BASELINE configs[4] on the same ranks (N >= 2): one 16384^2 grid on row slabs, 5 timed steps (strong scaling)
    slab_extra = None
    if world > 1 and not args.no_extras and not args.members:
        del ens
        torch.cuda.empty_cache()
        try:
            r = slab_measure("slab_cavity16384", 5, 2, rank, world, local, with_clocks=False)
            if rank == 0:
                ref_ms = 88.3      # one GPU, profiles/r1_slab16384_n1_bench.json (the N = 1 run does not fit this run's time budget)
                slab_extra = {"ms_per_step": r["ms_per_step"], "value": r["value"], "unit": UNIT, "scaling": "strong",
                              "steps": 5, "ticks": r["config"]["ticks_per_step"], "exchange_mode": r["config"]["exchange_mode"],
                              "efficiency_vs_n1": ref_ms / r["ms_per_step"] / world, "n1_ms_per_step": ref_ms,
                              "roofline": r["roofline"], "finite": r["config"]["finite"]}
        except Exception as exc:
            if rank == 0:
                slab_extra = {"error": repr(exc)}
    if rank == 0:
        peak, peak_src = peaks()
        kms = float(np.mean(kern_ms))
        achieved = BYTES_PER_CELL_UPDATE * B * nx * ny / (kms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, nx, ny, B, world, nit, dt),
            "checks": {"sweeps_per_step_last": [int(sweeps_last.min()), int(sweeps_last.max())], "finite": finite},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "chorin_stream_kernel", "bytes_per_cell_update": BYTES_PER_CELL_UPDATE,
                         "kernel_ms": kms,
                         "note": "one launch = one fused step of all members; FP64-pipe / shared-memory bound before "
                                 "HBM (49 exact-order SOR sweeps x 6 FP64 instr per cell: FP64 floor 1.2-1.4 ms/step "
                                 "vs HBM floor 0.66 ms); see DESIGN.md 4.1"},
        }
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    out["roofline"]["traffic"] = json.load(f).get(args.workload)
            except Exception:
                pass
        if world == 1 and not args.no_extras and not args.members:
            out["extra"] = {"configs": extra_configs()}
        if world > 1 and slab_extra is not None:
            out["extra"] = {"slab16384": slab_extra}
        if world == 1 and not args.no_cpu_baseline:
            val, used, members, el, _ = cpu_port_throughput(nx, ny, nit, dt, budget_s=12.0)
            out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                                   "sample": "%d of %d members x 1 step in %.1f s (oracle/oracle.c, bit-exact "
                                             "restatement of the reference, OpenMP over members)" % (members, B, el)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
