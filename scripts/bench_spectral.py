"""chorin_spectral throughput (BASELINE config 3): the 28-GEMM step chain for N = 127 (odd N: real spectrum), device-
resident, CUDA events.  The reference scheme overflows within ~10 steps: values become inf/NaN, which does
not change fp64 GEMM timing -- timing only.  Usage: python scripts/bench_spectral.py [N] [batch ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nns_b200
from nns_b200.ensemble import SpectralEnsemble
D = nns_b200.DirichletBoundaryCondition
N = int(sys.argv[1]) if len(sys.argv) > 1 else 127
batches = [int(x) for x in sys.argv[2:]] or [1, 64, 1024]
dx = dy = 2. / (N - 1.)
u_bc = [D(0, 'left', dx, dy), D(1, 'right', dx, dy), D(0, 'top', dx, dy), D(0, 'bottom', dx, dy)]
v_bc = [D(0, s, dx, dy) for s in ('left', 'right', 'top', 'bottom')]
for B in batches:
    for mode in ("graph", "plain"):
        os.environ["NNS_SPECTRAL_NOGRAPH"] = "0" if mode == "graph" else "1"
        ens = SpectralEnsemble(B, N, N, u_bc=u_bc, v_bc=v_bc, dt=1e-3, rho=1)
        steps = 999 if B == 1 else 30
        ens.run(9)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ens.launches
        e0.record(); ens.run(steps); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flops = 2.0 * 28 * (N - 2) ** 3 * B
        print(json.dumps({"N": N, "batch": B, "mode": mode, "steps": steps, "ms_per_step": ms, "launches_per_step": (ens.launches - l0) / steps,
                          "cell_updates_per_s": B * N * N / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}))
