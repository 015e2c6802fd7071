"""Trajectory sink throughput (nns_traj_coarsen / nns_traj_observations) against the HBM roofline.
Algorithmic bytes per fine cell: 24 B read + 24 / (agg_x agg_y) B (float64) or 12 / (agg_x agg_y) B (float32) written."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nns_b200 import trajectory  # noqa: E402

peak = 6548.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
B, T, n = 4096, 8, 128          # 4096 members x 8 stored frames: 3 x 4.3 GB of trajectories (larger than L2)
tu, tv, tp = (torch.randn((B, T, n, n), dtype=torch.float64, device="cuda") for _ in range(3))
for name, fn, wbytes in (("coarsen 4x4 (float64)", lambda: trajectory.coarsen_device(tu, tv, tp, 4, 4), 24 / 16),
                         ("observations 4x4 (float32)", lambda: trajectory.observations_device(tu, tv, tp, 4, 4), 12 / 16),
                         ("observations 1x1 (float32)", lambda: trajectory.observations_device(tu, tv, tp, 1, 1), 12)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    cells = B * T * n * n
    gbs = cells * (24 + wbytes) / ms / 1e6
    print(json.dumps({"kernel": "traj_pack_kernel", "case": name, "ms": ms, "cells_per_s": cells / ms * 1e3,
                      "achieved_GBs": gbs, "peak_GBs": peak, "frac": gbs / peak}))
