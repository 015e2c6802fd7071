#!/usr/bin/env python
"""Golden fixture of the trajectory sink, made by RUNNING THE REFERENCE's own ``utils.spatial_coarsen``
(src/utils.py:13-60) in the build container:

    PYTHONPATH=/root/repo python tests/golden/make_golden_traj.py

Inputs are seeded smooth + noisy fields (written into the fixture so that the GPU box needs nothing else); outputs
are the reference's coarsened sequences for (agg_x, agg_y) = (4, 4), (2, 2), (4, 2) [the ``ny // agg_x`` loop bound
of utils.py:49 leaves the right half of the output zero] and (3, 3) [block of 9: NumPy's 8-accumulator pass plus a
tail].  The script also checks the oracle restatement (oracle/traj.py) against the reference -- bit-exact -- and
records that in MANIFEST.json.  /root/reference does not exist on the GPU box: only this script reads it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import traj as otraj  # noqa: E402
from src import utils as rutils  # noqa: E402  (the reference)


def fields(T, nx, ny, seed):
    rng = np.random.default_rng(seed)
    x, y = np.meshgrid(np.linspace(0, 2, nx), np.linspace(0, 2, ny), indexing="ij")
    out = []
    for k in range(3):
        a = np.stack([np.sin((k + 1) * x + 0.1 * t) * np.cos((k + 2) * y - 0.05 * t) * (10.0 ** (k - 1)) for t in range(T)])
        out.append(a + 1e-3 * rng.standard_normal((T, nx, ny)))
    return out


def main():
    cases = {"a": (5, 24, 24, 4, 4), "b": (3, 24, 24, 2, 2), "c": (3, 16, 16, 4, 2), "d": (2, 18, 18, 3, 3)}
    save, pins = {}, {}
    for name, (T, nx, ny, ax, ay) in cases.items():
        u, v, p = fields(T, nx, ny, 100 + ord(name))
        X, Y = np.meshgrid(np.linspace(0, 2, nx), np.linspace(0, 2, ny))
        X, Y = X.T.copy(), Y.T.copy()          # (nx, ny), like the solvers' meshes
        ref = rutils.spatial_coarsen(X, Y, u, v, p, agg_x=ax, agg_y=ay)
        ora = otraj.spatial_coarsen(X, Y, u, v, p, agg_x=ax, agg_y=ay)
        diffs = [float(np.max(np.abs(a - b))) for a, b in zip(ref, ora)]
        # the written-out addition order against np.mean on every block of the first field
        pw = np.zeros_like(ref[2])
        for t in range(T):
            for i in range(nx // ax):
                for j in range(ny // ax):
                    pw[t, i, j] = otraj.pairwise_mean(u[t, i * ax:(i + 1) * ax, j * ay:(j + 1) * ay].reshape(-1))
        diffs.append(float(np.max(np.abs(pw - ref[2]))))
        assert max(diffs) == 0.0, (name, diffs)
        pins[name] = {"T": T, "nx": nx, "ny": ny, "agg_x": ax, "agg_y": ay, "max_abs_diff_oracle_vs_reference": max(diffs)}
        save.update({name + "_u": u, name + "_v": v, name + "_p": p, name + "_cfg": np.array([T, nx, ny, ax, ay]),
                     name + "_X": ref[0], name + "_Y": ref[1], name + "_cu": ref[2], name + "_cv": ref[3], name + "_cp": ref[4]})
    np.savez_compressed(os.path.join(HERE, "traj_coarsen.npz"), **save)
    mpath = os.path.join(HERE, "MANIFEST.json")
    man = json.load(open(mpath))
    man["traj_coarsen"] = {"reference": "src/utils.py:13-60 (spatial_coarsen), run unmodified", "cases": pins,
                           "numpy": np.__version__}
    json.dump(man, open(mpath, "w"), indent=1, sort_keys=True)
    print(json.dumps(pins, indent=1))


if __name__ == "__main__":
    main()
