import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def manifest():
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        return json.load(f)


def make_bcs(spec, nx, ny):
    """[[side, type, value], ...] (golden fixture format) -> nns_b200 BC objects."""
    import nns_b200
    dx, dy = 2. / (nx - 1), 2. / (ny - 1)
    cls = {"dirichlet": nns_b200.DirichletBoundaryCondition, "neumann": nns_b200.NeumannBoundaryCondition}
    return [cls[t](v, s, dx, dy) for s, t, v in spec]


def rel_l2(a, b):
    d = np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())
    n = np.linalg.norm(np.asarray(b).ravel())
    return d / n if n > 0 else d


def smooth_ic(nx, ny, seed, amp=0.3):
    rng = np.random.default_rng(seed)
    x = np.linspace(-1, 1, nx)[:, None]
    y = np.linspace(-1, 1, ny)[None, :]
    out = []
    for _ in range(3):
        f = np.zeros((nx, ny))
        for _k in range(4):
            a, b, c, d = rng.normal(size=4)
            f += a * np.sin(np.pi * (b * x + c * y) + d)
        out.append(amp * f / 4)
    return out


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
