"""CPU: the oracle of the trajectory sink (oracle/traj.py) against the fixture recorded from the reference's own
utils.spatial_coarsen (tests/golden/make_golden_traj.py), and the host-side checks of the drop-in wrapper."""
import os

import numpy as np
import pytest

from tests._util import ROOT, has_gpu, manifest

from oracle import traj as otraj

CASES = "abcd"


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "traj_coarsen.npz"))


def test_manifest_records_bit_exact_pin():
    m = manifest()["traj_coarsen"]
    assert set(m["cases"]) == set(CASES)
    for k, v in m["cases"].items():
        assert v["max_abs_diff_oracle_vs_reference"] == 0.0, k


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_coarsen_equals_reference_fixture(name):
    g = _golden()
    T, nx, ny, ax, ay = (int(x) for x in g[name + "_cfg"])
    X, Y = np.meshgrid(np.linspace(0, 2, nx), np.linspace(0, 2, ny))
    out = otraj.spatial_coarsen(X.T, Y.T, g[name + "_u"], g[name + "_v"], g[name + "_p"], agg_x=ax, agg_y=ay)
    for got, key in zip(out, ("_X", "_Y", "_cu", "_cv", "_cp")):
        assert got.shape == g[name + key].shape
        assert np.array_equal(got, g[name + key]), key
    if ax > ay:      # the ny // agg_x loop bound (utils.py:49) leaves the remaining output columns zero
        assert np.all(out[2][:, :, ny // ax:] == 0.0) and np.any(out[2][:, :, :ny // ax] != 0.0)


@pytest.mark.parametrize("n", [1, 2, 4, 7, 8, 9, 16, 24, 30, 64, 100, 128])
def test_written_out_summation_order_is_numpys(n):
    rng = np.random.default_rng(n)
    for scale in (1.0, 1e8, 1e-8):
        a = rng.standard_normal((50, n)) * scale + rng.standard_normal((50, 1))
        want = np.mean(a, axis=1)
        got = np.array([otraj.pairwise_mean(r) for r in a])
        assert np.array_equal(got, want)


def test_observations_layout():
    rng = np.random.default_rng(0)
    u, v, p = (rng.standard_normal((4, 5, 6)) for _ in range(3))
    o = otraj.observations(u, v, p)
    assert o.shape == (4, 3, 5, 6) and o.dtype == np.float32
    assert np.array_equal(o[:, 1], v.astype(np.float32))


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_wrapper_has_no_cpu_fallback():
    from nns_b200 import utils
    X = np.zeros((8, 8))
    with pytest.raises(AssertionError):                      # utils.py:39-40
        utils.spatial_coarsen(X, X, np.zeros((1, 8, 8)), np.zeros((1, 8, 8)), np.zeros((1, 8, 8)), agg_x=3, agg_y=4)
    with pytest.raises(RuntimeError):
        utils.spatial_coarsen(X, X, np.zeros((1, 8, 8)), np.zeros((1, 8, 8)), np.zeros((1, 8, 8)))
