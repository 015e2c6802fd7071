"""GPU: the trajectory sink (trajectory.cu through nns_traj_coarsen / nns_traj_observations) against the fixture
recorded from the reference's utils.spatial_coarsen and against the oracle.  Block means are computed in NumPy's
summation order: BIT-EXACT."""
import os

import numpy as np
import pytest

from tests._util import ROOT

from oracle import traj as otraj

pytestmark = pytest.mark.gpu


def _golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "traj_coarsen.npz"))


@pytest.mark.parametrize("name", list("abcd"))
def test_spatial_coarsen_dropin_equals_reference_fixture(name):
    from nns_b200 import utils
    g = _golden()
    T, nx, ny, ax, ay = (int(x) for x in g[name + "_cfg"])
    X, Y = np.meshgrid(np.linspace(0, 2, nx), np.linspace(0, 2, ny))
    out = utils.spatial_coarsen(X.T, Y.T, g[name + "_u"], g[name + "_v"], g[name + "_p"], agg_x=ax, agg_y=ay)
    for got, key in zip(out, ("_X", "_Y", "_cu", "_cv", "_cp")):
        assert isinstance(got, np.ndarray) and got.dtype == np.float64
        assert np.array_equal(got, g[name + key]), key


def test_coarsen_ensemble_trajectory_full_size_properties():
    """[members, nt, 128, 128] device trajectories: every block mean equals np.mean of the block bit for bit on a
    sample, and coarsening commutes with member / frame slicing."""
    import torch
    from nns_b200 import trajectory
    rng = np.random.default_rng(3)
    B, T, n = 37, 5, 128
    host = [rng.standard_normal((B, T, n, n)) * s for s in (1.0, 1e3, 1e-3)]
    dev = [torch.from_numpy(h).cuda() for h in host]
    cu, cv, cp = trajectory.coarsen_device(*dev, agg_x=4, agg_y=4)
    assert cu.shape == (B, T, 32, 32)
    for h, c in zip(host, (cu, cv, cp)):
        c = c.cpu().numpy()
        for (b, t, i, j) in [(0, 0, 0, 0), (36, 4, 31, 31), (17, 2, 5, 30), (3, 1, 16, 7)]:
            assert c[b, t, i, j] == np.mean(h[b, t, 4 * i:4 * i + 4, 4 * j:4 * j + 4].reshape(-1))
    sub = trajectory.coarsen_device(dev[0][5:9, 1:3].contiguous(), dev[1][5:9, 1:3].contiguous(), dev[2][5:9, 1:3].contiguous())
    assert torch.equal(sub[0], cu[5:9, 1:3]) and torch.equal(sub[2], cp[5:9, 1:3])


def test_observation_tensor_and_ragged_factors():
    import torch
    from nns_b200 import trajectory
    rng = np.random.default_rng(4)
    u, v, p = (rng.standard_normal((7, 30, 20)) for _ in range(3))
    du, dv, dp = (torch.from_numpy(a).cuda() for a in (u, v, p))
    obs = trajectory.observations_device(du, dv, dp)
    assert obs.dtype == torch.float32 and obs.shape == (7, 3, 30, 20)
    assert np.array_equal(obs.cpu().numpy(), otraj.observations(u, v, p))
    # coarsened observations = float32 of the reference's coarsened fields (blocks of 5 x 2 = 10 cells: NumPy's
    # 8-accumulator pass plus a tail of 2)
    X = np.zeros((30, 20))
    _, _, cu, cv, cp = otraj.spatial_coarsen(X, X, u, v, p, agg_x=5, agg_y=2)
    # ny // agg_x = 4 < ny // agg_y = 10: the reference leaves the other columns zero, and so does the kernel
    obs2 = trajectory.observations_device(du, dv, dp, agg_x=5, agg_y=2)
    assert np.array_equal(obs2.cpu().numpy(), otraj.observations(cu, cv, cp))


def test_errors_mirror_the_reference():
    import torch
    from nns_b200 import trajectory
    z = torch.zeros((2, 12, 12), dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):          # utils.py:39-40
        trajectory.coarsen_device(z, z, z, agg_x=5, agg_y=4)
    with pytest.raises(IndexError):              # ny // agg_x > ny // agg_y: utils.py:55 indexes past the output
        trajectory.coarsen_device(z, z, z, agg_x=2, agg_y=4)
    with pytest.raises(RuntimeError):            # no CPU fallback
        trajectory.coarsen_device(z.cpu(), z.cpu(), z.cpu())
    e = torch.zeros((0, 12, 12), dtype=torch.float64, device="cuda")
    assert trajectory.coarsen_device(e, e, e, 4, 4)[0].shape == (0, 3, 3)      # empty trajectory


def test_ensemble_trajectory_to_data_files(tmp_path):
    """End of the path: an ensemble run leaves [members, nt, nx, ny] trajectories on the device; save_npz writes
    the reference's data-file format (keys u, v, p; (nt, nx, ny) float64) that np.load-based consumers read."""
    from nns_b200 import trajectory
    from nns_b200.ensemble import ChorinEnsemble, cavity_bcs
    nx = ny = 32
    dx = dy = 2. / (nx - 1)
    u_bc, v_bc, p_bc = cavity_bcs(dx, dy)
    ens = ChorinEnsemble(3, nx, ny, u_bc=u_bc, v_bc=v_bc, p_bc=p_bc, nit=20, dt=1e-3, rho=1, nu=0.1, beta=1.25,
                         method='explicit')
    ens.init_variables()
    tu, tv, tp, _ = ens.run(4, trajectory=True, sweeps=True)
    paths = trajectory.save_ensemble(str(tmp_path / "data"), tu, tv, tp)
    assert len(paths) == 3
    d = np.load(paths[1])
    assert sorted(d.files) == ["p", "u", "v"]
    assert d["u"].shape == (4, nx, ny) and d["u"].dtype == np.float64
    assert np.array_equal(d["p"], tp[1].cpu().numpy())
    trajectory.save_npz(str(tmp_path / "coarse.npz"), tu[2], tv[2], tp[2], agg_x=4, agg_y=4)
    c = np.load(str(tmp_path / "coarse.npz"))
    X = np.zeros((nx, ny))
    want = otraj.spatial_coarsen(X, X, tu[2].cpu().numpy(), tv[2].cpu().numpy(), tp[2].cpu().numpy())
    assert np.array_equal(c["u"], want[2]) and np.array_equal(c["p"], want[4])
