"""Trajectory sink: the interface between the time-step path and the rest of the reference.

The solvers leave trajectories on the device (``[members, nt, nx, ny]`` float64 u, v, p).  The reference's
consumers want (a) ``np.savez(path, u=..., v=..., p=...)`` files with ``(nt, nx, ny)`` float64 arrays
(writers: src/chorin_fd/simulate.py:323-324, src/direct_fd/simulate.py:193-194; paths: src/constants.py:3-5),
(b) spatially coarsened copies of them (src/utils.py:13-60) and (c) the float32 ``(nt, 3, nx, ny)`` observation
tensor the neural scripts build (src/neural_spectral/rnn.py:77-82, spectral_ode.py:158-163).  (b) and (c) run
on the device through ``nns_traj_coarsen`` / ``nns_traj_observations`` so that only the (16x smaller) result
crosses PCIe; there is no CPU fallback.
"""
import os

import numpy as np
import torch

from . import _lib


def _check3(u, v, p):
    if not (u.is_cuda and v.is_cuda and p.is_cuda):
        raise RuntimeError("nns_b200.trajectory works on CUDA tensors (no CPU fallback)")
    if not (u.dtype == v.dtype == p.dtype == torch.float64):
        raise TypeError("trajectories must be float64")
    if not (u.shape == v.shape == p.shape) or u.dim() < 3:
        raise ValueError("u, v, p must have the same shape [..., nx, ny]")
    return u.contiguous(), v.contiguous(), p.contiguous()


def coarsen_device(u, v, p, agg_x=4, agg_y=4):
    """Block means of CUDA trajectories ``[..., nx, ny]`` -> ``[..., nx // agg_x, ny // agg_y]`` (float64), the
    u/v/p part of utils.spatial_coarsen, bit-identical to it (NumPy's summation order)."""
    u, v, p = _check3(u, v, p)
    nx, ny = u.shape[-2:]
    assert nx % agg_x == 0          # utils.py:39
    assert ny % agg_y == 0          # utils.py:40
    if ny // agg_x > ny // agg_y:
        raise IndexError("index %d is out of bounds for axis 2 with size %d" % (ny // agg_y, ny // agg_y))   # utils.py:55
    frames = int(np.prod(u.shape[:-2]))
    shape = tuple(u.shape[:-2]) + (nx // agg_x, ny // agg_y)
    out = [torch.empty(shape, dtype=torch.float64, device=u.device) for _ in range(3)]
    if frames == 0:
        return tuple(out)
    with torch.cuda.device(u.device):
        st = torch.cuda.current_stream(u.device).cuda_stream
        _lib.check(_lib.lib().nns_traj_coarsen(u.data_ptr(), v.data_ptr(), p.data_ptr(), frames, nx, ny, agg_x, agg_y,
                                               out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), st))
    return tuple(out)


def observations_device(u, v, p, agg_x=1, agg_y=1):
    """float32 observation tensor ``[..., nt, 3, nx', ny']`` of CUDA trajectories ``[..., nt, nx, ny]``:
    ``torch.stack([u, v, p]).permute(1, 0, 2, 3)`` of the ``.float()`` fields (rnn.py:77-82), optionally coarsened."""
    u, v, p = _check3(u, v, p)
    nx, ny = u.shape[-2:]
    assert nx % agg_x == 0 and ny % agg_y == 0
    frames = int(np.prod(u.shape[:-2]))
    out = torch.empty(tuple(u.shape[:-2]) + (3, nx // agg_x, ny // agg_y), dtype=torch.float32, device=u.device)
    if frames == 0:
        return out
    with torch.cuda.device(u.device):
        st = torch.cuda.current_stream(u.device).cuda_stream
        _lib.check(_lib.lib().nns_traj_observations(u.data_ptr(), v.data_ptr(), p.data_ptr(), frames, nx, ny, agg_x, agg_y,
                                                    out.data_ptr(), st))
    return out


def to_host(t):
    """Device tensor -> numpy through a pinned staging buffer (one DMA, no pageable bounce)."""
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def save_npz(path, u, v, p, agg_x=1, agg_y=1):
    """Write one simulation's trajectory in the reference's data-file format: ``np.savez(path, u=, v=, p=)`` with
    ``(nt, nx, ny)`` float64 arrays (chorin_fd/simulate.py:323-324), optionally coarsened on the device first."""
    if u.dim() != 3:
        raise ValueError("save_npz writes one simulation: tensors must be [nt, nx, ny]")
    if agg_x != 1 or agg_y != 1:
        u, v, p = coarsen_device(u, v, p, agg_x, agg_y)
    d = os.path.dirname(os.path.abspath(path))
    if not os.path.isdir(d):
        os.makedirs(d)
    np.savez(path, u=to_host(u), v=to_host(v), p=to_host(p))


def save_ensemble(out_dir, tu, tv, tp, agg_x=1, agg_y=1, pattern="member_{:05d}.npz"):
    """One data file per ensemble member from ``[members, nt, nx, ny]`` device trajectories (the training-data
    generation of BASELINE config 4)."""
    paths = []
    for b in range(tu.shape[0]):
        paths.append(os.path.join(out_dir, pattern.format(b)))
        save_npz(paths[-1], tu[b], tv[b], tp[b], agg_x, agg_y)
    return paths
