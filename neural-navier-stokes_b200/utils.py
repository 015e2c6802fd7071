"""Drop-in for the data helpers of src/utils.py that sit on the time-step path's output: ``spatial_coarsen``
(src/utils.py:13-60) and ``numpy_to_torch`` (src/utils.py:9-10).  The block means run on the GPU
(``nns_traj_coarsen``); the ML helpers of that file (AverageMeter, checkpoints, ...) are out of scope."""
import numpy as np
import torch

from . import trajectory


def numpy_to_torch(array, device):
    return torch.from_numpy(array).float().to(device)


def spatial_coarsen(X, Y, u_seq, v_seq, p_seq, agg_x=4, agg_y=4):
    """Same signature, return values and quirks as the reference: ``(new_X, new_Y, new_u_seq, new_v_seq,
    new_p_seq)`` with ``new_X, new_Y = np.meshgrid(np.linspace(0, 2, nx // agg_x), np.linspace(0, 2, ny // agg_y))``
    and ``(T, nx // agg_x, ny // agg_y)`` float64 block means (bit-identical to np.mean over the flattened block).
    u_seq, v_seq, p_seq: numpy arrays or CUDA tensors (tensors stay on the device)."""
    nx, ny = X.shape[0], X.shape[1]
    assert nx % agg_x == 0
    assert ny % agg_y == 0
    new_x = np.linspace(0, 2, nx // agg_x)
    new_y = np.linspace(0, 2, ny // agg_y)
    new_X, new_Y = np.meshgrid(new_x, new_y)
    on_host = isinstance(u_seq, np.ndarray)
    if not torch.cuda.is_available():
        raise RuntimeError("nns_b200.utils.spatial_coarsen needs a CUDA device (no CPU fallback)")
    dev = torch.device('cuda', torch.cuda.current_device())
    seqs = [torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).to(dev) if isinstance(s, np.ndarray) else s
            for s in (u_seq, v_seq, p_seq)]
    out = trajectory.coarsen_device(*seqs, agg_x=agg_x, agg_y=agg_y)
    if on_host:
        out = tuple(trajectory.to_host(o) for o in out)
    return (new_X, new_Y) + tuple(out)
