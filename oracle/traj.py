"""ORACLE (test infrastructure only -- never imported by the product path): numpy restatement of the trajectory
sink's arithmetic.

``spatial_coarsen`` follows src/utils.py:13-60 of the reference line by line (including the ``ny // agg_x`` bound
of the inner loop, :49).  ``pairwise_mean`` spells out the order in which ``np.mean`` adds the elements of a
flattened block (NumPy's pairwise_sum: eight running accumulators, then a fixed tree, then one division), which is
the order the CUDA kernel uses; ``tests/test_oracle_golden.py`` pins both against the reference's own function run
in the build container (tests/golden/traj_coarsen.npz, made by tests/golden/make_golden_traj.py).
``observations`` restates src/neural_spectral/rnn.py:77-82.
"""
import numpy as np


def spatial_coarsen(X, Y, u_seq, v_seq, p_seq, agg_x=4, agg_y=4):
    nx, ny = X.shape[0], X.shape[1]
    T = u_seq.shape[0]
    assert nx % agg_x == 0
    assert ny % agg_y == 0
    new = [np.zeros((T, nx // agg_x, ny // agg_y)) for _ in range(3)]
    new_x = np.linspace(0, 2, nx // agg_x)
    new_y = np.linspace(0, 2, ny // agg_y)
    new_X, new_Y = np.meshgrid(new_x, new_y)
    for i in range(nx // agg_x):
        for j in range(ny // agg_x):                     # sic: agg_x (utils.py:49)
            for out, seq in zip(new, (u_seq, v_seq, p_seq)):
                sub = seq[:, i * agg_x:(i + 1) * agg_x, j * agg_y:(j + 1) * agg_y].reshape(T, -1)
                out[:, i, j] = np.mean(sub, axis=1)
    return (new_X, new_Y) + tuple(new)


def pairwise_mean(block):
    """np.mean of a 1-D float64 array of at most 128 elements, with every addition written out."""
    a = np.asarray(block, dtype=np.float64)
    n = a.size
    if n < 8:
        res = np.float64(0.0)
        for x in a:
            res = res + x
    else:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
    return res / np.float64(n) if n != 1 else res


def observations(u, v, p):
    """(nt, 3, nx, ny) float32: torch.stack([u, v, p]).permute(1, 0, 2, 3) of the .float() fields."""
    return np.stack([u.astype(np.float32), v.astype(np.float32), p.astype(np.float32)]).transpose(1, 0, 2, 3)
