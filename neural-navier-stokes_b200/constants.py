"""Data-file locations, same names and values as the reference's ``src/constants.py:1-5``
(consumers such as ``src/neural_spectral/*`` read the solver output from these paths)."""
import os

SRC_DIR = os.path.realpath(os.path.dirname(__file__))
DATA_DIR = os.path.join(SRC_DIR, 'data')
CHORIN_FD_DATA_FILE = os.path.join(DATA_DIR, 'chorin_fd', 'data_semi_implicit.npz')
DIRECT_FD_DATA_FILE = os.path.join(DATA_DIR, 'direct_fd', 'data.npz')
