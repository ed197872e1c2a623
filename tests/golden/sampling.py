"""Shared by gen_fullsize.py (writer) and the tests (readers): which elements of a big tensor a golden file holds."""
import numpy as np

N_TAP, N_OUT = 10000, 100000


def sub_idx(key, numel, n):
    """Sample positions are a pure function of (tensor name, size): regenerated on the test side, not stored."""
    seed = sum((i + 1) * ord(ch) for i, ch in enumerate(key)) % (2 ** 31)
    return np.sort(np.random.RandomState(seed).choice(numel, size=min(n, numel), replace=False)).astype(np.int64)
