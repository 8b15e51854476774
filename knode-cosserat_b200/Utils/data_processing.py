"""normalize_data — imported (never called) by the reference's train_segment.py:6; kept so that import works."""
import numpy as np


def normalize_data(data):
    data = np.asarray(data)
    lo, hi = data.min(axis=0), data.max(axis=0)
    rng = np.where(hi - lo == 0, 1.0, hi - lo)
    return (data - lo) / rng, lo, rng
