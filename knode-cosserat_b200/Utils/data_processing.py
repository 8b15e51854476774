"""Host helpers of knode_cosserat_realworld/Utils/data_processing.py (imported by train_segment.py:6, never called on the
hot path).  Same behaviour as the reference: statistics over time for 2-D data and over time and space for 3-D data
(:18-21), ranges clipped to >= 1e-10 (:29), squeezed statistics returned (:31); denormalize_data is its literal inverse
formula with the (min, max) arguments the reference's signature names (:33-50)."""
import numpy as np


def normalize_data(data):
    data = np.asarray(data)
    if data.ndim == 2:
        axis = 0
    elif data.ndim == 3:
        axis = (0, 2)
    else:
        raise ValueError("normalize_data expects 2-D (time, channel) or 3-D (time, channel, node) data")
    lo = np.min(data, axis=axis, keepdims=True)
    span = np.clip(np.max(data, axis=axis, keepdims=True) - lo, 1e-10, np.inf)
    return (data - lo) / span, lo.squeeze(), span.squeeze()


def denormalize_data(normalized_data, min_vals, max_vals):
    return normalized_data * (max_vals - min_vals) + min_vals
