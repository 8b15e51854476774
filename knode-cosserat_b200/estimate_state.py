"""Drop-in for knode_cosserat_realworld/estimate_state.py — `estimate_state(data, tensions, robot)` (:158-242): the full
25-row state [p, h, n, m, q, w, v, u] on the N-node grid from measured positions and quaternions, the step that produces
the `*_estimated.npy` datasets train_segment.py trains on (:275-280).

One call = kc_estimate_state (csrc/kc_estimate.cu) on the GPU: finite differences in time, quaternion -> R, spatial
derivative of R, the backward recursion for n and m, the constitutive re-estimate of v and u with its BDF2 recurrence.
No CPU fallback.  Extensions (keyword-only / by shape, not in the reference): a leading batch dimension
(data[B,T,7,N], tensions[B,T,4] -> [B,T,25,N]) for many recordings at once, CUDA torch tensors in -> CUDA tensor out,
`dtype=np.float32`.

The reference's helper functions (compute_v_u, compute_angular_velocities, compute_internal_forces_and_moments,
compute_R_spatial_derivative) have no caller other than estimate_state and are not exported separately; the `__main__`
block of the reference (rosbag-derived inputs, interpolate_curve.fit_curve) is outside the path.
"""
from __future__ import annotations

import numpy as np
import torch

import _kc
import _ops


def estimate_state(data, tensions, robot, *, dtype=np.float64):
    """data [T,7,n] (x, y, z and the quaternion (w,x,y,z) at the robot's N nodes), tensions [T,4], robot (a CosseratRod /
    CosseratRodTorch with its derived terms computed) -> estimated_state [T,25,N] (float64 numpy, as the reference)."""
    if not torch.cuda.is_available():
        raise RuntimeError("knode-cosserat_b200 has no CPU fallback: estimate_state() needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device())
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    on_device = torch.is_tensor(data) and data.is_cuda
    d = data if torch.is_tensor(data) else torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64)))
    c = tensions if torch.is_tensor(tensions) else torch.from_numpy(np.ascontiguousarray(np.asarray(tensions, dtype=np.float64)))
    single = d.ndim == 3
    if single:
        d, c = d[None], c[None]
    if d.ndim != 4 or d.shape[2] != 7:
        raise ValueError(f"data must be [T,7,N] or [B,T,7,N], got {tuple(d.shape)}")
    if d.shape[3] != int(robot.N):
        # the reference assigns data into a [T,25,robot.N] array and fails on a mismatch (:174)
        raise ValueError(f"data has {d.shape[3]} nodes, robot.N = {int(robot.N)}")
    if d.shape[1] < 3:
        raise ValueError("Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) "
                         "elements are required.")                           # numpy.gradient's own message (:186)
    c = c[:, :d.shape[1]]                                                   # the reference only reads tensions[t], t < T
    est = _ops.estimate_state(_kc.rod_params(robot), float(robot.L), float(robot.del_t), d.to(dev, tdt), c.to(dev, tdt))
    if single:
        est = est[0]
    if on_device:
        return est
    out = est.cpu().numpy()
    if single:
        robot.vstar = out[0, 19:22, 0]    # side effect of the reference (:202): a view of the result's root v at t = 0
    return out
