"""quaternion_to_euler — drop-in for knode_cosserat/Utils/transformations.py:3-31.

Pure torch elementwise math on whatever device the input lives on (it is the loss helper of the *slow* reference loop;
the fused training step computes the same angles inside kc_train_fwd_kernel).  Kept differentiable like the original.
"""
import torch


def quaternion_to_euler(quaternions):
    """[4, a] quaternions (w, x, y, z) -> [3, a] Euler angles (roll, pitch, yaw); force-cast to fp32 as the
    reference does (:14)."""
    quaternions = quaternions.float()
    norms = quaternions.norm(p=2, dim=0, keepdim=True)
    q = quaternions / norms
    w, x, y, z = q[0], q[1], q[2], q[3]
    roll = torch.atan2(2 * (w * y + x * z), 1 - 2 * (y ** 2 + z ** 2))
    pitch = torch.asin(torch.clamp(2 * (w * z - x * y), -1.0, 1.0))
    yaw = torch.atan2(2 * (w * x + y * z), 1 - 2 * (x ** 2 + z ** 2))
    return torch.stack([roll, pitch, yaw], dim=0)
