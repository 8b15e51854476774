"""Tension input generators — drop-in for knode_cosserat/physics_controls.py (reference :3-33).

Host-side input synthesis only (no arithmetic of the hot path lives here).  `sine`, `step` and `random` reproduce the
reference's sequences exactly (same numpy RNG seeding); the reference's `ramp` branch reads an undefined name and can
only raise, so here it raises the same NameError-free, explicit Exception.
"""
import numpy as np


def calc_controls(control_type, control_arg, del_t, train_len):
    np.random.seed(int(control_arg))  # For random trajectory (physics_controls.py:4)
    controls = []
    for i in range(1, train_len + 1):
        if control_type == 'sine':
            sin_period = control_arg / del_t
            phase = 2 * np.pi / 4
            T1, T2, T3, T4 = (6 + np.sin(2 * np.pi * i / sin_period + k * phase) for k in range(4))
        elif control_type == 'step':
            step_tension = 0 if i * del_t < 1.5 else control_arg
            T1, T2, T3, T4 = 5 + step_tension, 5, 5, 5 + step_tension
        elif control_type == 'random':
            T1 = 5 + 5 * np.random.rand()
            T2 = 5 + 5 * np.random.rand()
            T3 = 5 + 5 * np.random.rand()
            T4 = 5 + 5 * np.random.rand()
        elif control_type == 'ramp':
            raise Exception("control type 'ramp' is broken in the reference (ramp_speed is undefined, "
                            "physics_controls.py:25-29)")
        else:
            raise Exception('Unknown control type ' + control_type)
        controls.append([T1, T2, T3, T4])
    return controls


def synthetic_tensions(B, T, del_t, seed=0, dtype=np.float32):
    """The C2/C5 benchmark inputs of SURVEY.md §8(d): half `sine` (period ~ U(0.5, 3.0) s, random phase, the four
    tendons 90 degrees apart as in calc_controls), half `random` 5 + 5*U(0,1) held per step.  -> [B, T, 4]."""
    rng = np.random.default_rng(seed)
    ctl = np.empty((B, T, 4), dtype)
    h = B // 2
    i = np.arange(1, T + 1)[None, :, None]
    per = rng.uniform(0.5, 3.0, (h, 1, 1)) / del_t
    ph = rng.uniform(0, 2 * np.pi, (h, 1, 1))
    ctl[:h] = 6 + np.sin(2 * np.pi * i / per + ph + np.arange(4)[None, None, :] * np.pi / 2)
    ctl[h:] = 5 + 5 * rng.random((B - h, T, 4))
    return ctl
