"""ctypes binding of libknode_cosserat_b200.so (include/knode_cosserat.h) — the thin C-ABI layer between the
reference-shaped Python surface and the hand-written sm_100a kernels.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a compute entry
point is called, the call raises.  Loading the library itself needs no GPU (the CPU test-suite checks that every
symbol of the header is exported).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libknode_cosserat_b200.so")

KC_F32, KC_F64 = 0, 1
KC_MARCH_EULER, KC_MARCH_RK4 = 0, 1


class kc_rod_params(C.Structure):
    _fields_ = [("N", C.c_int32), ("reserved", C.c_int32),
                ("ds", C.c_double), ("c0", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("rhoA", C.c_double),
                ("Kse_c0Bse_inv", C.c_double * 9), ("Kbt_c0Bbt_inv", C.c_double * 9), ("Bse", C.c_double * 9),
                ("Bbt", C.c_double * 9), ("rhoJ", C.c_double * 9),
                ("Kse_vstar", C.c_double * 3), ("rhoAg", C.c_double * 3), ("C", C.c_double * 3),
                ("F_tip", C.c_double * 3), ("M_tip", C.c_double * 3),
                ("p0", C.c_double * 3), ("h0", C.c_double * 4), ("q0", C.c_double * 3), ("w0", C.c_double * 3),
                ("tendon_dirs", C.c_double * 12)]


class kc_mlp(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("hidden", C.c_int32), ("out_dim", C.c_int32), ("reserved", C.c_int32),
                ("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p)]


_SIGNATURES = {
    "kc_version": (C.c_int, []),
    "kc_last_error": (C.c_char_p, []),
    "kc_ode_fwd": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64] + [C.c_void_p] * 7),
    "kc_ode_bwd_workspace_bytes": (C.c_int64, [C.c_int, C.POINTER(kc_mlp), C.c_int64]),
    "kc_ode_bwd": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64] + [C.c_void_p] * 15
                   + [C.c_int64, C.c_void_p]),
    "kc_mlp_fwd": (C.c_int, [C.c_int, C.POINTER(kc_mlp), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kc_mlp_bwd": (C.c_int, [C.c_int, C.POINTER(kc_mlp), C.c_int64] + [C.c_void_p] * 8 + [C.c_int64, C.c_void_p]),
    "kc_march": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int, C.c_int64]
                 + [C.c_void_p] * 8),
    "kc_segment_fwd": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int32,
                                 C.POINTER(C.c_int32)] + [C.c_void_p] * 6),
    "kc_rollout_workspace_bytes": (C.c_int64, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64,
                                               C.c_int64]),
    "kc_rollout_fwd": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "kc_rollout_fwd_rk4": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "kc_rollout_fwd_range": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "kc_rollout_resumable": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64,
                                       C.c_int32]),
    "kc_rollout_host_device_bytes": (C.c_int64, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64,
                                                 C.c_int64, C.c_int32]),
    "kc_rollout_host": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "kc_rollout_bwd_workspace_bytes": (C.c_int64, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64,
                                                   C.c_int64]),
    "kc_rollout_bwd": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64]
                       + [C.c_void_p] * 9 + [C.c_int64, C.c_void_p]),
    "kc_train_step_workspace_bytes": (C.c_int64, [C.c_int, C.POINTER(kc_mlp), C.c_int64, C.c_int64, C.c_int32]),
    "kc_train_step": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.POINTER(kc_mlp), C.c_int64, C.c_int64, C.c_int32,
                                C.POINTER(C.c_int32)] + [C.c_void_p] * 9 + [C.c_int64, C.c_void_p]),
    "kc_adam_clamp": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_void_p]),
    "kc_adam_clamp_multi": (C.c_int, [C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "kc_estimate_state": (C.c_int, [C.c_int, C.POINTER(kc_rod_params), C.c_double, C.c_double, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kc_rollout_loss": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                  C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kc_plateau_step": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kc_peer_region_bytes": (C.c_int64, [C.c_int, C.c_int64]),
    "kc_peer_publish": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "kc_peer_gather_adam": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                      C.c_void_p]),
    "kc_eval_metrics": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kc_umma_selftest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "kc_fma_peak": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """The loaded shared library (raises if it has not been built: run `python -c 'import __graft_entry__ as g;
    g.build()'` or `make -C knode-cosserat_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA library first (__graft_entry__.build()); "
                               "there is no CPU fallback for this path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class kc_adam_tensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("clamp_min_zero", C.c_int32), ("reserved", C.c_int32)]


def exported_symbols():
    return list(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().kc_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def _f64(x):
    """Python float / numpy / torch tensor -> flat float64 numpy array (host)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().double().numpy()
    return np.asarray(x, dtype=np.float64).reshape(-1)


_ATTRS = [("Kse_c0Bse_inv", "Kse_plus_c0_Bse_inv", 9), ("Kbt_c0Bbt_inv", "Kbt_plus_c0_Bbt_inv", 9),
          ("Bse", "Bse", 9), ("Bbt", "Bbt", 9), ("rhoJ", "rhoJ", 9), ("Kse_vstar", "Kse_vstar", 3),
          ("rhoAg", "rhoAg", 3), ("C", "C", 3), ("F_tip", "F_tip", 3), ("M_tip", "M_tip", 3),
          ("p0", "p0", 3), ("h0", "h0", 4), ("q0", "q0", 3), ("w0", "w0", 3), ("tendon_dirs", "tendon_dirs", 12)]
_SCALARS = ["N", "ds", "c0", "c1", "c2", "rhoA"]


def _fingerprint(robot):
    """Cheap identity of every attribute the struct depends on: object id + in-place version counter for tensors,
    id + bytes for numpy arrays, value for python scalars.  No device synchronisation."""
    fp = []
    for name in _SCALARS:
        fp.append(float(getattr(robot, name)))
    for _, attr, _n in _ATTRS:
        v = getattr(robot, attr)
        if hasattr(v, "_version"):
            fp.append((id(v), v._version))
        elif isinstance(v, np.ndarray):
            fp.append((id(v), v.tobytes()))
        else:
            fp.append(repr(v))
    return tuple(fp)


def rod_params(robot) -> kc_rod_params:
    """Snapshot the derived constants of a CosseratRodTorch / CosseratRod-shaped object
    (cosserat_ode_torch.py:108-129).  Callers mutate attributes and then call compute_intermediate_terms()
    (knode.py:11-53), so the struct is rebuilt whenever any attribute object, its in-place version or a scalar changed;
    otherwise the cached struct is reused (reading device tensors back costs a synchronisation per attribute)."""
    fp = _fingerprint(robot)
    cache = robot.__dict__.get("_kc_params_cache") if hasattr(robot, "__dict__") else None
    if cache is not None and cache[0] == fp:
        return cache[1]
    p = kc_rod_params()
    p.N = int(robot.N)
    p.reserved = 0
    p.ds, p.c0, p.c1, p.c2, p.rhoA = (float(robot.ds), float(robot.c0), float(robot.c1), float(robot.c2),
                                      float(robot.rhoA))
    for name, attr, n in _ATTRS:
        v = _f64(getattr(robot, attr))
        if v.size != n:
            raise ValueError(f"robot.{attr} has {v.size} elements, expected {n}")
        getattr(p, name)[:] = v.tolist()
    if hasattr(robot, "__dict__"):
        # keep the fingerprinted objects alive: a freed tensor's id() could otherwise be reused by its replacement
        keep = [getattr(robot, attr) for _, attr, _n in _ATTRS]
        robot.__dict__["_kc_params_cache"] = (fp, p, keep)
    return p
