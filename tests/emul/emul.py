"""TEST-ONLY loader of the host harness (tests/emul/kc_emul.cpp)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import _kc  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libkc_emul.so")
        src = os.path.join(HERE, "kc_emul.cpp")
        hdrs = [os.path.join(ROOT, "knode-cosserat_b200", "csrc", h) for h in
                ("kc_common.cuh", "kc_rod.cuh", "kc_rollout_core.cuh", "kc_rollout_wide.cuh", "kc_adjoint.cuh",
                 "kc_bptt_core.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
        _lib = C.CDLL(so)
        _lib.kc_emul_rollout.restype = C.c_int
    return _lib


def rollout(P, ctl, dtype=np.float64, mlp=None, tol=0.0, max_iter=60, wide=False):
    """P: object with the derived rod attributes; ctl[B,T,4] -> traj[B,T,25,N], iters[B,T], G[B,T,6]."""
    ctl = np.ascontiguousarray(ctl, dtype=dtype)
    B, T, _ = ctl.shape
    N = int(P.N)
    traj = np.zeros((B, T, 25, N), dtype)
    iters = np.zeros((B, T), np.int32)
    G = np.zeros((B, T, 6), dtype)
    p = _kc.rod_params(P)
    if mlp is None:
        in_dim = hidden = 0
        ptrs = [None] * 4
        keep = []
    else:
        keep = [np.ascontiguousarray(mlp[k], dtype=dtype) for k in ("W1", "b1", "W2", "b2")]
        hidden, in_dim = keep[0].shape
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in keep]
    rc = lib().kc_emul_rollout(C.c_int(0 if dtype == np.float32 else 1), C.byref(p), C.c_int(in_dim), C.c_int(hidden),
                               *ptrs, C.c_int64(B), C.c_int64(T), ctl.ctypes.data_as(C.c_void_p),
                               traj.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p),
                               G.ctypes.data_as(C.c_void_p), C.c_double(tol), C.c_int(max_iter), C.c_int(int(wide)))
    assert rc == 0
    # device layout per rod is [T][25*N] with k = row*N + node: already [T,25,N]
    return traj, iters, G


def bptt(P, ctl, traj, gtraj, mlp=None, dtype=np.float64):
    """Reverse mode through the rollout on the host harness: returns (g_tensions[B,T,4], xs[Q,in], gos[Q,25]) where the
    MLP samples (x, dL/do) still have to be reduced to weight gradients."""
    ctl = np.ascontiguousarray(ctl, dtype=dtype)
    traj = np.ascontiguousarray(traj, dtype=dtype)
    gtraj = np.ascontiguousarray(gtraj, dtype=dtype)
    B, T, _ = ctl.shape
    N = int(P.N)
    p = _kc.rod_params(P)
    gten = np.zeros((B, T, 4), dtype)
    if mlp is None:
        in_dim = hidden = 0
        ptrs, keep = [None] * 4, []
        xs = gos = np.zeros((1, 1), dtype)
    else:
        keep = [np.ascontiguousarray(mlp[k], dtype=dtype) for k in ("W1", "b1", "W2", "b2")]
        hidden, in_dim = keep[0].shape
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in keep]
        Q = B * (T - 1) * (N - 1) * 2
        xs, gos = np.zeros((Q, in_dim), dtype), np.zeros((Q, 25), dtype)
    L = lib()
    L.kc_emul_bptt.restype = C.c_int
    rc = L.kc_emul_bptt(C.c_int(0 if dtype == np.float32 else 1), C.byref(p), C.c_int(in_dim), C.c_int(hidden), *ptrs,
                        C.c_int64(B), C.c_int64(T), ctl.ctypes.data_as(C.c_void_p), traj.ctypes.data_as(C.c_void_p),
                        gtraj.ctypes.data_as(C.c_void_p), gten.ctypes.data_as(C.c_void_p),
                        xs.ctypes.data_as(C.c_void_p), gos.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return gten, xs, gos
