"""Tensor-level wrappers over the C ABI (include/knode_cosserat.h).

Every function takes CUDA torch tensors, allocates the outputs / workspace with torch (device memory plumbing only),
and launches the hand-written kernels on torch's current CUDA stream.  There is deliberately no CPU path: a CPU
tensor, a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import _kc

_DT = {torch.float32: _kc.KC_F32, torch.float64: _kc.KC_F64}


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise TypeError(f"knode-cosserat_b200 kernels compute in float32 or float64, got {t.dtype}")
    return _DT[t.dtype]


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("knode-cosserat_b200 has no CPU fallback: tensors must live on a CUDA device "
                               "(construct the robot with device='cuda')")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_SCRATCH = {}


def _scratch(device, nbytes, tag):
    """Grow-only workspace per (device, stream, call site): the backward entry points need hundreds of MB at the training
    sizes, and a fresh torch.empty per call can fall through the caching allocator to cudaMalloc (tens of ms).  The buffer
    is only touched by kernels on the stream it is keyed by, so back-to-back calls serialise on it naturally."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        _SCRATCH[key] = buf = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
    return buf


def _c(t, dtype=None):
    """contiguous (and optionally cast) view/copy"""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class Mlp:
    """Borrowed view of the four nn.Linear tensors in the layout kc_mlp wants."""

    def __init__(self, W1, b1, W2, b2):
        self.W1, self.b1, self.W2, self.b2 = (_c(W1.detach()), _c(b1.detach()), _c(W2.detach()), _c(b2.detach()))
        self.hidden, self.in_dim = self.W1.shape
        if self.W2.shape != (25, self.hidden):
            raise ValueError(f"second Linear must be [25,{self.hidden}], got {tuple(self.W2.shape)}")
        self.c = _kc.kc_mlp(self.in_dim, self.hidden, 25, 0, self.W1.data_ptr(), self.b1.data_ptr(),
                            self.W2.data_ptr(), self.b2.data_ptr())

    def cast(self, dtype):
        if self.W1.dtype == dtype:
            return self
        return Mlp(self.W1.to(dtype), self.b1.to(dtype), self.W2.to(dtype), self.b2.to(dtype))

    def ref(self):
        return C.byref(self.c)


def _mlp_ref(mlp, dtype):
    if mlp is None:
        return None, None
    m = mlp.cast(dtype)
    return m, m.ref()


def ode_fwd(P: _kc.kc_rod_params, mlp, y, yh, zh, tf):
    """kc_ode_fwd: y[Q,19], yh[Q,19], zh[Q,6], tf[Q,3] -> ys[Q,19], z[Q,6]."""
    _require_cuda(y, yh, zh, tf)
    dt = y.dtype
    y, yh, zh, tf = _c(y), _c(yh, dt), _c(zh, dt), _c(tf, dt)
    Q = y.shape[0]
    ys = torch.empty((Q, 19), dtype=dt, device=y.device)
    z = torch.empty((Q, 6), dtype=dt, device=y.device)
    keep, mref = _mlp_ref(mlp, dt)
    with torch.cuda.device(y.device):
        rc = _kc.lib().kc_ode_fwd(_dtype_code(y), C.byref(P), mref, Q, _ptr(y), _ptr(yh), _ptr(zh), _ptr(tf), _ptr(ys),
                                  _ptr(z), _stream(y.device))
    _kc.check(rc, "kc_ode_fwd")
    return ys, z


def ode_bwd(P, mlp, y, yh, zh, tf, g_ys, g_z, need_inputs=True, need_params=True):
    """kc_ode_bwd -> (g_y, g_yh, g_zh, g_tf, gW1, gb1, gW2, gb2) (None where not requested)."""
    _require_cuda(y, yh, zh, tf, g_ys, g_z)
    dt = y.dtype
    dev = y.device
    y, yh, zh, tf, g_ys, g_z = _c(y), _c(yh, dt), _c(zh, dt), _c(tf, dt), _c(g_ys, dt), _c(g_z, dt)
    Q = y.shape[0]
    gi = [torch.empty_like(t) for t in (y, yh, zh, tf)] if need_inputs else [None] * 4
    keep, mref = _mlp_ref(mlp, dt)
    gp = [None] * 4
    if mlp is not None and need_params:
        gp = [torch.empty_like(t) for t in (keep.W1, keep.b1, keep.W2, keep.b2)]
    with torch.cuda.device(dev):
        nbytes = _kc.lib().kc_ode_bwd_workspace_bytes(_dtype_code(y), mref, Q)
        ws = _scratch(dev, nbytes, "ode_bwd")
        rc = _kc.lib().kc_ode_bwd(_dtype_code(y), C.byref(P), mref, Q, _ptr(y), _ptr(yh), _ptr(zh), _ptr(tf),
                                  _ptr(g_ys), _ptr(g_z), *[_ptr(t) for t in gi], *[_ptr(t) for t in gp], _ptr(ws),
                                  int(nbytes), _stream(dev))
    _kc.check(rc, "kc_ode_bwd")
    return (*gi, *gp)


def mlp_fwd(mlp, x):
    """kc_mlp_fwd: x[Q,in] -> [Q,25]."""
    _require_cuda(x)
    x = _c(x)
    m = mlp.cast(x.dtype)
    out = torch.empty((x.shape[0], 25), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _kc.lib().kc_mlp_fwd(_dtype_code(x), m.ref(), x.shape[0], _ptr(x), _ptr(out), _stream(x.device))
    _kc.check(rc, "kc_mlp_fwd")
    return out


def mlp_bwd(mlp, x, g_out, need_input=True, need_params=True):
    """kc_mlp_bwd -> (g_x, gW1, gb1, gW2, gb2) (None where not requested)."""
    _require_cuda(x, g_out)
    x = _c(x)
    g_out = _c(g_out, x.dtype)
    m = mlp.cast(x.dtype)
    Q = x.shape[0]
    gx = torch.empty_like(x) if need_input else None
    gp = [torch.empty_like(t) for t in (m.W1, m.b1, m.W2, m.b2)] if need_params else [None] * 4
    with torch.cuda.device(x.device):
        nbytes = int(_kc.lib().kc_ode_bwd_workspace_bytes(_dtype_code(x), m.ref(), Q))
        ws = _scratch(x.device, nbytes, "mlp_bwd")
        rc = _kc.lib().kc_mlp_bwd(_dtype_code(x), m.ref(), Q, _ptr(x), _ptr(g_out), _ptr(gx), *[_ptr(t) for t in gp],
                                  _ptr(ws), nbytes, _stream(x.device))
    _kc.check(rc, "kc_mlp_bwd")
    return (gx, *gp)


def march(P, mlp, G, y, z, yh, zh, tensions, method=_kc.KC_MARCH_EULER):
    """kc_march, IN PLACE on y[B,19,N], z[B,6,N]; returns res[B,6]."""
    _require_cuda(G, y, z, yh, zh, tensions)
    dt = y.dtype
    if not (y.is_contiguous() and z.is_contiguous()):
        raise ValueError("y and z must be contiguous (they are updated in place)")
    G, yh, zh, tensions = _c(G, dt), _c(yh, dt), _c(zh, dt), _c(tensions, dt)
    B = y.shape[0]
    res = torch.empty((B, 6), dtype=dt, device=y.device)
    keep, mref = _mlp_ref(mlp, dt)
    with torch.cuda.device(y.device):
        rc = _kc.lib().kc_march(_dtype_code(y), C.byref(P), mref, method, B, _ptr(G), _ptr(y), _ptr(z), _ptr(yh),
                                _ptr(zh), _ptr(tensions), _ptr(res), _stream(y.device))
    _kc.check(rc, "kc_march")
    return res


def segment_fwd(P, mlp, Gs, key_idx, yh, zh, tensions):
    """kc_segment_fwd: Gs[S,25,N] -> [S,25,K] (key_idx given) or [S,25,N] (key_idx None)."""
    _require_cuda(Gs, yh, zh, tensions)
    dt = Gs.dtype
    Gs, yh, zh, tensions = _c(Gs), _c(yh, dt), _c(zh, dt), _c(tensions, dt)
    S, _, N = Gs.shape
    if key_idx is None:
        K, keys, cols = 0, None, N
    else:
        ki = np.ascontiguousarray(np.asarray(key_idx).reshape(-1), dtype=np.int32)
        K, cols = int(ki.size), int(ki.size)
        keys = ki.ctypes.data_as(C.POINTER(C.c_int32))
    out = torch.empty((S, 25, cols), dtype=dt, device=Gs.device)
    keep, mref = _mlp_ref(mlp, dt)
    with torch.cuda.device(Gs.device):
        rc = _kc.lib().kc_segment_fwd(_dtype_code(Gs), C.byref(P), mref, S, K, keys, _ptr(Gs), _ptr(yh), _ptr(zh),
                                      _ptr(tensions), _ptr(out), _stream(Gs.device))
    _kc.check(rc, "kc_segment_fwd")
    return out


class RolloutPlan:
    """Pre-allocated outputs + workspace for repeated rollouts of one shape (what bench.py times)."""

    def __init__(self, P, mlp, B, T, dtype, device, rows=25, want_G=False, want_iters=True, method="euler"):
        self.P, self.B, self.T, self.rows, self.dtype, self.device = P, B, T, rows, dtype, device
        if method not in ("euler", "rk4"):
            raise ValueError(f"method must be 'euler' or 'rk4', got {method!r}")
        self.method = method          # spatial march inside the shooting solve (getResidualEuler / getResidualRK4)
        self.mlp, self.mref = _mlp_ref(mlp, dtype)
        self.code = _DT[dtype]
        with torch.cuda.device(device):
            nbytes = _kc.lib().kc_rollout_workspace_bytes(self.code, C.byref(P), self.mref, B, T)
        if nbytes < 0:
            raise ValueError("kc_rollout_workspace_bytes: bad arguments")
        self.nbytes = int(nbytes)
        self.ws = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
        self.traj = torch.empty((B, T, rows, int(P.N)), dtype=dtype, device=device) if rows else None
        self.G = torch.empty((B, T, 6), dtype=dtype, device=device) if want_G else None
        self.iters = torch.empty((B, T), dtype=torch.int32, device=device) if want_iters else None

    def rebind(self, P, mlp):
        """Point a cached plan at the current rod constants / weights (same shapes)."""
        self.P = P
        self.mlp, self.mref = _mlp_ref(mlp, self.dtype)

    def run(self, tensions, y0=None, z0=None, tol=0.0, max_iter=0):
        _require_cuda(tensions, y0, z0)
        tensions = _c(tensions, self.dtype)
        if tuple(tensions.shape) != (self.B, self.T, 4):
            raise ValueError(f"tensions must be [{self.B},{self.T},4], got {tuple(tensions.shape)}")
        if y0 is not None:
            y0, z0 = _c(y0, self.dtype), _c(z0, self.dtype)
        fn = _kc.lib().kc_rollout_fwd_rk4 if self.method == "rk4" else _kc.lib().kc_rollout_fwd
        with torch.cuda.device(self.device):
            rc = fn(self.code, C.byref(self.P), self.mref, self.B, self.T, _ptr(tensions), _ptr(y0), _ptr(z0), float(tol),
                    int(max_iter), self.rows, _ptr(self.traj), _ptr(self.G), _ptr(self.iters), _ptr(self.ws), self.nbytes,
                    _stream(self.device))
        _kc.check(rc, "kc_rollout_fwd_rk4" if self.method == "rk4" else "kc_rollout_fwd")
        return self.traj


class HostRolloutPlan:
    """kc_rollout_host: host tensions in -> host trajectory out, transfers pipelined with the solve inside the call.

    Holds the device scratch and PINNED host buffers for one shape; `run` returns torch CPU tensors that alias the pinned
    buffers (they are overwritten by the next run) unless `out` is given.
    """

    def __init__(self, P, mlp, B, T, dtype, device, rows=25, want_G=False, want_iters=True, segments=0):
        if rows not in (25, 50):
            raise ValueError("rows must be 25 or 50")
        self.P, self.B, self.T, self.rows, self.dtype, self.device = P, B, T, rows, dtype, device
        self.mlp, self.mref = _mlp_ref(mlp, dtype)
        self.code = _DT[dtype]
        self.segments = int(segments)
        with torch.cuda.device(device):
            nbytes = _kc.lib().kc_rollout_host_device_bytes(self.code, C.byref(P), self.mref, B, T, rows)
        if nbytes < 0:
            raise ValueError("kc_rollout_host_device_bytes: bad arguments")
        self.nbytes = int(nbytes)
        self.dbuf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
        self.tens_h = torch.empty((B, T, 4), dtype=dtype).pin_memory()
        self.traj_h = None
        self.G_h = torch.empty((B, T, 6), dtype=dtype).pin_memory() if want_G else None
        self.iters_h = torch.empty((B, T), dtype=torch.int32).pin_memory() if want_iters else None

    def rebind(self, P, mlp):
        self.P = P
        self.mlp, self.mref = _mlp_ref(mlp, self.dtype)

    def run(self, tensions_host, y0=None, z0=None, tol=0.0, max_iter=0, out=None):
        """tensions_host: numpy array or CPU tensor [B,T,4]; out: optional CPU tensor [B,T,rows,N] (ideally pinned)."""
        th = torch.as_tensor(tensions_host)
        if th.is_cuda:
            raise ValueError("kc_rollout_host takes HOST tensions (use RolloutPlan for device tensors)")
        if tuple(th.shape) != (self.B, self.T, 4):
            raise ValueError(f"tensions must be [{self.B},{self.T},4], got {tuple(th.shape)}")
        if th.dtype == self.dtype and th.is_contiguous() and th.is_pinned():
            src = th
        else:
            self.tens_h.copy_(th)          # dtype conversion + staging into pinned memory
            src = self.tens_h
        if out is None:
            if self.traj_h is None:
                self.traj_h = torch.empty((self.B, self.T, self.rows, int(self.P.N)), dtype=self.dtype).pin_memory()
            out = self.traj_h
        if out.is_cuda or out.dtype != self.dtype or not out.is_contiguous() or \
                tuple(out.shape) != (self.B, self.T, self.rows, int(self.P.N)):
            raise ValueError("out must be a contiguous CPU tensor [B,T,rows,N] of the plan's dtype")
        if y0 is not None:
            _require_cuda(y0, z0)
            y0, z0 = _c(y0, self.dtype), _c(z0, self.dtype)
        with torch.cuda.device(self.device):
            rc = _kc.lib().kc_rollout_host(self.code, C.byref(self.P), self.mref, self.B, self.T, src.data_ptr(),
                                           _ptr(y0), _ptr(z0), float(tol), int(max_iter), self.rows, out.data_ptr(),
                                           self.G_h.data_ptr() if self.G_h is not None else None,
                                           self.iters_h.data_ptr() if self.iters_h is not None else None,
                                           _ptr(self.dbuf), self.nbytes, self.segments, _stream(self.device))
        _kc.check(rc, "kc_rollout_host")
        return out


def rollout(P, mlp, tensions, y0=None, z0=None, tol=0.0, max_iter=0, rows=25, want_G=False, method="euler"):
    """kc_rollout_fwd (method="rk4": kc_rollout_fwd_rk4): tensions[B,T,4] -> traj[B,T,rows,N] (+ G[B,T,6]) and iters[B,T]."""
    _require_cuda(tensions)
    B, T, _ = tensions.shape
    plan = RolloutPlan(P, mlp, B, T, tensions.dtype, tensions.device, rows, want_G, method=method)
    plan.run(tensions, y0, z0, tol, max_iter)
    return plan.traj, plan.G, plan.iters


def rollout_bwd(P, mlp, tensions, traj, g_traj, want_g_tensions=True, want_params=True):
    """kc_rollout_bwd: reverse mode through the rollout.  tensions[B,T,4], traj[B,T,25,N] (forward result), g_traj
    (dL/dtraj) -> (g_tensions|None, gW1, gb1, gW2, gb2 | Nones)."""
    _require_cuda(tensions, traj, g_traj)
    dt = traj.dtype
    dev = traj.device
    tensions, traj, g_traj = _c(tensions, dt), _c(traj), _c(g_traj, dt)
    B, T, rows, N = traj.shape
    if rows != 25:
        raise ValueError("rollout_bwd needs the 25-row trajectory [y;z]")
    keep, mref = _mlp_ref(mlp, dt)
    gt = torch.empty_like(tensions) if want_g_tensions else None
    gp = [None] * 4
    if mlp is not None and want_params:
        gp = [torch.empty_like(t) for t in (keep.W1, keep.b1, keep.W2, keep.b2)]
    with torch.cuda.device(dev):
        nbytes = int(_kc.lib().kc_rollout_bwd_workspace_bytes(_dtype_code(traj), C.byref(P), mref, B, T))
        ws = _scratch(dev, nbytes, "rollout_bwd")
        rc = _kc.lib().kc_rollout_bwd(_dtype_code(traj), C.byref(P), mref, B, T, _ptr(tensions), _ptr(traj), _ptr(g_traj),
                                      _ptr(gt), *[_ptr(g) for g in gp], _ptr(ws), nbytes, _stream(dev))
    _kc.check(rc, "kc_rollout_bwd")
    return (gt, *gp)


def rollout_loss(traj, target, key_idx, scale=1.0, g_traj=None, loss=None):
    """kc_rollout_loss: the 4-term training loss (physics_train.py:345-352) of a rollout against a target trajectory,
    fused with its cotangent.  traj, target [B,T,25,N] -> (loss float64[1] tensor, g_traj [B,T,25,N])."""
    _require_cuda(traj, target)
    traj, target = _c(traj), _c(target, traj.dtype)
    B, T, rows, N = traj.shape
    if rows != 25 or tuple(target.shape) != tuple(traj.shape):
        raise ValueError("rollout_loss needs traj and target of shape [B,T,25,N]")
    ki = np.ascontiguousarray(np.asarray(key_idx).reshape(-1), dtype=np.int32)
    if g_traj is None:
        g_traj = torch.empty_like(traj)
    if loss is None:
        loss = torch.zeros(1, dtype=torch.float64, device=traj.device)
    with torch.cuda.device(traj.device):
        rc = _kc.lib().kc_rollout_loss(_dtype_code(traj), B, T, N, int(ki.size), ki.ctypes.data_as(C.POINTER(C.c_int32)),
                                       _ptr(traj), _ptr(target), float(scale), _ptr(loss), _ptr(g_traj), _stream(traj.device))
    _kc.check(rc, "kc_rollout_loss")
    return loss, g_traj


class BpttStepPlan:
    """One rollout-training step with nothing but kernel launches: kc_rollout_fwd -> kc_rollout_loss -> kc_rollout_bwd,
    every buffer allocated once.  As in TrainStepPlan the four gradients are views of ONE flat buffer whose last element
    receives the loss, so a single all-reduce covers gradients and loss."""

    def __init__(self, P, weights, tensions, target, key_idx, scale=1.0):
        _require_cuda(tensions, target, *weights)
        self.P = P
        self.dt, self.dev = target.dtype, target.device
        self.tensions, self.target = _c(tensions, self.dt), _c(target)
        self.B, self.T, rows, self.N = self.target.shape
        if rows != 25:
            raise ValueError("target must be [B,T,25,N]")
        self.key, self.scale = np.asarray(key_idx).reshape(-1), float(scale)
        self.mlp = Mlp(*weights)
        sizes = [t.numel() for t in (self.mlp.W1, self.mlp.b1, self.mlp.W2, self.mlp.b2)]
        self.flat = torch.zeros(sum(sizes) + 1, dtype=self.dt, device=self.dev)
        self.grads, off = [], 0
        for t, n in zip((self.mlp.W1, self.mlp.b1, self.mlp.W2, self.mlp.b2), sizes):
            self.grads.append(self.flat[off:off + n].view_as(t))
            off += n
        self.loss64 = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.fwd = RolloutPlan(P, self.mlp, self.B, self.T, self.dt, self.dev, rows=25)
        self.g_traj = torch.empty_like(self.target)
        with torch.cuda.device(self.dev):
            self.nbytes = int(_kc.lib().kc_rollout_bwd_workspace_bytes(_DT[self.dt], C.byref(P), self.mlp.ref(), self.B, self.T))
        self.ws = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=self.dev)

    def run(self):
        self.fwd.P = self.P
        traj = self.fwd.run(self.tensions)
        rollout_loss(traj, self.target, self.key, self.scale, g_traj=self.g_traj, loss=self.loss64)
        with torch.cuda.device(self.dev):
            rc = _kc.lib().kc_rollout_bwd(_DT[self.dt], C.byref(self.P), self.mlp.ref(), self.B, self.T, _ptr(self.tensions),
                                          _ptr(traj), _ptr(self.g_traj), None, *[_ptr(g) for g in self.grads], _ptr(self.ws),
                                          self.nbytes, _stream(self.dev))
        _kc.check(rc, "kc_rollout_bwd")
        self.flat[-1:].copy_(self.loss64)
        return self.flat


def train_step(P, mlp, traj, controls, key_idx, want_pred=False):
    """kc_train_step: traj[B,T,25,N], controls[B,T,4] -> (loss float64[1] tensor, (gW1,gb1,gW2,gb2), pred|None)."""
    _require_cuda(traj, controls)
    dt = traj.dtype
    dev = traj.device
    traj, controls = _c(traj), _c(controls, dt)
    B, T, _, N = traj.shape
    ki = np.ascontiguousarray(np.asarray(key_idx).reshape(-1), dtype=np.int32)
    K = int(ki.size)
    keep, mref = _mlp_ref(mlp, dt)
    grads = [torch.empty_like(t) for t in (keep.W1, keep.b1, keep.W2, keep.b2)]
    loss = torch.empty(1, dtype=torch.float64, device=dev)
    pred = torch.empty((B, T - 1, 25, K), dtype=dt, device=dev) if want_pred else None
    with torch.cuda.device(dev):
        nbytes = int(_kc.lib().kc_train_step_workspace_bytes(_dtype_code(traj), mref, B, T, K))
        ws = _scratch(dev, nbytes, "train_step")
        rc = _kc.lib().kc_train_step(_dtype_code(traj), C.byref(P), mref, B, T, K,
                                     ki.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(traj), _ptr(controls),
                                     _ptr(loss), *[_ptr(g) for g in grads], _ptr(pred), _ptr(ws), nbytes, _stream(dev))
    _kc.check(rc, "kc_train_step")
    return loss, tuple(grads), pred


def adam_clamp(param, grad, exp_avg, exp_avg_sq, step, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
               clamp=True):
    """kc_adam_clamp, in place on param / exp_avg / exp_avg_sq (flat or any contiguous shape)."""
    _require_cuda(param, grad, exp_avg, exp_avg_sq)
    for t in (param, exp_avg, exp_avg_sq):
        if not t.is_contiguous():
            raise ValueError("adam_clamp updates in place: tensors must be contiguous")
    grad = _c(grad, param.dtype)
    with torch.cuda.device(param.device):
        rc = _kc.lib().kc_adam_clamp(_dtype_code(param), param.numel(), _ptr(param), _ptr(grad), _ptr(exp_avg),
                                     _ptr(exp_avg_sq), int(step), float(lr), float(betas[0]), float(betas[1]),
                                     float(eps), float(weight_decay), 1 if clamp else 0, _stream(param.device))
    _kc.check(rc, "kc_adam_clamp")


class TrainStepPlan:
    """kc_train_step with every output and the workspace allocated once: `run()` launches kernels only (no allocation, no
    host synchronisation), so it can be captured in a CUDA graph.  The four gradients are views of ONE flat buffer whose
    last element receives the loss (cast to the parameter dtype): a single all-reduce covers gradients and loss."""

    def __init__(self, P, weights, traj, controls, key_idx):
        _require_cuda(traj, controls, *weights)
        self.P = P
        self.dt, self.dev = traj.dtype, traj.device
        self.traj, self.controls = _c(traj), _c(controls, traj.dtype)
        self.B, self.T, _, self.N = self.traj.shape
        self.ki = np.ascontiguousarray(np.asarray(key_idx).reshape(-1), dtype=np.int32)
        self.K = int(self.ki.size)
        self.mlp = Mlp(*weights)            # borrows the parameter storage: in-place optimiser updates are seen
        for w, t in zip(weights, (self.mlp.W1, self.mlp.b1, self.mlp.W2, self.mlp.b2)):
            if w.data_ptr() != t.data_ptr():
                raise ValueError("parameters must be contiguous tensors of the trajectory dtype")
        sizes = [t.numel() for t in (self.mlp.W1, self.mlp.b1, self.mlp.W2, self.mlp.b2)]
        self.flat = torch.zeros(sum(sizes) + 1, dtype=self.dt, device=self.dev)
        self.grads, off = [], 0
        for t, n in zip((self.mlp.W1, self.mlp.b1, self.mlp.W2, self.mlp.b2), sizes):
            self.grads.append(self.flat[off:off + n].view_as(t))
            off += n
        self.loss64 = torch.zeros(1, dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            self.nbytes = int(_kc.lib().kc_train_step_workspace_bytes(_DT[self.dt], self.mlp.ref(), self.B, self.T, self.K))
        self.ws = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=self.dev)

    def run(self):
        with torch.cuda.device(self.dev):
            rc = _kc.lib().kc_train_step(_DT[self.dt], C.byref(self.P), self.mlp.ref(), self.B, self.T, self.K,
                                         self.ki.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(self.traj),
                                         _ptr(self.controls), _ptr(self.loss64), *[_ptr(g) for g in self.grads], None,
                                         _ptr(self.ws), self.nbytes, _stream(self.dev))
        _kc.check(rc, "kc_train_step")
        self.flat[-1:].copy_(self.loss64)    # f64 -> parameter dtype, on the stream
        return self.flat


class AdamClampMulti:
    """kc_adam_clamp_multi: Adam + non-negative clamp of all parameter tensors in one launch; step count and learning
    rate live on the device (graph-capturable)."""

    def __init__(self, params, grads, clamp_flags, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 exp_avg=None, exp_avg_sq=None, step=0):
        _require_cuda(*params, *grads)
        self.dev, self.dt = params[0].device, params[0].dtype
        self.params, self.grads = list(params), list(grads)
        self.exp_avg = exp_avg if exp_avg is not None else [torch.zeros_like(p) for p in params]
        self.exp_avg_sq = exp_avg_sq if exp_avg_sq is not None else [torch.zeros_like(p) for p in params]
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.step_dev = torch.full((1,), int(step), dtype=torch.int64, device=self.dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float64, device=self.dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
        n = len(params)
        self.arr = (_kc.kc_adam_tensor * n)()
        for i, (p, g, m, v, c) in enumerate(zip(self.params, self.grads, self.exp_avg, self.exp_avg_sq, clamp_flags)):
            for t in (p, g, m, v):
                if not t.is_contiguous() or t.dtype != self.dt:
                    raise ValueError("adam tensors must be contiguous and share one dtype")
            self.arr[i] = _kc.kc_adam_tensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                             1 if c else 0, 0)

    def set_lr(self, lr):
        self.lr_dev.fill_(float(lr))

    # -- gradient all-reduce fused with the update, over NVLink peer memory (kc_peer_publish / kc_peer_gather_adam) ------
    def enable_peer_allreduce(self, flat, regions=None, rank=None, world=None):
        """`flat` = the [gradients | loss] buffer whose leading views are this optimiser's gradients.  regions: list of
        device pointers to every rank's symmetric region (tests); default: one torch.distributed._symmetric_memory
        allocation, rendezvoused over the world group.  Returns False (and leaves the NCCL path in place) if symmetric memory
        cannot be set up on this system."""
        n = flat.numel()
        nbytes = int(_kc.lib().kc_peer_region_bytes(_DT[self.dt], n))
        if regions is None:
            import torch.distributed as dist
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._sym = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.dev)
                self._sym.zero_()
                self._sym_hdl = symm_mem.rendezvous(self._sym, dist.group.WORLD)
                regions = [int(p) for p in self._sym_hdl.buffer_ptrs]
                rank, world = dist.get_rank(), dist.get_world_size()
                torch.cuda.synchronize(self.dev)
                dist.barrier()            # every region is zeroed before anybody publishes
            except Exception as e:        # no peer access / no symmetric-memory support: keep the collective-library path
                self._peer_error = repr(e)
                return False
        if world > 8:
            return False
        self._peer_flat, self._peer_n = flat, n
        self._peer_rank, self._peer_world = int(rank), int(world)
        self._peer_regions = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in regions])
        self._peer_tickets = torch.zeros(2, dtype=torch.int32, device=self.dev)
        return True

    def run_peer(self):
        """publish -> gather + Adam + clamp (two launches; replaces all_reduce(flat) + run())."""
        L = _kc.lib()
        with torch.cuda.device(self.dev):
            st = _stream(self.dev)
            rc = L.kc_peer_publish(_DT[self.dt], self._peer_world, self._peer_rank, C.cast(self._peer_regions, C.c_void_p),
                                   self._peer_n, _ptr(self._peer_flat), _ptr(self.step_dev), _ptr(self._peer_tickets), st)
            _kc.check(rc, "kc_peer_publish")
            rc = L.kc_peer_gather_adam(_DT[self.dt], self._peer_world, self._peer_rank, C.cast(self._peer_regions, C.c_void_p),
                                       self._peer_n, _ptr(self._peer_flat), len(self.params), C.cast(self.arr, C.c_void_p),
                                       _ptr(self.step_dev), _ptr(self.lr_dev), float(self.betas[0]), float(self.betas[1]),
                                       float(self.eps), float(self.weight_decay), self._peer_tickets[1:].data_ptr(), st)
            _kc.check(rc, "kc_peer_gather_adam")

    def run(self):
        with torch.cuda.device(self.dev):
            rc = _kc.lib().kc_adam_clamp_multi(_DT[self.dt], len(self.params), C.cast(self.arr, C.c_void_p),
                                               _ptr(self.step_dev), _ptr(self.lr_dev), float(self.betas[0]),
                                               float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                               _ptr(self.ticket), _stream(self.dev))
        _kc.check(rc, "kc_adam_clamp_multi")


def eval_metrics(pred, ref, node=None, want_mse=True):
    """kc_eval_metrics: pred [E,Ta,rows,N], ref [E,Tb,rows,N] (or one pair without the leading axis) on the device ->
    (dtw float64[E], mse float64[E] | None): exact L1 DTW of the tip positions and the pos + Euler-'zyx' MSE x 1000."""
    _require_cuda(pred, ref)
    single = pred.ndim == 3
    if single:
        pred, ref = pred[None], ref[None]
    pred, ref = _c(pred), _c(ref, pred.dtype)
    E, Ta, rows, N = pred.shape
    Tb = ref.shape[1]
    if ref.shape[0] != E or ref.shape[2] != rows or ref.shape[3] != N:
        raise ValueError("pred and ref must share E, rows and N")
    node = N - 1 if node is None else int(node)
    dtw = torch.empty(E, dtype=torch.float64, device=pred.device)
    mse = torch.empty(E, dtype=torch.float64, device=pred.device) if want_mse else None
    with torch.cuda.device(pred.device):
        rc = _kc.lib().kc_eval_metrics(_dtype_code(pred), E, Ta, Tb, rows, N, node, _ptr(pred), _ptr(ref), _ptr(dtw),
                                       _ptr(mse), _stream(pred.device))
    _kc.check(rc, "kc_eval_metrics")
    if single:
        return dtw[0], (mse[0] if want_mse else None)
    return dtw, mse


class PlateauDevice:
    """kc_plateau_step: ReduceLROnPlateau('min', patience, factor) with its state and the learning rate on the device."""

    def __init__(self, lr_dev, patience=80, factor=0.5, threshold=1e-4, eps=1e-8):
        self.lr_dev = lr_dev
        self.state = torch.tensor([float("inf"), 0.0, float(patience), float(factor), float(threshold), float(eps)],
                                  dtype=torch.float64, device=lr_dev.device)

    def run(self, loss_elem):
        """loss_elem: a one-element device tensor (fp32 or fp64) holding this epoch's loss."""
        with torch.cuda.device(self.lr_dev.device):
            rc = _kc.lib().kc_plateau_step(_DT[loss_elem.dtype], loss_elem.data_ptr(), _ptr(self.state), _ptr(self.lr_dev),
                                           _stream(self.lr_dev.device))
        _kc.check(rc, "kc_plateau_step")

    def get_last_lr(self):
        return [float(self.lr_dev.item())]

    @property
    def lr(self):
        return float(self.lr_dev.item())

    @lr.setter
    def lr(self, value):               # a manual change of the learning rate reaches the (captured) optimiser kernels
        self.lr_dev.fill_(float(value))

    @property
    def best(self):
        return float(self.state[0].item())

    @property
    def bad(self):
        return int(self.state[1].item())


def fma_peak(dtype, iters, device):
    """Measured FP32/FP64 FMA-pipe throughput in FLOP/s (CUDA-event timed), used as the rollout's roofline peak."""
    code = _DT[dtype]
    scratch = torch.empty(148 * 8 * 256 * 2, dtype=dtype, device=device)
    flops = C.c_double(0.0)
    with torch.cuda.device(device):
        st = _stream(device)
        for _ in range(2):
            _kc.check(_kc.lib().kc_fma_peak(code, int(iters), C.byref(flops), _ptr(scratch), st), "kc_fma_peak")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _kc.check(_kc.lib().kc_fma_peak(code, int(iters), C.byref(flops), _ptr(scratch), st), "kc_fma_peak")
        e1.record()
        e1.synchronize()
    return flops.value / (e0.elapsed_time(e1) * 1e-3)


def estimate_state(P, L, del_t, data, tensions):
    """kc_estimate_state: data[B,T,7,N] (p, h on the full grid), tensions[B,T,4] -> est[B,T,25,N]."""
    _require_cuda(data, tensions)
    dt = data.dtype
    data, tensions = _c(data), _c(tensions, dt)
    B, T, rows, N = data.shape
    if rows != 7 or N != int(P.N) or tuple(tensions.shape) != (B, T, 4):
        raise ValueError(f"estimate_state: data must be [B,T,7,{int(P.N)}] and tensions [B,T,4], got "
                         f"{tuple(data.shape)} and {tuple(tensions.shape)}")
    est = torch.empty((B, T, 25, N), dtype=dt, device=data.device)
    with torch.cuda.device(data.device):
        rc = _kc.lib().kc_estimate_state(_dtype_code(data), C.byref(P), float(L), float(del_t), B, T, _ptr(data),
                                         _ptr(tensions), _ptr(est), _stream(data.device))
    _kc.check(rc, "kc_estimate_state")
    return est
