"""CPU oracle for the KNODE-Cosserat hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy (float64 unless a dtype is passed) restatement of the reference's algorithm, function by
function, each citing the reference file:line it follows (paths relative to /root/reference/).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module; nothing under `knode-cosserat_b200/` does.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here against the vectors in
`tests/golden/*.npz`, which `tests/golden/make_golden.py` produced by importing and running the unmodified
reference in the build container (the reference ships no tests or fixtures of its own — SURVEY.md §4).

Everything is batched over leading dimensions: a "rod" is y[...,19,N], z[...,6,N]; a node sample is y[...,19].
State layout (cosserat_ode_torch.py:141-152): y = [p(0:3), h(3:7) wxyz, n(7:10), m(10:13), q(13:16), w(16:19)],
z = [v(0:3), u(3:6)].
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------------------


class RodParams:
    """Physical constants + derived terms.

    Defaults: cosserat_ode.py:15-47 (== cosserat_ode_torch.py:14-45).
    Derived:  compute_intermediate_terms, cosserat_ode.py:58-78 (== cosserat_ode_torch.py:108-129).
    """

    def __init__(self):
        self.L = 0.4
        self.N = 10
        self.E = 109e9
        self.r = 0.0012
        self.rho = 8000.0
        self.vstar = np.array([0.0, 0.0, 1.0])
        self.g = np.array([0.0, 0.0, -9.81])
        self.Bse = np.zeros((3, 3))
        self.Bbt = np.diag([3e-2, 3e-2, 3e-2])
        self.C = np.array([1e-4, 1e-4, 1e-4])
        self.del_t = 0.005
        self.F_tip = np.zeros(3)
        self.M_tip = np.zeros(3)
        self.n_tendons = 4
        theta = np.pi / self.n_tendons
        self.tendon_offset = 0.02
        self.tendon_dirs = np.array([[np.cos(theta + k * np.pi / 2), np.sin(theta + k * np.pi / 2), 0.0]
                                     for k in range(4)])
        self.p0 = np.zeros(3)
        self.h0 = np.array([1.0, 0.0, 0.0, 0.0])
        self.q0 = np.zeros(3)
        self.w0 = np.zeros(3)
        self.compute_intermediate_terms()

    def compute_intermediate_terms(self):
        self.A = np.pi * self.r ** 2
        self.G = self.E / (2 * (1 + 0.3))
        self.ds = self.L / (self.N - 1)
        self.J = np.diag([np.pi * self.r ** 4 / 4, np.pi * self.r ** 4 / 4, np.pi * self.r ** 4 / 2])
        self.Kse = np.diag([self.G * self.A, self.G * self.A, self.E * self.A])
        self.Kbt = np.diag([self.E * self.J[0, 0], self.E * self.J[1, 1], self.G * self.J[2, 2]])
        self.c0 = 1.5 / self.del_t
        self.c1 = -2 / self.del_t
        self.c2 = 0.5 / self.del_t
        self.Kse_plus_c0_Bse_inv = np.linalg.inv(self.Kse + self.c0 * self.Bse)
        self.Kbt_plus_c0_Bbt_inv = np.linalg.inv(self.Kbt + self.c0 * self.Bbt)
        self.Kse_vstar = self.Kse @ self.vstar
        self.rhoA = self.rho * self.A
        self.rhoAg = self.rho * self.A * self.g
        self.rhoJ = self.rho * self.J


def setup_params(P: RodParams, mod=None) -> RodParams:
    """knode.setup_robot, knode.py:6-53."""
    P.del_t = 0.05
    P.L = 0.635
    P.tendon_offset = 0.04445
    P.r = 0.003175
    P.rho = 1411.6751
    P.E = 2.757903e9
    Bbt = 3e-2
    if mod is None:
        pass
    elif mod == "noair":
        P.C = np.zeros(3)
    elif mod == "nsw":
        P.g = np.zeros(3)
    elif mod == "short":
        P.L = 0.4
    elif mod == "damping":
        Bbt = 0.2
    elif mod == "dampstiff":
        Bbt = 0.2
        P.E = 10e9
    elif mod == "lengthstiff":
        P.L = 0.4
        P.E = 10e9
    elif mod == "youngs":
        P.E = 10e9
    else:
        raise Exception("Unknown mod " + mod)
    P.Bbt = np.diag([Bbt, Bbt, Bbt])
    P.compute_intermediate_terms()
    return P


def calc_controls(control_type, control_arg, del_t, train_len):
    """physics_controls.py:3-33 (the 'ramp' branch raises NameError in the reference; here: same Exception class
    as an unknown type)."""
    np.random.seed(int(control_arg))
    controls = []
    for i in range(1, train_len + 1):
        if control_type == "sine":
            sin_period = control_arg / del_t
            ph = 2 * np.pi / 4
            T = [6 + np.sin(2 * np.pi * i / sin_period + k * ph) for k in range(4)]
        elif control_type == "step":
            s = 0 if i * del_t < 1.5 else control_arg
            T = [5 + s, 5, 5, 5 + s]
        elif control_type == "random":
            T = [5 + 5 * np.random.rand() for _ in range(4)]
        else:
            raise Exception("Unknown control type " + control_type)
        controls.append(T)
    return controls


# ----------------------------------------------------------------------------------------------------------
# MLP (cosserat_ode_torch.py:60-62,131-134; numpy twin cosserat_ode.py:90-112)
# ----------------------------------------------------------------------------------------------------------


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def mlp_forward(mlp, x):
    """mlp = dict(W1[H,in], b1[H], W2[25,H], b2[25]); x[...,in] -> [...,25]."""
    a = elu(x @ mlp["W1"].T.astype(x.dtype) + mlp["b1"].astype(x.dtype))
    return a @ mlp["W2"].T.astype(x.dtype) + mlp["b2"].astype(x.dtype)


# ----------------------------------------------------------------------------------------------------------
# per-node ODE (cosserat_ode.py:114-186 == cosserat_ode_torch.py:137-214 == :217-322)
# ----------------------------------------------------------------------------------------------------------


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _mv(M, x):
    """(batched or constant) 3x3 @ vector."""
    return np.einsum("...ij,...j->...i", M, x)


def quat_to_R(h):
    """Eq(10): cosserat_ode.py:133-137."""
    h1, h2, h3, h4 = h[..., 0], h[..., 1], h[..., 2], h[..., 3]
    s = 2.0 / np.sum(h * h, axis=-1)
    M = np.stack([np.stack([-h3 ** 2 - h4 ** 2, h2 * h3 - h4 * h1, h2 * h4 + h3 * h1], -1),
                  np.stack([h2 * h3 + h4 * h1, -h2 ** 2 - h4 ** 2, h3 * h4 - h2 * h1], -1),
                  np.stack([h2 * h4 - h3 * h1, h3 * h4 + h2 * h1, -h2 ** 2 - h3 ** 2], -1)], -2)
    return np.eye(3, dtype=h.dtype) + s[..., None, None] * M


def ode(P: RodParams, y, yh, zh, tf, mlp=None, nn_input_history=False):
    """Returns (ys[...,19], z[...,6]).  tf is the global-frame distributed tendon force [...,3]."""
    dt = y.dtype
    c = lambda a: np.asarray(a, dtype=dt)
    h, n, m = y[..., 3:7], y[..., 7:10], y[..., 10:13]
    q, w, vh, uh = y[..., 13:16], y[..., 16:19], zh[..., 0:3], zh[..., 3:6]
    R = quat_to_R(h)
    RT = np.swapaxes(R, -1, -2)
    # Eq(6), cosserat_ode.py:140-142
    v = _mv(c(P.Kse_plus_c0_Bse_inv), _mv(RT, n) + c(P.Kse_vstar) - _mv(c(P.Bse), vh))
    u = _mv(c(P.Kbt_plus_c0_Bbt_inv), _mv(RT, m) - _mv(c(P.Bbt), uh))
    z = np.concatenate([v, u], -1)
    # Eq(5), :146-148
    c0 = dt.type(P.c0)
    yt = c0 * y + yh
    zt = c0 * z + zh
    vt, ut, qt, wt = zt[..., 0:3], zt[..., 3:6], yt[..., 13:16], yt[..., 16:19]
    # Eq(3), :151
    f = c(P.rhoAg) - _mv(R, c(P.C) * q * np.abs(q)) + tf
    # Eq(7), :154-158
    ps = _mv(R, v)
    ns = dt.type(P.rhoA) * _mv(R, _cross(w, q) + qt) - f
    rhoJ = c(P.rhoJ)
    ms = _mv(R, _cross(w, _mv(rhoJ, w)) + _mv(rhoJ, wt)) - _cross(ps, n)
    qs = vt - _cross(u, q) + _cross(w, v)
    ws = ut - _cross(u, w)
    # Eq(9), :161-165
    u1, u2, u3 = u[..., 0], u[..., 1], u[..., 2]
    h1, h2, h3, h4 = h[..., 0], h[..., 1], h[..., 2], h[..., 3]
    hs = 0.5 * np.stack([-u1 * h2 - u2 * h3 - u3 * h4,
                         u1 * h1 + u3 * h3 - u2 * h4,
                         u2 * h1 - u3 * h2 + u1 * h4,
                         u3 * h1 + u2 * h2 - u1 * h3], -1)
    ys = np.concatenate([ps, hs, ns, ms, qs, ws], -1)
    if mlp is not None:
        # :169-184 — MLP sees the PRE-correction z; z is corrected after ys is formed
        if nn_input_history:
            x = np.concatenate([y, yh, z, zh, tf], -1)
        else:
            x = np.concatenate([y, z, tf], -1)
        o = mlp_forward(mlp, x)
        ys = ys + o[..., :19]
        z = z + o[..., 19:]
    return ys, z


# ----------------------------------------------------------------------------------------------------------
# shooting residual / spatial march (cosserat_ode.py:188-213, :215-255)
# ----------------------------------------------------------------------------------------------------------


def tendon_force(P, tensions):
    """cosserat_ode.py:195 — uniform global-frame distributed load."""
    return np.asarray(tensions) @ np.asarray(P.tendon_dirs, dtype=np.asarray(tensions).dtype)


def march_euler(P, G, y, z, yh, zh, tensions, mlp=None, nn_input_history=False):
    """getResidualEuler.  G[...,6]; y[...,19,N], z[...,6,N] are NOT modified (copies returned).
    Returns (res[...,6], y_new, z_new); z_new[..., N-1] keeps its input value (never written, :198-201)."""
    y = y.copy()
    z = z.copy()
    dt = y.dtype
    lead = y.shape[:-2]
    base = np.concatenate([np.broadcast_to(np.asarray(P.p0, dt), lead + (3,)),
                           np.broadcast_to(np.asarray(P.h0, dt), lead + (4,)),
                           G[..., 0:3], G[..., 3:6],
                           np.broadcast_to(np.asarray(P.q0, dt), lead + (3,)),
                           np.broadcast_to(np.asarray(P.w0, dt), lead + (3,))], -1)
    y[..., :, 0] = base
    tf = tendon_force(P, tensions)
    ds = dt.type(P.ds)
    for j in range(P.N - 1):
        yj = y[..., :, j]
        dyds, zj = ode(P, yj, yh[..., :, j], zh[..., :, j], tf, mlp, nn_input_history)
        z[..., :, j] = zj
        y[..., :, j + 1] = yj + ds * dyds
    nL = y[..., 7:10, -1]
    mL = y[..., 10:13, -1]
    res = np.concatenate([np.asarray(P.F_tip, dt) - nL, np.asarray(P.M_tip, dt) - mL], -1)
    return res, y, z


def march_rk4(P, G, y, z, yh, zh, tensions, mlp=None, nn_input_history=False):
    """getResidualRK4, cosserat_ode.py:215-255 (no caller in the reference).  Mid-point histories are the
    linear interpolations built in knode.py:80-81."""
    y = y.copy()
    z = z.copy()
    dt = y.dtype
    lead = y.shape[:-2]
    yh_int = 0.5 * (yh[..., :, :-1] + yh[..., :, 1:])
    zh_int = 0.5 * (zh[..., :, :-1] + zh[..., :, 1:])
    base = np.concatenate([np.broadcast_to(np.asarray(P.p0, dt), lead + (3,)),
                           np.broadcast_to(np.asarray(P.h0, dt), lead + (4,)),
                           G[..., 0:3], G[..., 3:6],
                           np.broadcast_to(np.asarray(P.q0, dt), lead + (3,)),
                           np.broadcast_to(np.asarray(P.w0, dt), lead + (3,))], -1)
    y[..., :, 0] = base
    tf = tendon_force(P, tensions)
    ds = dt.type(P.ds)
    for j in range(P.N - 1):
        yj = y[..., :, j]
        k1, zj = ode(P, yj, yh[..., :, j], zh[..., :, j], tf, mlp, nn_input_history)
        z[..., :, j] = zj
        k2, _ = ode(P, yj + k1 * ds / 2, yh_int[..., :, j], zh_int[..., :, j], tf, mlp, nn_input_history)
        k3, _ = ode(P, yj + k2 * ds / 2, yh_int[..., :, j], zh_int[..., :, j], tf, mlp, nn_input_history)
        k4, _ = ode(P, yj + k3 * ds, yh[..., :, j + 1], zh[..., :, j + 1], tf, mlp, nn_input_history)
        y[..., :, j + 1] = yj + ds * (k1 + 2 * (k2 + k3) + k4) / 6
    nL = y[..., 7:10, -1]
    mL = y[..., 10:13, -1]
    res = np.concatenate([np.asarray(P.F_tip, dt) - nL, np.asarray(P.M_tip, dt) - mL], -1)
    return res, y, z


# ----------------------------------------------------------------------------------------------------------
# time rollout (knode.simulate, knode.py:55-102)
# ----------------------------------------------------------------------------------------------------------


def initial_state(P, lead=(), dtype=np.float64):
    """knode.py:58-64 — straight rod."""
    N = P.N
    y = np.zeros(lead + (19, N), dtype)
    y[..., 2, :] = np.linspace(0, P.L, N)
    y[..., 3, :] = 1.0
    z = np.zeros(lead + (6, N), dtype)
    z[..., 2, :] = 1.0
    return y, z


def rollout_fsolve(P, ctl, mlp=None, nn_input_history=False, xtol=1.49012e-08):
    """Single rod, literally knode.simulate with scipy's hybrd: ctl[T,4] -> float64[T,50,N].
    The state after each solve is whatever the LAST residual evaluation left behind (in-place mutation,
    cosserat_ode.py:194,200-201 + knode.py:89,96)."""
    from scipy.optimize import fsolve

    y, z = initial_state(P)
    y_prev, z_prev = y.copy(), z.copy()
    G = np.zeros(6)
    trajectory = [np.vstack([y, z, y, z])]
    for controls in ctl:
        tens = np.array(controls, dtype=np.float64)
        yh = P.c1 * y + P.c2 * y_prev
        zh = P.c1 * z + P.c2 * z_prev
        y_prev, z_prev = y.copy(), z.copy()
        state = {}

        def resid(Gx):
            res, yn, zn = march_euler(P, Gx, y_prev, z_prev, yh, zh, tens, mlp, nn_input_history)
            state["y"], state["z"] = yn, zn
            return res

        G = fsolve(resid, G, xtol=xtol)
        y, z = state["y"], state["z"]
        trajectory.append(np.vstack([y, z, yh, zh]))
    return np.array(trajectory)[:-1]


def rollout_newton(P, ctl, mlp=None, nn_input_history=False, tol=1e-13, max_iter=30, rows=50,
                   return_info=False, method="euler"):
    """Batched rollout: ctl[B,T,4] -> float64[B,T,rows,N].  Same time loop as knode.simulate (knode.py:70-100) with
    the 6-unknown root found by Newton with a central-difference Jacobian instead of hybrd; agrees with
    `rollout_fsolve(xtol=1e-13)` to ~1e-13 (tests/test_oracle_golden.py).  method="rk4": the residual of the loop is
    getResidualRK4 (cosserat_ode.py:215-255; the loop already builds its mid-point histories, knode.py:80-81) — pinned to
    the reference's simulate run with that residual (tests/golden/make_rk4_rollout.py)."""
    march_euler = globals()["march_rk4" if method == "rk4" else "march_euler"]   # noqa: F841 — shadows the Euler march below
    ctl = np.asarray(ctl, dtype=np.float64)
    B, T, _ = ctl.shape
    y, z = initial_state(P, (B,))
    y_prev, z_prev = y.copy(), z.copy()
    G = np.zeros((B, 6))
    out = np.zeros((B, T, rows, P.N))
    out[:, 0, :19] = y
    out[:, 0, 19:25] = z
    if rows == 50:
        out[:, 0, 25:44] = y
        out[:, 0, 44:50] = z
    iters = np.zeros((B, T), np.int32)
    Gs = np.zeros((B, T, 6))
    for t in range(T - 1):  # the reference computes a T-th step and drops it (knode.py:102)
        tens = ctl[:, t]
        yh = P.c1 * y + P.c2 * y_prev
        zh = P.c1 * z + P.c2 * z_prev
        y_prev, z_prev = y, z
        for it in range(max_iter):
            res, yn, zn = march_euler(P, G, y_prev, z_prev, yh, zh, tens, mlp, nn_input_history)
            J = np.zeros((B, 6, 6))
            for k in range(6):
                e = 1e-6 * np.maximum(1.0, np.abs(G[:, k]))
                Gp, Gm = G.copy(), G.copy()
                Gp[:, k] += e
                Gm[:, k] -= e
                rp, _, _ = march_euler(P, Gp, y_prev, z_prev, yh, zh, tens, mlp, nn_input_history)
                rm, _, _ = march_euler(P, Gm, y_prev, z_prev, yh, zh, tens, mlp, nn_input_history)
                J[:, :, k] = (rp - rm) / (2 * e)[:, None]
            dG = np.linalg.solve(J, -res[..., None])[..., 0]
            if np.max(np.abs(res)) < tol:
                break
            G = G + dG
        iters[:, t + 1] = it + 1
        y, z = yn, zn
        Gs[:, t + 1] = G
        out[:, t + 1, :19] = y
        out[:, t + 1, 19:25] = z
        if rows == 50:
            out[:, t + 1, 25:44] = yh
            out[:, t + 1, 44:50] = zh
    if return_info:
        return out, Gs, iters
    return out


# ----------------------------------------------------------------------------------------------------------
# teacher-forced segment step (cosserat_ode_torch.py:370-399, :401-437)
# ----------------------------------------------------------------------------------------------------------


def next_segment_euler(P, Grod, yh, zh, tensions, mlp=None, nn_input_history=False):
    """getNextSegmentEuler: Grod[...,25,N] -> full_rod[...,25,N]; y,z are NOT propagated (:391)."""
    y = Grod[..., :19, :]
    z = Grod[..., 19:, :]
    tf = tendon_force(P, tensions)
    N = Grod.shape[-1]
    full = np.empty_like(Grod)
    full[..., :, 0] = np.concatenate([y[..., :, 0], z[..., :, 0]], -1)
    for j in range(N - 1):
        dy, zj = ode(P, y[..., :, j], yh[..., :, j], zh[..., :, j], tf, mlp, nn_input_history)
        full[..., :19, j + 1] = y[..., :, j] + Grod.dtype.type(P.ds) * dy
        full[..., 19:, j + 1] = zj
    return full


def parallel_next_segment_euler(P, Gs, segment_idxs, yh, zh, tensions, mlp=None, nn_input_history=False):
    """parallelGetNextSegmentEuler: Gs[S,25,N], yh[S,19,N], zh[S,6,N], tensions[S,4] -> [S,25,K]."""
    idx = np.asarray(segment_idxs) - 1
    y = np.swapaxes(Gs[:, :19, :], 1, 2)[:, idx]  # [S,K,19]
    yhk = np.swapaxes(yh, 1, 2)[:, idx]
    zhk = np.swapaxes(zh, 1, 2)[:, idx]
    tf = tendon_force(P, tensions)[:, None, :].repeat(len(idx), 1)
    dy, zj = ode(P, y, yhk, zhk, tf, mlp, nn_input_history)
    yn = y + Gs.dtype.type(P.ds) * dy
    return np.swapaxes(np.concatenate([yn, zj], -1), 1, 2)


# ----------------------------------------------------------------------------------------------------------
# loss (Utils/transformations.py:3-31, physics_train.py:345-352) and the training step
# ----------------------------------------------------------------------------------------------------------


def quaternion_to_euler(qt):
    """qt[4,...] -> [3,...] (roll, pitch, yaw) exactly as Utils/transformations.py:16-29."""
    nrm = np.sqrt(np.sum(qt * qt, axis=0, keepdims=True))
    w, x, y, z = qt / nrm
    roll = np.arctan2(2 * (w * y + x * z), 1 - 2 * (y ** 2 + z ** 2))
    pitch = np.arcsin(np.clip(2 * (w * z - x * y), -1.0, 1.0))
    yaw = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x ** 2 + z ** 2))
    return np.stack([roll, pitch, yaw], 0)


def quaternion_to_euler_vjp(qt, g):
    """Cotangent g[3,...] of quaternion_to_euler -> cotangent of qt[4,...] (what torch autograd computes;
    clamp passes gradient only strictly inside (-1,1))."""
    n2 = np.sum(qt * qt, axis=0, keepdims=True)
    nrm = np.sqrt(n2)
    w, x, y, z = qt / nrm
    gr, gp, gy = g
    # roll = atan2(a, b)
    a, b = 2 * (w * y + x * z), 1 - 2 * (y ** 2 + z ** 2)
    da, db = gr * b / (a * a + b * b), -gr * a / (a * a + b * b)
    gw = da * 2 * y
    gx = da * 2 * z
    gyq = da * 2 * w + db * (-4 * y)
    gz = da * 2 * x + db * (-4 * z)
    # pitch = asin(clamp(s))
    s = 2 * (w * z - x * y)
    inside = (s > -1.0) & (s < 1.0)
    ds_ = np.where(inside, gp / np.sqrt(np.maximum(1 - s * s, 1e-300)), 0.0)
    gw = gw + ds_ * 2 * z
    gz = gz + ds_ * 2 * w
    gx = gx - ds_ * 2 * y
    gyq = gyq - ds_ * 2 * x
    # yaw = atan2(c, d)
    c, d = 2 * (w * x + y * z), 1 - 2 * (x ** 2 + z ** 2)
    dc, dd = gy * d / (c * c + d * d), -gy * c / (c * c + d * d)
    gw = gw + dc * 2 * x
    gx = gx + dc * 2 * w + dd * (-4 * x)
    gyq = gyq + dc * 2 * z
    gz = gz + dc * 2 * y + dd * (-4 * z)
    gn = np.stack([gw, gx, gyq, gz], 0)  # cotangent of the normalised quaternion
    qn = qt / nrm
    return (gn - qn * np.sum(gn * qn, axis=0, keepdims=True)) / nrm


def teacher_forced_loss_and_grads(P, traj, controls, key_idx, mlp, batch_len=None, nn_input_history=False,
                                  want_grads=True):
    """One full-batch loss of physics_train.py's fast path (:313-368) over traj[B,T,25,N], controls[B,T,4]:
    samples (b, t in 0..T-2, k): ODE at the NEXT ground-truth state's node key[k]-1 (teacher forcing, :325),
    BDF2 history from steps t, t-1 (:321-322,330-331; t=0 uses y_prev=y), Euler step, 4-term MSE (:345-352),
    summed over b,t and divided by (batch_len-1) (:367-368).
    Gradients w.r.t. the four MLP tensors by hand-written reverse mode; the physics is a constant offset
    (SURVEY §0 fact 1).  Returns loss, dict(W1,b1,W2,b2), pred[B,T-1,25,K]."""
    traj = np.asarray(traj)
    dt = traj.dtype
    B, T, _, N = traj.shape
    if batch_len is None:
        batch_len = T
    key = np.asarray(key_idx)
    K = len(key)
    ys = traj[:, :batch_len - 1, :19]
    zs = traj[:, :batch_len - 1, 19:]
    y_prevs = np.concatenate([ys[:, :1], ys[:, :-1]], 1)
    z_prevs = np.concatenate([zs[:, :1], zs[:, :-1]], 1)
    yh = dt.type(P.c1) * ys + dt.type(P.c2) * y_prevs
    zh = dt.type(P.c1) * zs + dt.type(P.c2) * z_prevs
    Gs = traj[:, 1:batch_len]
    idx = key - 1
    y = np.swapaxes(Gs[:, :, :19], 2, 3)[:, :, idx]  # [B,S,K,19]
    yhk = np.swapaxes(yh, 2, 3)[:, :, idx]
    zhk = np.swapaxes(zh, 2, 3)[:, :, idx]
    tf = tendon_force(P, controls[:, :batch_len - 1])[:, :, None, :].repeat(K, 2)
    # forward with the MLP internals kept
    ys_phys, z_phys = ode(P, y, yhk, zhk, tf, None)
    if nn_input_history:
        x = np.concatenate([y, yhk, z_phys, zhk, tf], -1)
    else:
        x = np.concatenate([y, z_phys, tf], -1)
    W1, b1, W2, b2 = (mlp[k].astype(dt) for k in ("W1", "b1", "W2", "b2"))
    z1 = x @ W1.T + b1
    a1 = elu(z1)
    o = a1 @ W2.T + b2
    ds = dt.type(P.ds)
    y_next = y + ds * (ys_phys + o[..., :19])
    z_new = z_phys + o[..., 19:]
    pred = np.swapaxes(np.concatenate([y_next, z_new], -1), 2, 3)  # [B,S,25,K]
    nxt = traj[:, 1:batch_len]
    tgt_y = nxt[:, :, :19][:, :, :, key]  # [B,S,19,K]
    tgt_z = nxt[:, :, 19:][:, :, :, key - 1]
    S = batch_len - 1
    e_p = pred[:, :, :3] - tgt_y[:, :, :3]
    e_f = pred[:, :, 7:19] - tgt_y[:, :, 7:19]
    qp = np.moveaxis(pred[:, :, 3:7], 2, 0)  # [4,B,S,K]
    qtgt = np.moveaxis(tgt_y[:, :, 3:7], 2, 0)
    e_e = quaternion_to_euler(qp) - quaternion_to_euler(qtgt)  # [3,B,S,K]
    e_z = pred[:, :, 19:] - tgt_z
    loss = (np.sum(e_p ** 2) / (3 * K) + np.sum(e_f ** 2) / (12 * K) + np.sum(e_e ** 2) / (3 * K)
            + np.sum(e_z ** 2) / (6 * K)) / S
    if not want_grads:
        return loss, None, pred
    g_pred = np.zeros_like(pred)
    g_pred[:, :, :3] = 2 * e_p / (3 * K) / S
    g_pred[:, :, 7:19] = 2 * e_f / (12 * K) / S
    g_pred[:, :, 19:] = 2 * e_z / (6 * K) / S
    g_q = quaternion_to_euler_vjp(qp, 2 * e_e / (3 * K) / S)  # [4,B,S,K]
    g_pred[:, :, 3:7] = np.moveaxis(g_q, 0, 2)
    g_o = np.swapaxes(g_pred, 2, 3).copy()  # [B,S,K,25]
    g_o[..., :19] *= ds
    go = g_o.reshape(-1, 25)
    a1f = a1.reshape(-1, a1.shape[-1])
    z1f = z1.reshape(-1, z1.shape[-1])
    xf = x.reshape(-1, x.shape[-1])
    gW2 = go.T @ a1f
    gb2 = go.sum(0)
    ga1 = go @ W2
    gz1 = ga1 * np.where(z1f > 0, 1.0, np.exp(np.minimum(z1f, 0)))
    gW1 = gz1.T @ xf
    gb1 = gz1.sum(0)
    return loss, {"W1": gW1, "b1": gb1, "W2": gW2, "b2": gb2}, pred


def adam_clamp_step(params, grads, state, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                    clamp_weights=True):
    """torch.optim.Adam (L2 weight decay added to the gradient) followed by the reference's clamp of both Linear
    weights to >= 0 (physics_train.py:199,296-304).  state = dict(step, m{...}, v{...}), updated in place."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    new = {}
    for k in ("W1", "b1", "W2", "b2"):
        g = grads[k] + weight_decay * params[k]
        m = state.setdefault("m", {}).get(k, np.zeros_like(g))
        v = state.setdefault("v", {}).get(k, np.zeros_like(g))
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        state["m"][k], state["v"][k] = m, v
        mhat = m / (1 - betas[0] ** t)
        denom = np.sqrt(v) / np.sqrt(1 - betas[1] ** t) + eps
        p = params[k] - lr * mhat / denom
        if clamp_weights and k in ("W1", "W2"):
            p = np.maximum(p, 0)
        new[k] = p
    return new


# ---- evaluation metrics (physics_train.py:159, physics_multitrain.py:211-222) ---------------------------------------------
def dtw_l1(a, b):
    """Exact dynamic-time-warping distance with the L1 point distance, a[Ta,d], b[Tb,d] — the quantity
    fastdtw(a, b)[0] (physics_train.py:159, physics_multitrain.py:211) approximates with radius 1.  fastdtw is a
    third-party package the reference neither vendors nor pins; its published algorithm (Salvador & Chan, "FastDTW", 2007)
    is a multi-resolution approximation of exactly this recurrence, plain cell-by-cell loops here:
        acc[i,j] = |a_i - b_j|_1 + min(acc[i-1,j], acc[i,j-1], acc[i-1,j-1])."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    Ta, Tb = len(a), len(b)
    acc = np.full((Ta + 1, Tb + 1), np.inf)
    acc[0, 0] = 0.0
    for i in range(1, Ta + 1):
        for j in range(1, Tb + 1):
            d = a[i - 1] - b[j - 1]
            cost = (abs(d[0]) + abs(d[1])) + abs(d[2]) if d.shape[0] == 3 else np.abs(d).sum()
            acc[i, j] = cost + min(min(acc[i - 1, j], acc[i, j - 1]), acc[i - 1, j - 1])
    return float(acc[Ta, Tb])


def euler_zyx(q):
    """scipy Rotation.from_quat(q, scalar_first=True).as_euler('zyx') (physics_multitrain.py:216-217) in closed form,
    q[n,4] = (w,x,y,z): extrinsic rotations about z, y, x, i.e. R = Rx(c) Ry(b) Rz(a) -> [a, b, c]."""
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    r00, r01, r02 = 1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)
    r12, r22 = 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)
    return np.stack([np.arctan2(-r01, r00), np.arcsin(np.clip(r02, -1, 1)), np.arctan2(-r12, r22)], axis=1)


def pos_euler_mse(trajectory, interpolated):
    """physics_multitrain.py:213-222: 1000 * mean over [squared Euler-angle differences ; squared position errors]."""
    se_pos = (trajectory[:, :3] - interpolated[:, :3]).reshape((-1, 3)) ** 2
    eq = trajectory[:, 3:7].transpose((0, 2, 1)).reshape((-1, 4))
    rq = interpolated[:, 3:7].transpose((0, 2, 1)).reshape((-1, 4))
    se_euler = (euler_zyx(eq) - euler_zyx(rq)) ** 2
    return float(np.mean(np.concatenate([se_euler, se_pos])) * 1000)
