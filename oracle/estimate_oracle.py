"""TEST INFRASTRUCTURE — numpy fp64 restatement of the reference's state estimator:

    knode_cosserat_realworld/estimate_state.py:158-242   estimate_state(data[T,7,N], tensions[T,4], robot) -> [T,25,N]
    and its helpers :11-46 (spatial derivative of R through the matrix logarithm), :48-95 (v, u from p, h),
    :97-123 (angular velocities from quaternion pairs), :126-156 (n, m by a backward recursion from the tip).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path never does.
PARITY PINNED: tests/test_oracle_golden.py checks it against tests/golden/estimate_state.npz, which
tests/golden/make_estimate_state.py produced by calling the unmodified reference function.

Differences in *how* (not what): scipy.linalg.logm of a relative rotation is replaced by its closed form (the rotation
vector, taken from the relative quaternion, as a skew matrix); loops over time are vectorised except the one true
recurrence (v_prev, u_prev).  `P` is an oracle.rod_oracle.RodParams.
"""
import numpy as np


def quat_to_R(h):
    """estimate_state.py:70-77 (== cosserat_ode.py:131-135): h[...,4] (w,x,y,z), not necessarily unit -> R[...,3,3]."""
    h1, h2, h3, h4 = (h[..., k] for k in range(4))
    s = 2.0 / (h1 * h1 + h2 * h2 + h3 * h3 + h4 * h4)
    R = np.empty(h.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 + s * (-h3 * h3 - h4 * h4)
    R[..., 0, 1] = s * (h2 * h3 - h4 * h1)
    R[..., 0, 2] = s * (h2 * h4 + h3 * h1)
    R[..., 1, 0] = s * (h2 * h3 + h4 * h1)
    R[..., 1, 1] = 1 + s * (-h2 * h2 - h4 * h4)
    R[..., 1, 2] = s * (h3 * h4 - h2 * h1)
    R[..., 2, 0] = s * (h2 * h4 - h3 * h1)
    R[..., 2, 1] = s * (h3 * h4 + h2 * h1)
    R[..., 2, 2] = 1 + s * (-h2 * h2 - h3 * h3)
    return R


def relative_rotation_vector(h_cur, h_next):
    """Rotation vector phi with expm([phi]x) = R(h_next) R(h_cur)^T, i.e. the vee of logm(R_rel) at estimate_state.py:29-32.
    R(a (x) b) = R(a) R(b) and R(conj a) = R(a)^T, so R_rel = R(h_next (x) conj(h_cur))."""
    a0, a1, a2, a3 = (h_next[..., k] for k in range(4))
    b0, b1, b2, b3 = h_cur[..., 0], -h_cur[..., 1], -h_cur[..., 2], -h_cur[..., 3]
    w = a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3
    x = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2
    y = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1
    z = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0
    flip = np.where(w < 0, -1.0, 1.0)
    w, x, y, z = w * flip, x * flip, y * flip, z * flip
    s = np.sqrt(x * x + y * y + z * z)
    theta = 2.0 * np.arctan2(s, w)
    nrm = np.sqrt(w * w + s * s)
    k = np.where(s > 1e-12 * nrm, theta / np.maximum(s, 1e-300), 2.0 / nrm)
    return np.stack([k * x, k * y, k * z], -1)


def skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 2, 1], S[..., 1, 2] = v[..., 0], -v[..., 0]
    S[..., 0, 2], S[..., 2, 0] = v[..., 1], -v[..., 1]
    S[..., 1, 0], S[..., 0, 1] = v[..., 2], -v[..., 2]
    return S


def spatial_difference(x, arc):
    """estimate_state.py:64-68 / :138-142: forward difference along the nodes, last column repeated.  x[...,3,N]."""
    d = np.empty_like(x)
    d[..., :-1] = (x[..., 1:] - x[..., :-1]) / (arc[1:] - arc[:-1])
    d[..., -1] = d[..., -2]
    return d


def compute_v_u(p, h, arc):
    """estimate_state.py:48-95 for all time steps at once.  p[T,3,N], h[T,4,N] -> v[T,3,N], u[T,3,N], R[T,N,3,3], p_s."""
    T, _, N = p.shape
    p_s = spatial_difference(p, arc)
    R = quat_to_R(np.moveaxis(h, 1, 2))                                   # [T,N,3,3]
    hq = np.moveaxis(h, 1, 2)
    phi = relative_rotation_vector(hq[:, :-1], hq[:, 1:])                 # [T,N-1,3]
    ang = skew(phi) / (arc[1:] - arc[:-1])[None, :, None, None]           # :35-36
    R_s = np.empty((T, N, 3, 3))
    R_s[:, :-1] = R[:, :-1] @ ang                                         # :39
    R_s[:, -1] = R_s[:, -2]                                               # :42
    v = np.einsum("tnji,tjn->tin", R, p_s)                                # :82  R^T p_s
    u_hat = np.swapaxes(R, -1, -2) @ R_s                                  # :84
    u = np.stack([u_hat[..., 2, 1], u_hat[..., 0, 2], u_hat[..., 1, 0]], 1)   # :85-87  -> [T,3,N]
    v[:, 0:2, 0] = 0                                                      # :90-91
    v[:, 2, 0] = 1
    return v, u, R, p_s


def compute_angular_velocities(h, del_t):
    """estimate_state.py:97-123.  h[T,4,N] -> w[T,3,N]; w[t+1] from the pair (h[t], h[t+1]), w[0] = w[1]."""
    q1, q2 = h[:-1], h[1:]
    w = np.zeros((h.shape[0], 3, h.shape[2]))
    w[1:, 0] = q1[:, 0] * q2[:, 1] - q1[:, 1] * q2[:, 0] - q1[:, 2] * q2[:, 3] + q1[:, 3] * q2[:, 2]
    w[1:, 1] = q1[:, 0] * q2[:, 2] + q1[:, 1] * q2[:, 3] - q1[:, 2] * q2[:, 0] - q1[:, 3] * q2[:, 1]
    w[1:, 2] = q1[:, 0] * q2[:, 3] - q1[:, 1] * q2[:, 2] + q1[:, 2] * q2[:, 1] - q1[:, 3] * q2[:, 0]
    w *= 2 / del_t
    w[0] = w[1]
    return w


def internal_forces_and_moments(P, p_s, R, q, w, qt, wt, tensions):
    """estimate_state.py:126-156 for all time steps at once, loop over the nodes as written — including the hard-coded
    `if i != 9` (:147, :153) and the write to index N-i-2, which is index -1 (the tip) when i = N-1 != 9."""
    T, N = q.shape[0], P.N
    n = np.zeros((T, 3, N))
    m = np.zeros((T, 3, N))
    tf = tensions @ P.tendon_dirs                                          # [T,3]  :134
    for i in range(N):
        k = N - i - 1
        Rk = R[:, k]
        f = P.rhoAg - np.einsum("tij,tj->ti", Rk, P.C * q[:, :, k] * np.abs(q[:, :, k])) + tf          # :145
        ns = P.rhoA * np.einsum("tij,tj->ti", Rk, np.cross(w[:, :, k], q[:, :, k]) + qt[:, :, k]) - f  # :146
        if i != 9:
            n[:, :, k - 1] = n[:, :, k] - ns * P.L / N                                                  # :148
    for i in range(N):
        k = N - i - 1
        Rk = R[:, k]
        ms = np.einsum("tij,tj->ti", Rk, np.cross(w[:, :, k], w[:, :, k] @ P.rhoJ.T) + wt[:, :, k] @ P.rhoJ.T) \
            - np.cross(p_s[:, :, k], n[:, :, k])                                                        # :151-152
        if i != 9:
            m[:, :, k - 1] = m[:, :, k] - ms * P.L / N                                                  # :154
    return n, m


def estimate_state(P, data, tensions):
    """estimate_state.py:158-242.  data[T,7,N], tensions[T,4] -> [T,25,N] (float64)."""
    data = np.asarray(data, dtype=np.float64)
    tensions = np.asarray(tensions, dtype=np.float64)
    T, _, N = data.shape
    assert N == P.N
    arc = np.linspace(0, P.L, N)                                            # :167
    est = np.zeros((T, 25, N))
    est[:, :3] = data[:, :3]                                                # :174
    est[:, :2, 0] = 0                                                       # :175
    est[:, 3:7] = data[:, 3:7]                                              # :178
    vel = np.gradient(est[:, :3], P.del_t, axis=0, edge_order=1)            # :180
    est[:, 13:16] = vel
    ang = compute_angular_velocities(est[:, 3:7], P.del_t)                  # :183
    est[:, 16:19] = ang
    qt = np.gradient(vel, P.del_t, axis=0, edge_order=2)                    # :186
    wt = np.gradient(ang, P.del_t, axis=0, edge_order=2)                    # :187
    v_raw, u_raw, R, p_s = compute_v_u(est[:, :3], est[:, 3:7], arc)        # :196
    n, m = internal_forces_and_moments(P, p_s, R, vel, ang, qt, wt, tensions)   # :215-223
    est[:, 7:10, :-1] = n[:, :, :-1]                                        # :226-227 (tip rows stay 0)
    est[:, 10:13, :-1] = m[:, :, :-1]
    Rt_n = np.einsum("tnji,tjn->tin", R, est[:, 7:10])
    Rt_m = np.einsum("tnji,tjn->tin", R, est[:, 10:13])
    v_prev = u_prev = None
    for t in range(T):                                                      # the one recurrence: :229-240
        v, u = v_raw[t].copy(), u_raw[t].copy()
        if t == 0:
            v_prev, u_prev = v, u                                           # :197-199 — ALIASES of v, u (updated in place)
        for i in range(N):
            vh = P.c1 * v[:, i] + P.c2 * v_prev[:, i]
            uh = P.c1 * u[:, i] + P.c2 * u_prev[:, i]
            v[:, i] = P.Kse_plus_c0_Bse_inv @ (Rt_n[t, :, i] + P.Kse_vstar - P.Bse @ vh)     # :233
            u[:, i] = P.Kbt_plus_c0_Bbt_inv @ (Rt_m[t, :, i] - P.Bbt @ uh)                    # :234
        est[t, 19:22], est[t, 22:25] = v, u                                 # :236-237
        v_prev, u_prev = v, u                                               # :240-241
    est[:, 4:7, 0] = 0                                                      # :238
    return est
