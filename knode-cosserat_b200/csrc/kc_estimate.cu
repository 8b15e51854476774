// kc_estimate.cu — kc_estimate_state: the full 25-row state of a rod from measured positions and quaternions
// (knode_cosserat_realworld/estimate_state.py:158-242 and its helpers :11-156), batched over recordings.
//
// Two kernels.  (A) kc_estimate_kernel: everything that is local in time — finite differences in time (numpy.gradient,
// edge orders 1 and 2), quaternion -> R, the spatial derivative of R through the rotation vector of the relative
// rotation (the reference's scipy logm, in closed form from the relative quaternion), the backward recursions for n and
// m, and the constitutive re-estimate of v, u WITHOUT the previous time step's contribution.  One thread per (time step,
// node); a CTA owns a tile of consecutive time steps of one recording: the measurement rows t0-3 .. t0+TT+2 are staged
// in shared memory with contiguous loads, the [TT][25][N] result tile leaves through shared memory with contiguous
// stores (HBM traffic = algorithmic bytes + the 6-row halo).  (B) kc_estimate_recur_kernel: the one true recurrence of
// the reference (v_prev, u_prev: x_t = a_t + M x_{t-1}, M = -c2 (K + c0 B)^-1 B), one thread per (recording, node),
// in place on rows 19:25.
#include <cuda_runtime.h>
#include "kc_common.cuh"

namespace {

template <typename T>
struct EstC {
    T inv_dt, inv_ds, L_over_N;
    T Mv[9], Mu[9];   // recurrence matrices, see kernel B
    int T_len, TT, N;
};

template <typename T> KC_D T t_sqrt(T x);
template <> KC_D float t_sqrt<float>(float x) { return sqrtf(x); }
template <> KC_D double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> KC_D T t_atan2(T y, T x);
template <> KC_D float t_atan2<float>(float y, float x) { return atan2f(y, x); }
template <> KC_D double t_atan2<double>(double y, double x) { return atan2(y, x); }

template <typename T>
KC_D void quat_R(const T* h, T* R) {   // estimate_state.py:70-77
    const T h1 = h[0], h2 = h[1], h3 = h[2], h4 = h[3];
    const T s = T(2) / (h1 * h1 + h2 * h2 + h3 * h3 + h4 * h4);
    R[0] = T(1) + s * (-h3 * h3 - h4 * h4); R[1] = s * (h2 * h3 - h4 * h1); R[2] = s * (h2 * h4 + h3 * h1);
    R[3] = s * (h2 * h3 + h4 * h1); R[4] = T(1) + s * (-h2 * h2 - h4 * h4); R[5] = s * (h3 * h4 - h2 * h1);
    R[6] = s * (h2 * h4 - h3 * h1); R[7] = s * (h3 * h4 + h2 * h1); R[8] = T(1) + s * (-h2 * h2 - h3 * h3);
}
template <typename T> KC_D void mv(const T* M, const T* x, T* y) {
    y[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
    y[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
    y[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
}
template <typename T> KC_D void mtv(const T* M, const T* x, T* y) {
    y[0] = M[0] * x[0] + M[3] * x[1] + M[6] * x[2];
    y[1] = M[1] * x[0] + M[4] * x[1] + M[7] * x[2];
    y[2] = M[2] * x[0] + M[5] * x[1] + M[8] * x[2];
}
template <typename T> KC_D void cross3(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

// Rotation vector phi with expm([phi]x) = R(hn) R(hc)^T (vee of logm(R_rel), estimate_state.py:29-32): R(a (x) b) =
// R(a) R(b), R(conj a) = R(a)^T, so R_rel = R(hn (x) conj hc); phi = 2 atan2(|vec|, w) vec/|vec| — far better conditioned
// for the small node-to-node angles than the skew part of R_rel.
template <typename T>
KC_D void rel_rotvec(const T* hc, const T* hn, T* phi) {
    const T a0 = hn[0], a1 = hn[1], a2 = hn[2], a3 = hn[3], b0 = hc[0], b1 = -hc[1], b2 = -hc[2], b3 = -hc[3];
    T w = a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3;
    T x = a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2;
    T y = a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1;
    T z = a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0;
    if (w < T(0)) { w = -w; x = -x; y = -y; z = -z; }
    const T s = t_sqrt(x * x + y * y + z * z), nrm = t_sqrt(w * w + s * s);
    const T k = (s > T(1e-12) * nrm) ? T(2) * t_atan2(s, w) / s : T(2) / nrm;
    phi[0] = k * x; phi[1] = k * y; phi[2] = k * z;
}

// Staged measurements: row r of `in` is time t0 - 3 + r, each row [7][N].
template <typename T>
struct Meas {
    const T* in;
    int t0, N, Tlen;
    T inv_dt;
    KC_D T p(int t, int k, int j) const {   // estimate_state.py:174-175: base x, y forced to 0
        return (j == 0 && k < 2) ? T(0) : in[(size_t)(t - t0 + 3) * 7 * N + k * N + j];
    }
    KC_D T h(int t, int k, int j) const { return in[(size_t)(t - t0 + 3) * 7 * N + (3 + k) * N + j]; }
    // numpy.gradient(p, dt, axis=0, edge_order=1), estimate_state.py:180
    KC_D void vel(int t, int j, T* v) const {
        for (int k = 0; k < 3; ++k) {
            if (t == 0) v[k] = (p(1, k, j) - p(0, k, j)) * inv_dt;
            else if (t == Tlen - 1) v[k] = (p(t, k, j) - p(t - 1, k, j)) * inv_dt;
            else v[k] = (p(t + 1, k, j) - p(t - 1, k, j)) * (T(0.5) * inv_dt);
        }
    }
    // compute_angular_velocities, estimate_state.py:97-123: w[t] from the pair (h[t-1], h[t]), w[0] = w[1]
    KC_D void angvel(int t, int j, T* w) const {
        if (t == 0) t = 1;
        const T q10 = h(t - 1, 0, j), q11 = h(t - 1, 1, j), q12 = h(t - 1, 2, j), q13 = h(t - 1, 3, j);
        const T q20 = h(t, 0, j), q21 = h(t, 1, j), q22 = h(t, 2, j), q23 = h(t, 3, j);
        const T f = T(2) * inv_dt;
        w[0] = f * (q10 * q21 - q11 * q20 - q12 * q23 + q13 * q22);
        w[1] = f * (q10 * q22 + q11 * q23 - q12 * q20 - q13 * q21);
        w[2] = f * (q10 * q23 - q11 * q22 + q12 * q21 - q13 * q20);
    }
};

// numpy.gradient(f, dt, axis=0, edge_order=2) at time t for a quantity f(t') given by a functor (estimate_state.py:186-187)
template <typename T, typename F>
KC_D void grad2(int t, int Tlen, T inv_dt, F f, T* out) {
    T a[3], b[3], c[3];
    if (t == 0) {
        f(0, a); f(1, b); f(2, c);
        for (int k = 0; k < 3; ++k) out[k] = (T(-1.5) * a[k] + T(2) * b[k] - T(0.5) * c[k]) * inv_dt;
    } else if (t == Tlen - 1) {
        f(t - 2, a); f(t - 1, b); f(t, c);
        for (int k = 0; k < 3; ++k) out[k] = (T(0.5) * a[k] - T(2) * b[k] + T(1.5) * c[k]) * inv_dt;
    } else {
        f(t + 1, a); f(t - 1, b);
        for (int k = 0; k < 3; ++k) out[k] = (a[k] - b[k]) * (T(0.5) * inv_dt);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) kc_estimate_kernel(const __grid_constant__ RodC<T> c, const __grid_constant__ EstC<T> e,
                                                          int ntiles, const T* __restrict__ data,
                                                          const T* __restrict__ tens, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = e.N, TT = e.TT, Tlen = e.T_len;
    T* in_s = reinterpret_cast<T*>(smem_raw);            // [TT+6][7][N]
    T* out_s = in_s + (size_t)(TT + 6) * 7 * N;          // [TT][25][N]
    T* ns_s = out_s + (size_t)TT * 25 * N;               // [TT][3][N]  ns, then the R-term of ms
    T* ps_s = ns_s + (size_t)TT * 3 * N;                 // [TT][3][N]  p_s
    T* n_s = ps_s + (size_t)TT * 3 * N;                  // [TT][3][N]  n of the recursion (tip entry included)
    T* m_s = n_s + (size_t)TT * 3 * N;                   // [TT][3][N]

    const int64_t b = blockIdx.x / ntiles;
    const int t0 = (int)(blockIdx.x % ntiles) * TT;
    const int nt = min(TT, Tlen - t0);
    {   // stage rows t0-3 .. t0+nt+2 (clipped to the recording): one contiguous block of the input
        const int lo = max(t0 - 3, 0), hi = min(t0 + nt + 3, Tlen);
        const T* src = data + ((size_t)b * Tlen + lo) * 7 * N;
        T* dst = in_s + (size_t)(lo - (t0 - 3)) * 7 * N;
        const int cnt = (hi - lo) * 7 * N;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int tl = threadIdx.x / N, j = threadIdx.x - tl * N;
    const bool active = tl < nt;
    const int t = t0 + tl;
    const Meas<T> M{in_s, t0, N, Tlen, e.inv_dt};
    T R[9], q[3], w[3], vraw[3], uraw[3], hq[4], pp[3];
    if (active) {
        for (int k = 0; k < 4; ++k) hq[k] = M.h(t, k, j);
        for (int k = 0; k < 3; ++k) pp[k] = M.p(t, k, j);
        quat_R(hq, R);
        M.vel(t, j, q);
        M.angvel(t, j, w);
        T qt[3], wt[3];
        grad2(t, Tlen, e.inv_dt, [&](int tt, T* o) { M.vel(tt, j, o); }, qt);
        grad2(t, Tlen, e.inv_dt, [&](int tt, T* o) { M.angvel(tt, j, o); }, wt);
        // p_s (estimate_state.py:64-68) and the rotation vector to the next node; the last node repeats its neighbour's
        const int jj = (j < N - 1) ? j : N - 2;
        T ps[3], hc[4], hn[4], phi[3];
        for (int k = 0; k < 3; ++k) ps[k] = (M.p(t, k, jj + 1) - M.p(t, k, jj)) * e.inv_ds;
        for (int k = 0; k < 4; ++k) { hc[k] = M.h(t, k, jj); hn[k] = M.h(t, k, jj + 1); }
        rel_rotvec(hc, hn, phi);
        for (int k = 0; k < 3; ++k) phi[k] *= e.inv_ds;
        if (j < N - 1) {   // u_hat = R^T (R [phi]x / ds) = [phi]x / ds   (:39, :84-87)
            for (int k = 0; k < 3; ++k) uraw[k] = phi[k];
        } else {           // R_s[N-1] = R_s[N-2] (:42): u_hat = R_{N-1}^T R_{N-2} [phi]x / ds, entries (2,1), (0,2), (1,0)
            T Rc[9], A[9], S[9] = {T(0), -phi[2], phi[1], phi[2], T(0), -phi[0], -phi[1], phi[0], T(0)};
            quat_R(hc, Rc);
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) A[r * 3 + cc] = Rc[r * 3] * S[cc] + Rc[r * 3 + 1] * S[3 + cc] + Rc[r * 3 + 2] * S[6 + cc];
            uraw[0] = R[2] * A[1] + R[5] * A[4] + R[8] * A[7];   // (R^T A)[2][1]
            uraw[1] = R[0] * A[2] + R[3] * A[5] + R[6] * A[8];   // (R^T A)[0][2]
            uraw[2] = R[1] * A[0] + R[4] * A[3] + R[7] * A[6];   // (R^T A)[1][0]
        }
        mtv(R, ps, vraw);                                         // :82
        if (j == 0) { vraw[0] = T(0); vraw[1] = T(0); vraw[2] = T(1); }   // :90-91
        // ns and the R-term of ms (estimate_state.py:145-146, :151)
        const T* tn = tens + ((size_t)b * Tlen + t) * 4;
        T tf[3], d[3], Rd[3], x[3], Rx[3];
        for (int k = 0; k < 3; ++k) tf[k] = tn[0] * c.tdirs[k] + tn[1] * c.tdirs[3 + k] + tn[2] * c.tdirs[6 + k] + tn[3] * c.tdirs[9 + k];
        for (int k = 0; k < 3; ++k) d[k] = c.C[k] * q[k] * fabs(q[k]);
        mv(R, d, Rd);
        cross3(w, q, x);
        for (int k = 0; k < 3; ++k) x[k] += qt[k];
        mv(R, x, Rx);
        T Jw[3], Jwt[3], wJw[3];
        mv(c.rhoJ, w, Jw); mv(c.rhoJ, wt, Jwt);
        cross3(w, Jw, wJw);
        for (int k = 0; k < 3; ++k) x[k] = wJw[k] + Jwt[k];
        T Rm[3];
        mv(R, x, Rm);
        for (int k = 0; k < 3; ++k) {
            const T f = c.rhoAg[k] - Rd[k] + tf[k];
            ns_s[(tl * 3 + k) * N + j] = c.rhoA * Rx[k] - f;
            ps_s[(tl * 3 + k) * N + j] = ps[k];
            m_s[(tl * 3 + k) * N + j] = Rm[k];      // parked here until the m recursion starts
            n_s[(tl * 3 + k) * N + j] = T(0);
        }
    }
    __syncthreads();
    if (active && j == 0) {
        // compute_internal_forces_and_moments (estimate_state.py:144-154) as written: backwards from the tip, step L/N,
        // the write skipped at loop index 9 whatever N is, and index N-i-2 = -1 wrapping to the tip when i = N-1 != 9.
        T* nn = n_s + tl * 3 * N; T* mm = m_s + tl * 3 * N; const T* nsv = ns_s + tl * 3 * N; const T* psv = ps_s + tl * 3 * N;
        for (int i = 0; i < N; ++i) {
            const int k = N - 1 - i, dst = (k == 0) ? N - 1 : k - 1;
            if (i != 9)
                for (int a = 0; a < 3; ++a) nn[a * N + dst] = nn[a * N + k] - nsv[a * N + k] * e.L_over_N;
        }
        // m in place: mm[k] holds the R-term of node k until node k is visited; mt carries m[k] down the rod
        T mt[3] = {T(0), T(0), T(0)};
        for (int i = 0; i < N; ++i) {
            const int k = N - 1 - i;
            const T pk[3] = {psv[k], psv[N + k], psv[2 * N + k]}, nk[3] = {nn[k], nn[N + k], nn[2 * N + k]};
            T pxn[3];
            cross3(pk, nk, pxn);
            for (int a = 0; a < 3; ++a) {
                const T ms = mm[a * N + k] - pxn[a];
                mm[a * N + k] = mt[a];                                        // m[k] is final
                mt[a] = (i != 9) ? mt[a] - ms * e.L_over_N : T(0);            // m[k-1]; stays 0 when the write is skipped
            }
        }   // (for N != 10 the last write wraps to the tip entry, which is never read: rows 10:13 of the tip stay 0)
    }
    __syncthreads();
    if (active) {
        T nj[3], mj[3];
        for (int k = 0; k < 3; ++k) {   // :226-227: the tip keeps n = m = 0
            nj[k] = (j < N - 1) ? n_s[(tl * 3 + k) * N + j] : T(0);
            mj[k] = (j < N - 1) ? m_s[(tl * 3 + k) * N + j] : T(0);
        }
        // constitutive re-estimate (:231-234) without the previous step's term; at t = 0 v_prev aliases v (:197-199)
        const T ch = (t == 0) ? c.c1 + c.c2 : c.c1;
        T Rn[3], Rmm[3], Bv[3], Bu[3], rv[3], ru[3], av[3], au[3];
        mtv(R, nj, Rn); mtv(R, mj, Rmm);
        mv(c.Bse, vraw, Bv); mv(c.Bbt, uraw, Bu);
        for (int k = 0; k < 3; ++k) { rv[k] = Rn[k] + c.KseVstar[k] - ch * Bv[k]; ru[k] = Rmm[k] - ch * Bu[k]; }
        mv(c.KseInv, rv, av); mv(c.KbtInv, ru, au);
        T* o = out_s + (size_t)tl * 25 * N + j;
        for (int k = 0; k < 3; ++k) o[k * N] = pp[k];
        o[3 * N] = hq[0];
        for (int k = 1; k < 4; ++k) o[(3 + k) * N] = (j == 0) ? T(0) : hq[k];   // :238
        for (int k = 0; k < 3; ++k) {
            o[(7 + k) * N] = nj[k]; o[(10 + k) * N] = mj[k]; o[(13 + k) * N] = q[k]; o[(16 + k) * N] = w[k];
            o[(19 + k) * N] = av[k]; o[(22 + k) * N] = au[k];
        }
    }
    __syncthreads();
    {
        T* dst = out + ((size_t)b * Tlen + t0) * 25 * N;
        const int cnt = nt * 25 * N;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = out_s[i];
    }
}

// x_t = a_t + M x_{t-1} for t >= 1, in place on rows 19:22 (v, M = Mv) and 22:25 (u, M = Mu); x_0 is already final.
template <typename T>
__global__ void __launch_bounds__(128) kc_estimate_recur_kernel(const __grid_constant__ EstC<T> e, int64_t B, T* __restrict__ out) {
    const int N = e.N, Tlen = e.T_len;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * N) return;
    const int64_t b = g / N;
    const int j = (int)(g - b * N);
    T* base = out + (size_t)b * Tlen * 25 * N + 19 * N + j;
    const size_t step = (size_t)25 * N;
    T v[3], u[3];
    for (int k = 0; k < 3; ++k) { v[k] = base[k * N]; u[k] = base[(3 + k) * N]; }
    constexpr int U = 8;
    for (int t = 1; t < Tlen; t += U) {
        T a[U][6];
#pragma unroll
        for (int s = 0; s < U; ++s)
            if (t + s < Tlen)
#pragma unroll
                for (int k = 0; k < 6; ++k) a[s][k] = base[(size_t)(t + s) * step + k * N];
#pragma unroll
        for (int s = 0; s < U; ++s)
            if (t + s < Tlen) {
                T nv[3], nu[3];
                mv(e.Mv, v, nv); mv(e.Mu, u, nu);
#pragma unroll
                for (int k = 0; k < 3; ++k) { v[k] = a[s][k] + nv[k]; u[k] = a[s][3 + k] + nu[k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) { base[(size_t)(t + s) * step + k * N] = v[k]; base[(size_t)(t + s) * step + (3 + k) * N] = u[k]; }
            }
    }
}

void mat3(const double* A, const double* B, double s, double* out) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[r * 3 + c] = s * (A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c]);
}

template <typename T>
int launch(const kc_rod_params* P, double L, double del_t, int64_t B, int64_t Tlen, const void* data, const void* tens,
           void* out, cudaStream_t st) {
    const int N = P->N;
    EstC<T> e;
    e.inv_dt = (T)(1.0 / del_t);
    e.inv_ds = (T)((N - 1) / L);
    e.L_over_N = (T)(L / N);
    double Mv[9], Mu[9];
    mat3(P->Kse_c0Bse_inv, P->Bse, -P->c2, Mv);
    mat3(P->Kbt_c0Bbt_inv, P->Bbt, -P->c2, Mu);
    bool recur = false;
    for (int i = 0; i < 9; ++i) { e.Mv[i] = (T)Mv[i]; e.Mu[i] = (T)Mu[i]; recur = recur || Mv[i] != 0 || Mu[i] != 0; }
    e.T_len = (int)Tlen; e.N = N;
    e.TT = 256 / N > 0 ? 256 / N : 1;
    if (e.TT > Tlen) e.TT = (int)Tlen;
    const int threads = ((e.TT * N + 31) / 32) * 32;
    const size_t smem = sizeof(T) * (size_t)N * ((size_t)(e.TT + 6) * 7 + (size_t)e.TT * (25 + 12));
    const int ntiles = (int)((Tlen + e.TT - 1) / e.TT);
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "fp32 / fp64");
    if (smem > 48 * 1024) {
        cudaError_t er = cudaFuncSetAttribute(kc_estimate_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (er != cudaSuccess) { kc_set_error("kc_estimate_state: %zu bytes of shared memory: %s", smem, cudaGetErrorString(er)); return KC_ECUDA; }
    }
    const RodC<T> c = make_rodc<T>(*P);
    kc_estimate_kernel<T><<<(unsigned)(B * ntiles), threads, smem, st>>>(c, e, ntiles, (const T*)data, (const T*)tens, (T*)out);
    KC_CHECK_LAUNCH("kc_estimate_kernel");
    if (recur && Tlen > 1) {
        const int64_t n = B * N;
        kc_estimate_recur_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(e, B, (T*)out);
        KC_CHECK_LAUNCH("kc_estimate_recur_kernel");
    }
    return KC_OK;
}

}  // namespace

extern "C" int kc_estimate_state(int dtype, const kc_rod_params* P, double L, double del_t, int64_t B, int64_t T,
                                 const void* data, const void* tensions, void* est, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2 && P->N <= 256, "rod params missing or N outside [2, 256]");
    KC_CHECK_ARG(L > 0 && del_t > 0, "L and del_t must be positive");
    KC_CHECK_ARG(B >= 0 && T >= 3 && T < (1 << 30), "need B >= 0 and T >= 3 (numpy.gradient with edge_order=2, estimate_state.py:186)");
    KC_CHECK_ARG(B * ((T + 0) / 1) < ((int64_t)1 << 31), "B*T too large for one launch");
    if (B == 0) return KC_OK;
    KC_CHECK_ARG(data && tensions && est, "NULL data / tensions / est pointer");
    return dtype == KC_F32 ? launch<float>(P, L, del_t, B, T, data, tensions, est, (cudaStream_t)stream)
                           : launch<double>(P, L, del_t, B, T, data, tensions, est, (cudaStream_t)stream);
}
