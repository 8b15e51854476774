// kc_rollout.cu — batched time rollout of the Cosserat rod / KNODE model (knode.simulate, knode.py:55-102).
//
// Mapping: ONE ROD PER THREAD.  The 19 state values of the node being marched, the 6 base reactions, the 6x6 inverse
// shooting Jacobian and every temporary of the node ODE live in registers; the BDF2 history of the step being solved
// (12 values per node: qh, wh, vh, uh — the only history rows the physics reads) lives in shared memory laid out
// [slot][lane] so every access is conflict-free; the trajectory goes to HBM in a rod-fastest "device layout"
// [T][25*N][Bpad] so that each of a warp's stores is one full 128-byte line, and is read back (L2 hits) to form the
// next step's history.  A second kernel transposes the device layout into the reference's [B][T][rows][N] through a
// padded shared-memory tile (coalesced on both sides) and synthesises rows 25:50 (yh, zh) when asked.
//
// Roofline: the rollout is FP32/FP64-pipe and latency bound, not HBM bound: per rod-node-step it moves
// 100 B (fp32) to HBM against several kFLOP of dependent arithmetic.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "kc_rollout_core.cuh"
#include "kc_rollout_wide.cuh"
#include "kc_mlp_coop.cuh"
#include <type_traits>

// kc_knode_tc.cu: KNODE march with the MLP on tcgen05 (fp32, 28 inputs, hidden <= 512, Euler march)
bool kc_knode_tc_eligible(int dtype, const kc_mlp* mlp, int N, int method);
size_t kc_knode_tc_img_bytes();
int kc_knode_tc_fwd(const RodC<float>& P, const kc_mlp* mlp, int64_t B, int T_, const float* tensions, const float* y0,
                    const float* z0, float* trajD, float tol, int max_iter, float fd_eps, float* Gout, int32_t* iters,
                    unsigned char* img, unsigned char* scratch, cudaStream_t st);
size_t kc_knode_tc_fwd_scratch_bytes(int N, int64_t B);

constexpr int KC_LS = 32;  // lane stride of every per-rod array (one warp-wide tile)

// A rod's base pointer inside the tiled device layout [tile][T][N][25][32].
template <typename T>
__device__ __forceinline__ T* rod_base(T* trajD, int64_t b, int T_, int N) {
    return trajD + (size_t)(b >> 5) * ((size_t)T_ * N * 25 * KC_LS) + (b & 31);
}

// One warp per CTA; `rpw` (rods per warp, 1..32) lanes are active.  With few rods the launcher spreads them over more
// warps (the kernel is latency bound: a half-empty warp costs nothing, a longer per-warp critical path does).
template <typename T, bool DIAG, int IN, int NH, int METHOD = KC_MARCH_EULER>
__global__ void __launch_bounds__(32)
kc_rollout_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t B, int T_, int rpw,
                  const T* __restrict__ tensions, const T* __restrict__ y0, const T* __restrict__ z0, T* trajD,
                  size_t Bpad, T* state, int t_begin, int t_end, T tol, int max_iter, T fd_eps, T* Gout,
                  int32_t* iters) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = P.N;
    // shared memory: [N-1][NH][32] history ([N] for the RK4 march), then [KC_SHOOT_SLOTS][32] solver state
    T* Hs = reinterpret_cast<T*>(kc_smem) + threadIdx.x;
    const int hn = METHOD == KC_MARCH_RK4 ? N : N - 1;
    const ShootMem<T, KC_LS> st{reinterpret_cast<T*>(kc_smem) + (size_t)NH * hn * KC_LS + threadIdx.x};
    const int64_t b = (int64_t)blockIdx.x * rpw + threadIdx.x;
    if ((int)threadIdx.x >= rpw || b >= B) return;
    T* traj_b = rod_base(trajD, b, T_, N);
    T* sb = state + b;  // state workspace is [KC_SHOOT_SLOTS][Bpad]
    if (t_begin == 0) {
        st.reset();
        rollout_init<T, KC_LS>(P, y0 ? y0 + (size_t)b * 19 * N : nullptr, z0 ? z0 + (size_t)b * 6 * N : nullptr, traj_b);
        if (Gout) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = T(0);
        }
        if (iters) iters[(size_t)b * T_] = 0;
    } else {
        for (int i = 0; i < KC_SHOOT_SLOTS; ++i) st.p[i * KC_LS] = sb[(size_t)i * Bpad];
    }
    rollout_rod<T, DIAG, IN, NH, KC_LS, KC_LS, METHOD>(P, M, st, tensions + (size_t)b * T_ * 4, traj_b, Hs, t_begin, t_end, tol,
                                        max_iter, fd_eps, Gout ? Gout + (size_t)b * T_ * 6 : nullptr,
                                        iters ? iters + (size_t)b * T_ : nullptr);
    for (int i = 0; i < KC_SHOOT_SLOTS; ++i) sb[(size_t)i * Bpad] = st.p[i * KC_LS];
}

// KNODE rollout, warp-cooperative: ONE ROD PER WARP.  Every lane runs the same per-rod code on the same rod (physics
// redundantly, shared-memory state at lane-independent addresses, identical stores), the MLP inside each node evaluation
// is split over the lanes (kc_mlp_coop.cuh).  Chosen when the MLP is in the march and the batch is small.
constexpr int KC_COOP_WARPS = 8;   // rods (warps) per CTA: they share one shared-memory copy of the MLP weights
template <typename T, bool DIAG, int IN, int NH>
__global__ void __launch_bounds__(32 * KC_COOP_WARPS, 1)
kc_rollout_coop_kernel(const __grid_constant__ RodC<T> P, MlpCoop<T> M, int64_t B, int T_,
                       const T* __restrict__ tensions, const T* __restrict__ y0, const T* __restrict__ z0, T* trajD,
                       T tol, int max_iter, T fd_eps, T* Gout, int32_t* iters, int wc_elems) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = P.N;
    // the packed weights (110 KB at H = 512, fp32) are staged once per CTA when they fit (wc_elems != 0): every node
    // evaluation of every march re-reads all of them, and from L1/L2 a quarter of those reads missed L1
    T* Wsm = reinterpret_cast<T*>(kc_smem);
    if (wc_elems) {
        for (int e = threadIdx.x; e < wc_elems; e += blockDim.x) Wsm[e] = M.Wc[e];
        __syncthreads();
        M.Wc = Wsm;
    }
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * KC_COOP_WARPS + warp;
    if (b >= B) return;
    const int per_warp = NH * (N - 1) + KC_SHOOT_SLOTS;     // lane stride 1: the warp owns one rod
    T* Hs = Wsm + wc_elems + (size_t)warp * per_warp;
    const ShootMem<T, 1> st{Hs + (size_t)NH * (N - 1)};
    T* traj_b = rod_base(trajD, b, T_, N);
    st.reset();
    rollout_init<T, KC_LS>(P, y0 ? y0 + (size_t)b * 19 * N : nullptr, z0 ? z0 + (size_t)b * 6 * N : nullptr, traj_b);
    if (Gout) {
#pragma unroll
        for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = T(0);
    }
    if (iters) iters[(size_t)b * T_] = 0;
    __syncwarp();
    rollout_rod<T, DIAG, IN, NH, KC_LS, 1>(P, M, st, tensions + (size_t)b * T_ * 4, traj_b, Hs, 0, T_ - 1, tol, max_iter,
                                           fd_eps, Gout ? Gout + (size_t)b * T_ * 6 : nullptr,
                                           iters ? iters + (size_t)b * T_ : nullptr);
}

// Wide mode: 4 rods per warp, 8 lanes per rod (see kc_rollout_wide.cuh).  Shared memory: history [N-1][NH][4].
constexpr int KC_WG = 4;  // rods (lane groups) per warp
template <typename T, bool DIAG, int IN, int NH>
__global__ void __launch_bounds__(32)
kc_rollout_wide_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t B, int T_,
                       const T* __restrict__ tensions, const T* __restrict__ y0, const T* __restrict__ z0, T* trajD,
                       T tol, int max_iter, T fd_eps, T* Gout, int32_t* iters) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = P.N;
    const int lane = threadIdx.x, g = lane >> 3, k = lane & 7;
    const unsigned full = 0xffffffffu;
    const int64_t b_raw = (int64_t)blockIdx.x * KC_WG + g;
    const bool valid = b_raw < B;
    const int64_t b = valid ? b_raw : B - 1;  // surplus groups shadow the last rod and never store
    T* Hs = reinterpret_cast<T*>(kc_smem) + g;
    T* traj_b = rod_base(trajD, b, T_, N);
    const size_t tstride = (size_t)25 * N * KC_LS;
    if (k == 0 && valid) {
        rollout_init<T, KC_LS>(P, y0 ? y0 + (size_t)b * 19 * N : nullptr, z0 ? z0 + (size_t)b * 6 * N : nullptr, traj_b);
        if (Gout) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = T(0);
        }
        if (iters) iters[(size_t)b * T_] = 0;
    }
    __syncwarp();
    // cooperative history build H = c1*state[t] + c2*state[t-1] from the trajectory just written (L2 hits): the 8 lanes of
    // a group split the slots; all loads of up to 4 nodes are issued before the first use so their latencies overlap.
    auto build_hist = [&](const T* cur, const T* prev) {
        constexpr int SPL = (NH + 7) / 8;  // slots per lane
        for (int j0 = 0; j0 < N - 1; j0 += 4) {
            T a[4][SPL], c[4][SPL];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const int s = k + 8 * q, j = j0 + jj;
                    const bool ok = s < NH && j < N - 1;
                    const size_t o = ((size_t)(ok ? j : 0) * 25 + slot_row<NH>(ok ? s : 0)) * KC_LS;
                    a[jj][q] = cur[o];
                    c[jj][q] = prev[o];
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const int s = k + 8 * q, j = j0 + jj;
                    if (s < NH && j < N - 1) Hs[(size_t)(j * NH + s) * KC_WG] = P.c1 * a[jj][q] + P.c2 * c[jj][q];
                }
            }
        }
    };
    build_hist(traj_b, traj_b);
    T zlast[6];  // z[:, N-1] is never written by the march: constant over the whole rollout (cosserat_ode.py:198-201)
#pragma unroll
    for (int c = 0; c < 6; ++c) zlast[c] = traj_b[((size_t)(N - 1) * 25 + 19 + c) * KC_LS];
    __syncwarp();
    T G[6], Gm1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { G[i] = T(0); Gm1[i] = T(0); }
    const T* ten = tensions + (size_t)b * T_ * 4;
    T tn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tn[i] = ten[i];
    for (int t = 0; t < T_ - 1; ++t) {
        T tf[3];
        tendon_force(P, tn, tf);
        if (t + 1 < T_ - 1) {  // prefetch the next step's tensions: their latency hides behind this step's marches
#pragma unroll
            for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)(t + 1) * 4 + i];
        }
        T* nxt = traj_b + (size_t)(t + 1) * tstride;
        T Gp[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { Gp[i] = G[i]; G[i] = G[i] + (G[i] - Gm1[i]); }   // linear predictor
        bool done = false;
        int status = 0, marches = 0;
        HistView<T, NH, KC_WG> H{Hs};
        while (true) {
            T eps[6], Ge[6], F[6];
            wide_eps(G, fd_eps, eps);
#pragma unroll
            for (int i = 0; i < 6; ++i) Ge[i] = G[i] + ((k == i + 1) ? eps[i] : T(0));
            const bool on = k == 0 && !done;
            TrajSinkPred<T, KC_LS, NH, KC_WG> S{nxt, nullptr, on && valid, N - 1};
            rod_march<T, DIAG, IN, NH>(P, M, Ge, tf, H, S, F);
            T Fall[7][6];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Fall[c][i] = __shfl_sync(full, F[i], (lane & ~7) | c);
            }
            if (!done) {
                ++marches;
                const int r = wide_decide(Fall, G, eps, tol);
                if (r != 0) { done = true; status = r; }
                else if (marches >= max_iter) { done = true; status = -1; }
            }
            if (__all_sync(full, done)) break;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) Gm1[i] = Gp[i];
        if (k == 0 && valid) {
            const size_t o = (size_t)(N - 1) * 25 * KC_LS;
#pragma unroll
            for (int c = 0; c < 6; ++c) nxt[o + (19 + c) * KC_LS] = zlast[c];
            if (Gout) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Gout[((size_t)b * T_ + t + 1) * 6 + i] = G[i];
            }
            if (iters) iters[(size_t)b * T_ + t + 1] = status > 0 ? marches : -marches;
        }
        __syncwarp();
        build_hist(nxt, nxt - tstride);
        __syncwarp();
    }
}

// Wide mode with the linearised final correction (kc_rollout_wide.cuh): the 7 marched states of a rod live in shared
// memory, the accepted state is written STRAIGHT into the reference layout traj[B][T][rows][N] by the 8 lanes of the
// group (no device-layout trajectory, no transpose pass), and the next history is formed in shared memory.
// Shared memory per warp, all in "output order" e = r*N + j (row r, node j — the order of one [25][N] time slice):
//   S [25*N][28 slots]  marched states (slot = 7*rod + point)      H, A [NH*N][4 rods]  history / previous accepted state
// NC = compile-time node count (0: use P.N) — with NC all shared-memory offsets are immediates.
constexpr int KC_WS = 28;  // state slots per warp: 4 rods x 7 points
template <typename T, bool DIAG, int IN, int NH, int NC>
__global__ void __launch_bounds__(32)
kc_rollout_wide_lin_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t B, int T_,
                           const T* __restrict__ tensions, const T* __restrict__ y0, const T* __restrict__ z0,
                           T* out, int64_t tstride, T* gstate, size_t Bpad, int t_begin, int t_end, T tol, int max_iter,
                           T fd_eps, T* Gout, int32_t* iters) {
    // steps t in [t_begin, t_end) are solved (time indices t_begin+1 .. t_end are written; t_begin == 0 also writes
    // index 0).  A later launch resumes from the trajectory itself plus gstate[12][Bpad] = the last two roots.
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = NC ? NC : P.N;
    const int NV = 25 * N;
    const int H0 = (NH == 12) ? 13 * N : 0;   // first element that has a history (rows 13..24, or all rows)
    const int lane = threadIdx.x, g = lane >> 3, k = lane & 7;
    const unsigned full = 0xffffffffu;
    const int64_t b_raw = (int64_t)blockIdx.x * KC_WG + g;
    const bool valid = b_raw < B;
    const int64_t b = valid ? b_raw : B - 1;
    T* Sall = reinterpret_cast<T*>(kc_smem);
    T* Sg = Sall + g * 7;                       // this rod's 7 slots
    T* Hs = Sall + (size_t)NV * KC_WS + g;
    T* As = Hs + (size_t)NH * N * KC_WG;
    T* out_b = out + (size_t)b * T_ * tstride;
    // initial state -> time index 0, A, H and EVERY slot of S (rows 19..24 of the tip node are never written by the
    // march, cosserat_ode.py:198-201: they keep the initial z for the whole rollout)
    for (int e = k; e < NV; e += 8) {
        const int r = e / N, j = e - r * N;
        T v, vp;
        if (t_begin == 0) {
            if (r < 19) v = y0 ? y0[(size_t)b * 19 * N + e] : ((r == 2) ? P.ds * T(j) : (r == 3 ? T(1) : T(0)));
            else v = z0 ? z0[(size_t)b * 6 * N + (e - 19 * N)] : ((r == 21) ? T(1) : T(0));
            vp = v;                                             // state[-1] := state[0] (knode.py:65-66)
            if (valid) out_b[e] = v;
        } else {
            v = out_b[(size_t)t_begin * tstride + e];
            vp = out_b[(size_t)(t_begin - 1) * tstride + e];
        }
        if (e >= H0) {
            As[(size_t)(e - H0) * KC_WG] = v;
            Hs[(size_t)(e - H0) * KC_WG] = P.c1 * v + P.c2 * vp;
        }
#pragma unroll
        for (int c = 0; c < 7; ++c) Sg[(size_t)e * KC_WS + c] = v;
    }
    if (k == 0 && valid && t_begin == 0) {
        if (Gout) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = T(0);
        }
        if (iters) iters[(size_t)b * T_] = 0;
    }
    __syncwarp();
    T G[6], Gm1[6], D[6], Dm1[6];   // D: root change expected from the coming tension change (wide_decide_lin)
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        G[i] = t_begin == 0 ? T(0) : gstate[(size_t)i * Bpad + b];
        Gm1[i] = t_begin == 0 ? T(0) : gstate[(size_t)(6 + i) * Bpad + b];
        D[i] = t_begin == 0 ? T(0) : gstate[(size_t)(12 + i) * Bpad + b];
        Dm1[i] = t_begin == 0 ? T(0) : gstate[(size_t)(18 + i) * Bpad + b];
    }
    T Cest = T(0);
    const T* ten = tensions + (size_t)b * T_ * 4;
    T tn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t_begin * 4 + i];
    for (int t = t_begin; t < t_end; ++t) {
        T tf[3], tfn[3];
        tendon_force(P, tn, tf);
#pragma unroll
        for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)(t + 1) * 4 + i];   // t + 1 <= T - 1: always a valid row
        tendon_force(P, tn, tfn);
        T* nxt = out_b + (size_t)(t + 1) * tstride;
        T Gp[6], w[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            Gp[i] = G[i];
            G[i] = G[i] + ((G[i] - Gm1[i]) + (D[i] - Dm1[i]));   // linear extrapolation + anticipated tension change
            Dm1[i] = D[i];
            w[i] = T(0);
        }
        bool done = false;
        int status = 0, marches = 0;
        T sprev = T(0);
        Cest = T(0);      // the curvature estimate must come from THIS step's own iterations (inputs may jump between steps)
        HistViewE<T, NH, KC_WG, NC> H{Hs, N};
        // Gm: the point the group marches.  It is frozen once the rod is done, so while other rods of the warp still
        // iterate its re-marches reproduce the marched states bit for bit and the sink stores without a predicate.
        T Gm[6];
        bool first = true;
        while (true) {
            T eps[6], Ge[6], F[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) Gm[i] = done ? Gm[i] : G[i];
            wide_eps(Gm, fd_eps, eps);
#pragma unroll
            for (int i = 0; i < 6; ++i) Ge[i] = Gm[i] + ((k == i + 1) ? eps[i] : T(0));
            if (first) {
                // the first joint march of a step is never the accepted one (see the loop exit): no state stores; its
                // spare lane 7 marches the base point under the NEXT step's tendon load (tension-aware predictor)
                WideNoSink S0;
                T tfl[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) tfl[i] = (k == 7) ? tfn[i] : tf[i];
                rod_march<T, DIAG, IN, NH>(P, M, Ge, tfl, H, S0, F);
            } else {
                SmemStateSinkE<T, KC_WS, NC> S{Sg + (k < 7 ? k : 0), N};   // lane 7 repeats lane 0 (same values, same address)
                rod_march<T, DIAG, IN, NH>(P, M, Ge, tf, H, S, F);
            }
            T Fall[7][6];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Fall[c][i] = __shfl_sync(full, F[i], (lane & ~7) | c);
            }
            if (first) {
                T Fnext[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) Fnext[i] = __shfl_sync(full, F[i], lane | 7);
                ++marches;
                const int r = wide_decide_lin(Fall, G, eps, tol, Cest, sprev, w, Fnext, D);
                if (r != 0) { done = true; status = r; }
                else if (marches >= max_iter) { done = true; status = -1; }
            } else if (!done) {
                ++marches;
                const int r = wide_decide_lin(Fall, G, eps, tol, Cest, sprev, w);
                if (r != 0) { done = true; status = r; }
                else if (marches >= max_iter) { done = true; status = -1; }
            }
            // a rod that converged on the first march is re-marched once at its frozen point so that S holds its state
            if (!first && __all_sync(full, done)) break;
            first = false;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { Gm1[i] = Gp[i]; w[i] = (status == 2) ? w[i] : T(0); }
        __syncwarp();
        // accepted state = base state + first-order correction (w = 0 unless status == 2); lane k of the group handles
        // elements k, k+8, ... of the time slice: consecutive lanes write consecutive addresses of the reference layout.
        // Four elements per lane are loaded before any of them is stored so the shared-memory loads overlap.
        auto emit4 = [&](int e0) {
            T s0[4], sc[4][6], ap[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + 8 * u;
                const int ec = e < NV ? e : NV - 1;
                const T* sp = Sg + (size_t)ec * KC_WS;
                s0[u] = sp[0];
#pragma unroll
                for (int c = 0; c < 6; ++c) sc[u][c] = sp[c + 1];
                ap[u] = As[(size_t)(ec >= H0 ? ec - H0 : 0) * KC_WG];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + 8 * u;
                const T v0 = s0[u];
                T va = (sc[u][0] - v0) * w[0], vb = (sc[u][1] - v0) * w[1];
                va += (sc[u][2] - v0) * w[2]; vb += (sc[u][3] - v0) * w[3];
                va += (sc[u][4] - v0) * w[4]; vb += (sc[u][5] - v0) * w[5];
                const T v = v0 + (va + vb);
                if (e < NV) {
                    if (valid) nxt[e] = v;
                    if (e >= H0) {   // history for the next step; A <- accepted
                        const size_t hi = (size_t)(e - H0) * KC_WG;
                        Hs[hi] = P.c1 * v + P.c2 * ap[u];
                        As[hi] = v;
                    }
                }
            }
        };
        if (NC) {
#pragma unroll
            for (int i = 0; i < (25 * NC + 31) / 32; ++i) emit4(k + 32 * i);
        } else {
            for (int e0 = k; e0 < NV; e0 += 32) emit4(e0);
        }
        if (k == 0 && valid) {
            if (Gout) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Gout[((size_t)b * T_ + t + 1) * 6 + i] = G[i];
            }
            if (iters) iters[(size_t)b * T_ + t + 1] = status > 0 ? marches : -marches;
        }
        __syncwarp();
    }
    if (k == 0 && valid) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            gstate[(size_t)i * Bpad + b] = G[i]; gstate[(size_t)(6 + i) * Bpad + b] = Gm1[i];
            gstate[(size_t)(12 + i) * Bpad + b] = D[i]; gstate[(size_t)(18 + i) * Bpad + b] = Dm1[i];
        }
    }
}

// rows == 50 after a direct-layout rollout: rows 25:50 of time index t are yh,zh = c1*state[t-1] + c2*state[t-2]
// (state[-1] := state[0]); index 0 repeats [y;z] (knode.py:68,74-75).  One thread per element of [B][T][25*N].
template <typename T>
__global__ void kc_hist_rows_kernel(T* __restrict__ traj, int64_t B, int T_, int NV, int t0, int nt, T c1, T c2) {
    // time indices t0 .. t0+nt-1 of every rod
    const int64_t total = B * nt * NV;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % NV);
        const int64_t bt = i / NV;
        const int t = t0 + (int)(bt % nt);
        T* slice = traj + ((bt / nt) * T_ + t) * 2 * NV;
        T v;
        if (t == 0) v = slice[e];
        else {
            const T a = slice[e - 2 * NV];
            const T p = t >= 2 ? slice[e - 4 * NV] : a;
            v = c1 * a + c2 * p;
        }
        slice[NV + e] = v;
    }
}

// trajD[tile][T][N][25][32] -> traj[B][T][rows][N].  One CTA = one (32-rod tile, time index): the slab is contiguous,
// staged through a padded shared tile so both sides are coalesced.
// rows == 50 adds yh,zh = c1*state[t-1] + c2*state[t-2] (state[-1] := state[0]); index 0 repeats [y;z] (knode.py:68).
template <typename T>
__global__ void kc_traj_transpose_kernel(const T* __restrict__ trajD, T* __restrict__ traj, int64_t B, int T_, int N,
                                         int rows, int t_first, T c1, T c2) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    T* tile = reinterpret_cast<T*>(kc_smem);  // [N*25][33], k' = j*25 + r
    const int K = 25 * N;
    const int t = t_first + blockIdx.y;   // grid.y = number of time indices handled by this launch
    const int64_t b0 = (int64_t)blockIdx.x * 32;
    const int lane = threadIdx.x, wy = threadIdx.y, nwy = blockDim.y;
    const size_t slab = (size_t)K * KC_LS;
    const T* base = trajD + (size_t)blockIdx.x * T_ * slab;
    const int passes = rows == 50 ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
        for (int k = wy; k < K; k += nwy) {
            T v;
            if (pass == 0) {
                v = base[(size_t)t * slab + (size_t)k * KC_LS + lane];
            } else if (t == 0) {
                v = base[(size_t)k * KC_LS + lane];
            } else {
                const int tm1 = t - 1, tm2 = t >= 2 ? t - 2 : 0;
                v = c1 * base[(size_t)tm1 * slab + (size_t)k * KC_LS + lane] + c2 * base[(size_t)tm2 * slab + (size_t)k * KC_LS + lane];
            }
            tile[k * 33 + lane] = v;
        }
        __syncthreads();
        for (int rr = wy; rr < 32; rr += nwy) {
            if (b0 + rr < B) {
                T* dst = traj + (((size_t)(b0 + rr) * T_ + t) * rows + (size_t)pass * 25) * N;
                for (int kk = lane; kk < K; kk += 32) {  // kk = r*N + j in the reference layout
                    const int r = kk / N, j = kk - r * N;
                    dst[kk] = tile[(j * 25 + r) * 33 + rr];
                }
            }
        }
        __syncthreads();
    }
}

// W1[H][in], b1[H], W2[25][H] -> packed rows (see MlpC).
template <typename T>
__global__ void kc_pack_mlp_kernel(const T* __restrict__ W1, const T* __restrict__ b1, const T* __restrict__ W2, T* Wp,
                                   int in_dim, int inP, int hidden, int stride) {
    const int i = blockIdx.x;
    for (int k = threadIdx.x; k < stride; k += blockDim.x) {
        T v = T(0);
        if (k < in_dim) v = W1[(size_t)i * in_dim + k];
        else if (k == inP) v = b1[i];
        else if (k >= inP + 4 && k < inP + 4 + 25) v = W2[(size_t)(k - inP - 4) * hidden + i];
        Wp[(size_t)i * stride + k] = v;
    }
}

template <typename T>
int kc_pack_mlp(const kc_mlp* mlp, T* Wp, MlpC<T>& M, cudaStream_t st) {
    M.in_dim = mlp->in_dim;
    M.inP = (mlp->in_dim + 3) & ~3;
    M.hidden = mlp->hidden;
    M.stride = M.inP + 32;
    M.Wp = Wp;
    M.b2 = (const T*)mlp->b2;
    kc_pack_mlp_kernel<T><<<mlp->hidden, 64, 0, st>>>((const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Wp,
                                                       M.in_dim, M.inP, M.hidden, M.stride);
    KC_CHECK_LAUNCH("kc_pack_mlp");
    return KC_OK;
}
template int kc_pack_mlp<float>(const kc_mlp*, float*, MlpC<float>&, cudaStream_t);
template int kc_pack_mlp<double>(const kc_mlp*, double*, MlpC<double>&, cudaStream_t);

int kc_check_mlp(const kc_mlp* mlp) {
    if (!mlp) return KC_OK;
    KC_CHECK_ARG(mlp->in_dim == 28 || mlp->in_dim == 53, "kc_mlp.in_dim must be 28 or 53 (got %d)", mlp->in_dim);
    KC_CHECK_ARG(mlp->out_dim == 25, "kc_mlp.out_dim must be 25 (got %d)", mlp->out_dim);
    KC_CHECK_ARG(mlp->hidden >= 1, "kc_mlp.hidden must be >= 1");
    KC_CHECK_ARG(mlp->W1 && mlp->b1 && mlp->W2 && mlp->b2, "kc_mlp has a NULL weight pointer");
    return KC_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct RolloutWs {
    size_t trajD, wp, wc, state, tcimg, tcscr, total;
    size_t Bpad;
};
static RolloutWs rollout_ws(int dtype, int N, const kc_mlp* mlp, int64_t B, int64_t T_) {
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    RolloutWs w;
    w.Bpad = (size_t)((B + 31) / 32) * 32;
    w.trajD = 0;
    size_t off = align256((size_t)T_ * 25 * N * w.Bpad * sz);
    w.wp = off;
    if (mlp) off += align256((size_t)mlp->hidden * (((mlp->in_dim + 3) & ~3) + 32) * sz);
    w.wc = off;
    if (mlp) off += align256((size_t)kc_coop_row(mlp->in_dim) * (size_t)((mlp->hidden + 31) & ~31) * sz);
    w.state = off;
    off += align256((size_t)KC_SHOOT_SLOTS * w.Bpad * sz);
    w.tcimg = off;
    if (mlp) off += align256(kc_knode_tc_img_bytes());
    w.tcscr = off;
    if (mlp && dtype == KC_F32) off += align256(kc_knode_tc_fwd_scratch_bytes(N, B));
    w.total = off;
    return w;
}

extern "C" int64_t kc_rollout_workspace_bytes(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_) {
    if (!P || P->N < 2 || B < 0 || T_ < 1 || (dtype != KC_F32 && dtype != KC_F64)) return KC_EINVAL;
    return (int64_t)rollout_ws(dtype, P->N, mlp, B, T_).total;
}

template <typename T>
static int rollout_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t B, int64_t T_, const void* tensions,
                         const void* y0, const void* z0, double tol, int max_iter, int rows, void* traj, void* G_out,
                         int32_t* iters, void* workspace, int t_begin, int t_end, bool query_resumable, int method,
                         cudaStream_t st) {
    // steps [t_begin, t_end) of the rollout (full: 0 .. T-1).  query_resumable: launch nothing, return 1 if the mode this
    // call would select can be run in several time ranges (narrow and wide-lin keep their solver state), else 0.
    const RodC<T> P = make_rodc<T>(*Pp);
    const int N = P.N;
    const RolloutWs w = rollout_ws(sizeof(T) == 4 ? KC_F32 : KC_F64, N, mlp, B, T_);
    unsigned char* ws = (unsigned char*)workspace;
    T* trajD = (T*)(ws + w.trajD);
    T* state = (T*)(ws + w.state);
    MlpC<T> M{};
    int in_dim = 0;
    if (mlp) {
        if (!query_resumable) {
            int rc = kc_pack_mlp<T>(mlp, (T*)(ws + w.wp), M, st);
            if (rc) return rc;
        }
        in_dim = mlp->in_dim;
    }
    const T tl = tol > 0 ? (T)tol : (sizeof(T) == 4 ? T(2e-6) : T(1e-11));
    const T fd_eps = sizeof(T) == 4 ? T(1e-2) : T(1e-6);
    if (max_iter <= 0) max_iter = 60;
    const int NH = in_dim == 53 ? 25 : 12;
    const int threads = 32;
    const bool rk4 = method == KC_MARCH_RK4;   // RK4 spatial march: one rod per lane only (narrow kernel)
    const size_t smem = ((size_t)NH * (rk4 ? N : N - 1) + KC_SHOOT_SLOTS) * threads * sizeof(T);
    KC_CHECK_ARG(smem <= 227 * 1024, "N=%d too large for the shared-memory history (%zu B)", N, smem);
    // rods per warp: the kernel is latency bound, so with few rods spread them over the chip's 148 x 4 warp schedulers
    // (one warp each) before filling the lanes of a warp; never exceed one resident wave.
    int rpw = 32;
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t slots = (int64_t)sms * 4;
        while (rpw > 8 && (B + rpw / 2 - 1) / (rpw / 2) <= slots) rpw >>= 1;
        const char* e = getenv("KC_RPW");  // tuning knob for profiling runs
        if (e && atoi(e) >= 1 && atoi(e) <= 32) rpw = atoi(e);
    }
    const unsigned grid = (unsigned)((B + rpw - 1) / rpw);
    // mode: wide (8 lanes per rod, Newton with a fresh FD Jacobian per joint march) while the 8x lanes still fit the
    // chip's resident warps; narrow (one rod per lane, Broyden) beyond that.  KC_ROLLOUT_MODE=wide|narrow overrides.
    // Measured (tools/b_sweep.py, fp32 physics): wide-lin 1.25 ms @ 4096, 2.31 ms @ 8192, 3.51 ms @ 12288 rods against
    // 3.31 / 3.61 ms for narrow @ 8192 / 12288 -> wide up to three resident waves (7 warps x 4 rods per SM each).
    // Without the linearised variant (fp64, N > 10: shared memory) the older bound of 2 warps per scheduler stands.
    const bool lin_fits = rows != 0 &&
                          ((size_t)N * 25 * KC_WS + (size_t)2 * N * NH * KC_WG) * sizeof(T) <= 32 * 1024;
    bool wide = lin_fits ? B <= (int64_t)3 * 148 * 7 * KC_WG : B * 8 <= (int64_t)148 * 4 * 32 * 2;
    {
        const char* e = getenv("KC_ROLLOUT_MODE");
        if (e && e[0] == 'w') wide = true;
        if (e && e[0] == 'n') wide = false;
    }
    if (in_dim != 0) {   // with the MLP in the march every finite-difference lane would pay a full MLP: one rod per lane
        const char* e = getenv("KC_ROLLOUT_MODE");
        if (!(e && e[0] == 'w')) wide = false;
    }
    // linearised final correction (saves the last verification march) when its shared-memory state fits 7 warps per SM
    // (it writes the reference layout directly, so it needs rows != 0)
    const size_t lsmem = ((size_t)N * 25 * KC_WS + (size_t)2 * N * NH * KC_WG) * sizeof(T);
    bool lin = wide && rows != 0 && lsmem <= 32 * 1024;
    // larger states (fp64, N > 10): still worth it while all warps are resident at once (it saves about one joint march
    // in four); beyond one wave the plain wide kernel, whose footprint is a few KB, wins
    if (wide && rows != 0 && !lin && lsmem <= 200 * 1024) {
        const int64_t per_sm = (int64_t)(227 * 1024) / (int64_t)(lsmem + 1024);
        lin = (B + KC_WG - 1) / KC_WG <= 148 * per_sm;
    }
    {
        const char* e = getenv("KC_ROLLOUT_LIN");
        if (e && e[0] == '0') lin = false;
        if (e && e[0] == '1' && lsmem <= 200 * 1024) lin = wide && rows != 0;
    }
    if (rk4) { wide = false; lin = false; }
    // warp-cooperative KNODE rollout: MLP in the march and too few rods to fill the chip with one rod per thread
    bool coop = in_dim != 0 && !wide && B <= 8192 && !rk4;
    {
        const char* e = getenv("KC_ROLLOUT_COOP");
        if (e && e[0] == '0') coop = false;
        if (e && e[0] == '1' && in_dim != 0 && !rk4) coop = true;
    }
    // MLP of the march on the tensor cores (kc_knode_tc.cu): 16 rods x 8 shooting points per CTA.  Default for every KNODE
    // rollout of the eligible shape; KC_ROLLOUT_TC=0 falls back to the SIMT kernels above, an explicit KC_ROLLOUT_COOP /
    // KC_ROLLOUT_MODE request is honoured (the parity tests run every mode).
    bool tc = kc_knode_tc_eligible(sizeof(T) == 4 ? KC_F32 : KC_F64, mlp, N, method) && !getenv("KC_ROLLOUT_COOP") &&
              !getenv("KC_ROLLOUT_MODE");
    {
        const char* e = getenv("KC_ROLLOUT_TC");
        if (e && e[0] == '1' && kc_knode_tc_eligible(sizeof(T) == 4 ? KC_F32 : KC_F64, mlp, N, method)) tc = true;
    }
    if (tc) { coop = false; wide = false; lin = false; }
    const bool full_range = t_begin == 0 && t_end == (int)T_ - 1;
    const bool resumable = !tc && !coop && (!wide || lin);
    if (query_resumable) return resumable ? 1 : 0;
    KC_CHECK_ARG(full_range || resumable, "this rollout mode cannot be run in time ranges");
    const int t_first = t_begin == 0 ? 0 : t_begin + 1, n_idx = t_end - t_first + 1;   // time indices this call produces
    if (B > 0 && n_idx > 0) {
        if (tc) {
            if constexpr (std::is_same<T, float>::value) {
                int rc = kc_knode_tc_fwd(P, mlp, B, (int)T_, (const float*)tensions, (const float*)y0, (const float*)z0, trajD,
                                         tl, max_iter, fd_eps, (float*)G_out, iters, ws + w.tcimg, ws + w.tcscr, st);
                if (rc) return rc;
            }
        } else if (coop) {
        MlpCoop<T> MC;
        static_cast<MlpC<T>&>(MC) = M;
        MC.Hp = (mlp->hidden + 31) & ~31;
        T* Wc = (T*)(ws + w.wc);
        MC.Wc = Wc;
        const int inP = (in_dim + 3) & ~3;
        kc_pack_mlp_coop_kernel<T><<<64, 256, 0, st>>>((const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Wc, in_dim, inP,
                                                      mlp->hidden, MC.Hp);
        KC_CHECK_LAUNCH("kc_pack_mlp_coop_kernel");
#define KC_LAUNCH_COOP(D, I, H)                                                                                        \
    do {                                                                                                               \
        auto kern = kc_rollout_coop_kernel<T, D, I, H>;                                                                \
        const size_t state_b = (size_t)KC_COOP_WARPS * (H * (N - 1) + KC_SHOOT_SLOTS) * sizeof(T);                     \
        const size_t wc_b = (size_t)kc_coop_row(I) * MC.Hp * sizeof(T);                                                \
        const int wc_elems = wc_b + state_b <= 200 * 1024 ? (int)(wc_b / sizeof(T)) : 0;                               \
        const size_t csmem = state_b + (size_t)wc_elems * sizeof(T);                                                   \
        if (csmem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);    \
        kern<<<(unsigned)((B + KC_COOP_WARPS - 1) / KC_COOP_WARPS), 32 * KC_COOP_WARPS, csmem, st>>>(                  \
            P, MC, B, (int)T_, (const T*)tensions, (const T*)y0, (const T*)z0, trajD, tl, max_iter, fd_eps,            \
            (T*)G_out, iters, wc_elems);                                                                               \
    } while (0)
        if (P.diag) { if (in_dim == 28) KC_LAUNCH_COOP(true, 28, 12); else KC_LAUNCH_COOP(true, 53, 25); }
        else { if (in_dim == 28) KC_LAUNCH_COOP(false, 28, 12); else KC_LAUNCH_COOP(false, 53, 25); }
#undef KC_LAUNCH_COOP
        } else if (wide && lin) {
            // writes the reference layout directly; rows 25:50 (if asked for) are filled by an elementwise pass
            const unsigned wgrid = (unsigned)((B + KC_WG - 1) / KC_WG);
            const int64_t tstride = (int64_t)rows * N;
#define KC_LAUNCH_WLIN(D, I, H, NC)                                                                                    \
    do {                                                                                                               \
        auto kern = kc_rollout_wide_lin_kernel<T, D, I, H, NC>;                                                        \
        if (lsmem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem);    \
        kern<<<wgrid, 32, lsmem, st>>>(P, M, B, (int)T_, (const T*)tensions, (const T*)y0, (const T*)z0, (T*)traj,     \
                                       tstride, state, w.Bpad, t_begin, t_end, tl, max_iter, fd_eps, (T*)G_out, iters);\
    } while (0)
            if (P.diag) {
                if (in_dim == 0) { if (N == 10) KC_LAUNCH_WLIN(true, 0, 12, 10); else KC_LAUNCH_WLIN(true, 0, 12, 0); }
                else if (in_dim == 28) KC_LAUNCH_WLIN(true, 28, 12, 0);
                else KC_LAUNCH_WLIN(true, 53, 25, 0);
            } else {
                if (in_dim == 0) { if (N == 10) KC_LAUNCH_WLIN(false, 0, 12, 10); else KC_LAUNCH_WLIN(false, 0, 12, 0); }
                else if (in_dim == 28) KC_LAUNCH_WLIN(false, 28, 12, 0);
                else KC_LAUNCH_WLIN(false, 53, 25, 0);
            }
#undef KC_LAUNCH_WLIN
            KC_CHECK_LAUNCH("kc_rollout_wide_lin_kernel");
            if (rows == 50) {
                const int64_t total = B * n_idx * 25 * N;
                const unsigned hgrid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
                kc_hist_rows_kernel<T><<<hgrid, 256, 0, st>>>((T*)traj, B, (int)T_, 25 * N, t_first, n_idx, P.c1, P.c2);
                KC_CHECK_LAUNCH("kc_hist_rows_kernel");
            }
            return KC_OK;
        } else if (wide) {
            const size_t wsmem = (size_t)NH * (N - 1) * KC_WG * sizeof(T);
            const unsigned wgrid = (unsigned)((B + KC_WG - 1) / KC_WG);
#define KC_LAUNCH_WIDE(D, I, H)                                                                                        \
    do {                                                                                                               \
        auto kern = kc_rollout_wide_kernel<T, D, I, H>;                                                                \
        if (wsmem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);    \
        kern<<<wgrid, 32, wsmem, st>>>(P, M, B, (int)T_, (const T*)tensions, (const T*)y0, (const T*)z0, trajD, tl,    \
                                       max_iter, fd_eps, (T*)G_out, iters);                                            \
    } while (0)
            if (P.diag) {
                if (in_dim == 0) KC_LAUNCH_WIDE(true, 0, 12);
                else if (in_dim == 28) KC_LAUNCH_WIDE(true, 28, 12);
                else KC_LAUNCH_WIDE(true, 53, 25);
            } else {
                if (in_dim == 0) KC_LAUNCH_WIDE(false, 0, 12);
                else if (in_dim == 28) KC_LAUNCH_WIDE(false, 28, 12);
                else KC_LAUNCH_WIDE(false, 53, 25);
            }
#undef KC_LAUNCH_WIDE
        } else {
#define KC_LAUNCH_ROLL(D, I, H)                                                                                        \
    do {                                                                                                               \
        auto kern = rk4 ? kc_rollout_kernel<T, D, I, H, KC_MARCH_RK4> : kc_rollout_kernel<T, D, I, H, KC_MARCH_EULER>; \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        kern<<<grid, threads, smem, st>>>(P, M, B, (int)T_, rpw, (const T*)tensions, (const T*)y0, (const T*)z0,       \
                                          trajD, w.Bpad, state, t_begin, t_end, tl, max_iter, fd_eps, (T*)G_out,       \
                                          iters);                                                                      \
    } while (0)
            if (P.diag) {
                if (in_dim == 0) KC_LAUNCH_ROLL(true, 0, 12);
                else if (in_dim == 28) KC_LAUNCH_ROLL(true, 28, 12);
                else KC_LAUNCH_ROLL(true, 53, 25);
            } else {
                if (in_dim == 0) KC_LAUNCH_ROLL(false, 0, 12);
                else if (in_dim == 28) KC_LAUNCH_ROLL(false, 28, 12);
                else KC_LAUNCH_ROLL(false, 53, 25);
            }
#undef KC_LAUNCH_ROLL
        }
        KC_CHECK_LAUNCH("kc_rollout_kernel");
        if (rows == 0) return KC_OK;
        const int K = 25 * N;
        const size_t tsmem = (size_t)K * 33 * sizeof(T);
        auto tk = kc_traj_transpose_kernel<T>;
        if (tsmem > 48 * 1024) cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
        dim3 tgrid((unsigned)((B + 31) / 32), (unsigned)n_idx);
        tk<<<tgrid, dim3(32, 8), tsmem, st>>>(trajD, (T*)traj, B, (int)T_, N, rows, t_first, P.c1, P.c2);
        KC_CHECK_LAUNCH("kc_traj_transpose_kernel");
    }
    return KC_OK;
}

static int rollout_checked(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                           const void* tensions, const void* y0, const void* z0, double tol, int32_t max_iter,
                           int32_t rows, void* traj, void* G_out, int32_t* iters, void* workspace,
                           int64_t workspace_bytes, int64_t t_begin, int64_t t_end, bool query, void* stream,
                           int method = KC_MARCH_EULER) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(B >= 0 && T_ >= 1, "B must be >= 0 and T >= 1");
    KC_CHECK_ARG(rows == 25 || rows == 50 || rows == 0, "rows must be 25, 50 or 0");
    KC_CHECK_ARG(0 <= t_begin && t_begin <= t_end && t_end <= T_ - 1, "time range must satisfy 0 <= t_begin <= t_end <= T-1");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    if (!query) {
        KC_CHECK_ARG(B == 0 || (tensions && (traj || rows == 0) && workspace), "NULL tensions/traj/workspace");
        KC_CHECK_ARG((y0 == nullptr) == (z0 == nullptr), "y0 and z0 must be given together");
        const int64_t need = kc_rollout_workspace_bytes(dtype, P, mlp, B, T_);
        if (workspace_bytes < need) {
            kc_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
            return KC_ENOSPACE;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        return rollout_typed<float>(P, mlp, B, T_, tensions, y0, z0, tol, max_iter, rows, traj, G_out, iters, workspace,
                                    (int)t_begin, (int)t_end, query, method, st);
    return rollout_typed<double>(P, mlp, B, T_, tensions, y0, z0, tol, max_iter, rows, traj, G_out, iters, workspace,
                                 (int)t_begin, (int)t_end, query, method, st);
}

extern "C" int kc_rollout_fwd(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                              const void* tensions, const void* y0, const void* z0, double tol, int32_t max_iter,
                              int32_t rows, void* traj, void* G_out, int32_t* iters, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    return rollout_checked(dtype, P, mlp, B, T_, tensions, y0, z0, tol, max_iter, rows, traj, G_out, iters, workspace,
                           workspace_bytes, 0, T_ - 1, false, stream);
}

extern "C" int kc_rollout_fwd_rk4(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                                  const void* tensions, const void* y0, const void* z0, double tol, int32_t max_iter,
                                  int32_t rows, void* traj, void* G_out, int32_t* iters, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
    return rollout_checked(dtype, P, mlp, B, T_, tensions, y0, z0, tol, max_iter, rows, traj, G_out, iters, workspace,
                           workspace_bytes, 0, T_ - 1, false, stream, KC_MARCH_RK4);
}

extern "C" int kc_rollout_fwd_range(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                                    const void* tensions, const void* y0, const void* z0, double tol, int32_t max_iter,
                                    int32_t rows, void* traj, void* G_out, int32_t* iters, void* workspace,
                                    int64_t workspace_bytes, int64_t t_begin, int64_t t_end, void* stream) {
    return rollout_checked(dtype, P, mlp, B, T_, tensions, y0, z0, tol, max_iter, rows, traj, G_out, iters, workspace,
                           workspace_bytes, t_begin, t_end, false, stream);
}

extern "C" int kc_rollout_resumable(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                                    int32_t rows) {
    return rollout_checked(dtype, P, mlp, B, T_, nullptr, nullptr, nullptr, 0.0, 0, rows, nullptr, nullptr, nullptr,
                           nullptr, 0, 0, T_ - 1, true, nullptr);
}
