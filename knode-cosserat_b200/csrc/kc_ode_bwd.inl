// kc_ode_bwd.inl — reverse mode of kc_ode_fwd (included by kc_train.cu; shares its weight-gradient kernels).
// One sample per thread: recompute the physics forward (for the MLP input x), back-propagate through the MLP to its
// input with broadcast weights, then through the physics with the hand-written adjoint (kc_adjoint.cuh).  When
// parameter cotangents are wanted, the kernel also emits X[Q][XPG] and dO[Q][32] and the train-step kernels 3+4 reduce
// them to gW1, gb1, gW2, gb2.

// PH 0: everything in one kernel (SIMT MLP input VJP).  PH 1 / 2: the tensor-core split — PH 1 only emits X and dO (the
// samples the tcgen05 kernels consume), PH 2 takes the MLP input cotangent gx from GX[Q][32] (kc_train_tc_kernel<3>) and
// finishes with the physics adjoint.
template <typename T, bool DIAG, int IN, int PH = 0>
__global__ void __launch_bounds__(128)
kc_ode_bwd_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t Q, const T* __restrict__ y,
                  const T* __restrict__ yh, const T* __restrict__ zh, const T* __restrict__ tf,
                  const T* __restrict__ g_ys, const T* __restrict__ g_z, T* __restrict__ g_y, T* __restrict__ g_yh,
                  T* __restrict__ g_zh, T* __restrict__ g_tf, T* __restrict__ X, int XPG, T* __restrict__ dO,
                  const T* __restrict__ GX = nullptr) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    T ry[19], rh[25], rt[3], gys[19], gz[6];
#pragma unroll
    for (int k = 0; k < 19; ++k) { ry[k] = y[q * 19 + k]; rh[k] = yh[q * 19 + k]; gys[k] = g_ys[q * 19 + k]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) { rh[19 + k] = zh[q * 6 + k]; gz[k] = g_z[q * 6 + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) rt[k] = tf[q * 3 + k];
    T gy[19], gyh[19], gzh[6], gtf[3];
#pragma unroll
    for (int k = 0; k < 19; ++k) gyh[k] = T(0);
    T gy_nn[19], gtf_nn[3];
#pragma unroll
    for (int k = 0; k < 19; ++k) gy_nn[k] = T(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) gtf_nn[k] = T(0);
    T gzh_nn[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    if (IN > 0) {
        T ys0[19], z0[6];
        rod_ode<T, DIAG>(P, ry, rh + 13, rh + 16, rh + 19, rh + 22, rt, ys0, z0);
        constexpr int INX = IN > 0 ? IN : 28;
        T x[INX], go[25], gx[INX];
        if (IN == 28) {
#pragma unroll
            for (int i = 0; i < 19; ++i) x[i] = ry[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) x[19 + i] = z0[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) x[25 + i] = rt[i];
        } else {
#pragma unroll
            for (int i = 0; i < 19; ++i) { x[i] = ry[i]; x[19 + i] = rh[i]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { x[(38 + i) % INX] = z0[i]; x[(44 + i) % INX] = rh[19 + i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) x[(50 + i) % INX] = rt[i];
        }
#pragma unroll
        for (int i = 0; i < 19; ++i) go[i] = gys[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) go[19 + i] = gz[i];
        if (X && PH != 2) {
#pragma unroll
            for (int i = 0; i < INX; ++i) X[(size_t)q * XPG + i] = x[i];
            for (int i = INX; i < XPG; ++i) X[(size_t)q * XPG + i] = T(0);
#pragma unroll
            for (int i = 0; i < 25; ++i) dO[(size_t)q * 32 + i] = go[i];
#pragma unroll
            for (int i = 25; i < 32; ++i) dO[(size_t)q * 32 + i] = T(0);
        }
        if (PH == 1) return;
        if (PH == 2) {
#pragma unroll
            for (int i = 0; i < INX; ++i) gx[i] = GX[(size_t)q * 32 + i];
        } else {
            mlp_input_vjp<T, INX>(M, x, go, gx);
        }
        if (IN == 28) {
#pragma unroll
            for (int i = 0; i < 19; ++i) gy_nn[i] = gx[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) gz[i] += gx[19 + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) gtf_nn[i] = gx[25 + i];
        } else {
#pragma unroll
            for (int i = 0; i < 19; ++i) { gy_nn[i] = gx[i]; gyh[i] = gx[19 + i]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { gz[i] += gx[(38 + i) % INX]; gzh_nn[i] = gx[(44 + i) % INX]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) gtf_nn[i] = gx[(50 + i) % INX];
        }
    }
    T gqh[3], gwh[3], gvh[3], guh[3];
    rod_ode_vjp<T, DIAG>(P, ry, rh + 13, rh + 16, rh + 19, rh + 22, rt, gys, gz, gy, gqh, gwh, gvh, guh, gtf);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        gyh[13 + i] += gqh[i]; gyh[16 + i] += gwh[i];
        gzh[i] = gvh[i] + gzh_nn[i]; gzh[3 + i] = guh[i] + gzh_nn[3 + i];
        gtf[i] += gtf_nn[i];
    }
    if (g_y) {
#pragma unroll
        for (int k = 0; k < 19; ++k) g_y[q * 19 + k] = gy[k] + gy_nn[k];
    }
    if (g_yh) {
#pragma unroll
        for (int k = 0; k < 19; ++k) g_yh[q * 19 + k] = gyh[k];
    }
    if (g_zh) {
#pragma unroll
        for (int k = 0; k < 6; ++k) g_zh[q * 6 + k] = gzh[k];
    }
    if (g_tf) {
#pragma unroll
        for (int k = 0; k < 3; ++k) g_tf[q * 3 + k] = gtf[k];
    }
}

// Parameter gradients from (X[Q][32], dO[Q][32]) samples on the tensor cores (kc_train_tc_kernel<2>) when the shape is the
// reference's (fp32, 28 inputs, hidden <= 512) and there are enough samples to fill the chip; returns false to fall back
// to the SIMT kernel.  slices = number of partial-gradient slices written (< 0: launch error).
int kc_tc_launch_mode(int mode, const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                      const float* TGT, float* W1hl, float* W2c_, float* partial, int64_t NP, double* loss_part,
                      float* pred_out, const float* dOin, int grid, cudaStream_t st, int xpitch = 32, int dopitch = 32);
template <typename T>
static bool kc_param_grads_tc(const kc_mlp*, int64_t, const T*, const T*, T*, const TrainWs&, unsigned char*, cudaStream_t,
                              int&, int = 32, int = 32) { return false; }
template <>
bool kc_param_grads_tc<float>(const kc_mlp* mlp, int64_t Q, const float* X, const float* dO, float* part, const TrainWs& t,
                              unsigned char* ws, cudaStream_t st, int& slices, int xpitch, int dopitch) {
    if (mlp->in_dim != 28 || mlp->hidden > 512 || Q < 4096) return false;
    const char* e = getenv("KC_TRAIN_MODE");
    if (e && e[0] == 's') return false;
    float* tcw = (float*)(ws + t.tcw);
    slices = kc_train_tc_grid(Q);
    const int rc = kc_tc_launch_mode(2, mlp, 1.f, Q, 2, 1, X, nullptr, nullptr, tcw, tcw + 4 * 2 * 128 * 32, part, t.NP,
                                     (double*)(ws + t.lossp), nullptr, dO, slices, st, xpitch, dopitch);
    if (rc) slices = -1;
    return true;
}

// MLP input gradients on the tensor cores (kc_train_tc_kernel<3>): gx[Q][gxpitch] from x[Q][xpitch], dO[Q][dopitch]
template <typename T>
static int kc_input_grads_tc(const kc_mlp*, int64_t, const T*, int, const T*, int, T*, int, const TrainWs&, unsigned char*,
                             cudaStream_t) { return 1; }
template <>
int kc_input_grads_tc<float>(const kc_mlp* mlp, int64_t Q, const float* X, int xpitch, const float* dO, int dopitch,
                             float* gx, int gxpitch, const TrainWs& t, unsigned char* ws, cudaStream_t st) {
    float* tcw = (float*)(ws + t.tcw);
    return kc_tc_launch_mode(3, mlp, 1.f, Q, 2, gxpitch, X, nullptr, nullptr, tcw, tcw + 4 * 2 * 128 * 32, nullptr, 0,
                             (double*)(ws + t.lossp), gx, dO, kc_train_tc_grid(Q), st, xpitch, dopitch);
}
static bool kc_tc_bwd_ok(const kc_mlp* mlp, int64_t Q, bool is_float) {
    if (!is_float || !mlp || mlp->in_dim != 28 || mlp->hidden > 512 || Q < 4096) return false;
    const char* e = getenv("KC_TRAIN_MODE");
    return !(e && e[0] == 's');
}

struct OdeBwdWs { size_t X, dO, part, wp, total; TrainWs t; };
static OdeBwdWs ode_bwd_ws(int dtype, const kc_mlp* mlp, int64_t Q) {
    OdeBwdWs w{};
    if (!mlp) { w.total = 256; return w; }
    w.t = train_ws(dtype, mlp, Q);
    w.X = w.t.X; w.dO = w.t.dO; w.part = w.t.part; w.wp = w.t.wp; w.total = w.t.total;
    return w;
}

extern "C" int64_t kc_ode_bwd_workspace_bytes(int dtype, const kc_mlp* mlp, int64_t Q) {
    if ((dtype != KC_F32 && dtype != KC_F64) || Q < 0) return KC_EINVAL;
    if (mlp && kc_check_mlp(mlp)) return KC_EINVAL;
    return (int64_t)ode_bwd_ws(dtype, mlp, Q).total;
}

template <typename T>
static int ode_bwd_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t Q, const void* y, const void* yh,
                         const void* zh, const void* tf, const void* g_ys, const void* g_z, void* g_y, void* g_yh,
                         void* g_zh, void* g_tf, void* gW1, void* gb1, void* gW2, void* gb2, void* workspace,
                         cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    const OdeBwdWs w = ode_bwd_ws(sizeof(T) == 4 ? KC_F32 : KC_F64, mlp, Q);
    unsigned char* ws = (unsigned char*)workspace;
    MlpC<T> M{};
    const int in_dim = mlp ? mlp->in_dim : 0;
    const bool want_params = mlp && (gW1 || gb1 || gW2 || gb2);
    T* X = nullptr; T* dO = nullptr;
    const bool tc = kc_tc_bwd_ok(mlp, Q, sizeof(T) == 4);
    if (mlp) {
        if (!tc) {
            int rc = kc_pack_mlp<T>(mlp, (T*)(ws + w.wp), M, st);
            if (rc) return rc;
        }
        if (want_params || tc) { X = (T*)(ws + w.X); dO = (T*)(ws + w.dO); }
    }
    if (Q > 0 && tc) {
        // tensor-core split: samples -> (parameter gradients below) + input gradients -> physics adjoint
        const unsigned grid = (unsigned)((Q + 127) / 128);
        T* GX = (T*)(ws + w.t.PHYS);   // Q x 32 scratch: the PHYS | TGT regions of the training workspace are contiguous
#define BWD1(D) kc_ode_bwd_kernel<T, D, 28, 1><<<grid, 128, 0, st>>>(P, M, Q, (const T*)y, (const T*)yh, (const T*)zh, (const T*)tf, (const T*)g_ys, (const T*)g_z, nullptr, nullptr, nullptr, nullptr, X, w.t.XPG, dO)
        if (P.diag) BWD1(true); else BWD1(false);
#undef BWD1
        KC_CHECK_LAUNCH("kc_ode_bwd_kernel<1>");
        const int rc = kc_input_grads_tc<T>(mlp, Q, X, 32, dO, 32, GX, 32, w.t, ws, st);
        if (rc) return rc < 0 ? rc : KC_ECUDA;
#define BWD2(D) kc_ode_bwd_kernel<T, D, 28, 2><<<grid, 128, 0, st>>>(P, M, Q, (const T*)y, (const T*)yh, (const T*)zh, (const T*)tf, (const T*)g_ys, (const T*)g_z, (T*)g_y, (T*)g_yh, (T*)g_zh, (T*)g_tf, X, w.t.XPG, dO, GX)
        if (P.diag) BWD2(true); else BWD2(false);
#undef BWD2
        KC_CHECK_LAUNCH("kc_ode_bwd_kernel<2>");
    } else if (Q > 0) {
        const unsigned grid = (unsigned)((Q + 127) / 128);
        const int XPG = mlp ? w.t.XPG : 0;
#define BWD(D, I) kc_ode_bwd_kernel<T, D, I><<<grid, 128, 0, st>>>(P, M, Q, (const T*)y, (const T*)yh, (const T*)zh, (const T*)tf, (const T*)g_ys, (const T*)g_z, (T*)g_y, (T*)g_yh, (T*)g_zh, (T*)g_tf, X, XPG, dO)
        if (P.diag) { if (in_dim == 0) BWD(true, 0); else if (in_dim == 28) BWD(true, 28); else BWD(true, 53); }
        else { if (in_dim == 0) BWD(false, 0); else if (in_dim == 28) BWD(false, 28); else BWD(false, 53); }
#undef BWD
        KC_CHECK_LAUNCH("kc_ode_bwd_kernel");
    }
    if (want_params) {
        T* part = (T*)(ws + w.part);
        int tc_slices = 0;
        if (Q > 0 && kc_param_grads_tc<T>(mlp, Q, X, dO, part, w.t, ws, st, tc_slices)) {
            if (tc_slices < 0) return KC_ECUDA;
        } else if (Q > 0) {
            const size_t smem = bwd_smem_bytes(in_dim, sizeof(T));
            dim3 grid((unsigned)w.t.chunks, (unsigned)w.t.splits);
            if (in_dim == 28) {
                auto k = kc_train_bwd_kernel<T, 28>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.t.NP, w.t.tps);
            } else {
                auto k = kc_train_bwd_kernel<T, 53>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.t.NP, w.t.tps);
            }
            KC_CHECK_LAUNCH("kc_train_bwd_kernel");
        }
        kc_train_reduce_kernel<T><<<(unsigned)((w.t.NP + 63) / 64), 256, 0, st>>>(part, Q > 0 ? (tc_slices > 0 ? tc_slices : w.t.splits) : 0, w.t.NP, mlp->hidden, in_dim,
                                                                                  (T*)gW1, (T*)gb1, (T*)gW2, (T*)gb2, nullptr, 0, nullptr);
        KC_CHECK_LAUNCH("kc_train_reduce_kernel");
    }
    return KC_OK;
}

extern "C" int kc_ode_bwd(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t Q, const void* y, const void* yh,
                          const void* zh, const void* tf, const void* g_ys, const void* g_z, void* g_y, void* g_yh,
                          void* g_zh, void* g_tf, void* gW1, void* gb1, void* gW2, void* gb2, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(Q >= 0, "Q must be >= 0");
    KC_CHECK_ARG(Q == 0 || (y && yh && zh && tf && g_ys && g_z), "NULL data pointer");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    const int64_t need = kc_ode_bwd_workspace_bytes(dtype, mlp, Q);
    if (mlp && (workspace_bytes < need || !workspace)) {
        kc_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return KC_ENOSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        return ode_bwd_typed<float>(P, mlp, Q, y, yh, zh, tf, g_ys, g_z, g_y, g_yh, g_zh, g_tf, gW1, gb1, gW2, gb2, workspace, st);
    return ode_bwd_typed<double>(P, mlp, Q, y, yh, zh, tf, g_ys, g_z, g_y, g_yh, g_zh, g_tf, gW1, gb1, gW2, gb2, workspace, st);
}

// ---------------------------------------------------------------------------------------------------------------
// kc_mlp_fwd / kc_mlp_bwd — the MLP alone (forward(), get_nn_output())
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int IN>
__global__ void __launch_bounds__(128)
kc_mlp_fwd_kernel(const MlpC<T> M, int64_t Q, const T* __restrict__ x, T* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    T rx[IN], o[25];
#pragma unroll
    for (int k = 0; k < IN; ++k) rx[k] = x[q * IN + k];
    mlp_eval<T, IN>(M, rx, o);
#pragma unroll
    for (int c = 0; c < 25; ++c) out[q * 25 + c] = o[c];
}

template <typename T, int IN>
__global__ void __launch_bounds__(128)
kc_mlp_bwd_kernel(const MlpC<T> M, int64_t Q, const T* __restrict__ x, const T* __restrict__ g_out, T* __restrict__ g_x,
                  T* __restrict__ X, int XPG, T* __restrict__ dO) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    T rx[IN], go[25], gx[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) rx[k] = x[q * IN + k];
#pragma unroll
    for (int c = 0; c < 25; ++c) go[c] = g_out[q * 25 + c];
    if (X) {
#pragma unroll
        for (int i = 0; i < IN; ++i) X[(size_t)q * XPG + i] = rx[i];
        for (int i = IN; i < XPG; ++i) X[(size_t)q * XPG + i] = T(0);
#pragma unroll
        for (int i = 0; i < 25; ++i) dO[(size_t)q * 32 + i] = go[i];
#pragma unroll
        for (int i = 25; i < 32; ++i) dO[(size_t)q * 32 + i] = T(0);
    }
    if (g_x) {
        mlp_input_vjp<T, IN>(M, rx, go, gx);
#pragma unroll
        for (int k = 0; k < IN; ++k) g_x[q * IN + k] = gx[k];
    }
}

// Library-owned scratch for the tensor-core forward of the no-workspace entry points (grown on demand, never shrunk; the
// consumers run on the caller's stream after the producers).
void* kc_tc_scratch(size_t bytes) {
    static void* buf = nullptr;
    static size_t cap = 0;
    if (bytes > cap) {
        if (buf) cudaFree(buf);   // synchronises: safe w.r.t. earlier consumers
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { buf = nullptr; cap = 0; return nullptr; }
        cap = bytes;
    }
    return buf;
}
bool kc_tc_forward_ok(const kc_mlp* mlp, int64_t Q) {
    if (!mlp || mlp->in_dim != 28 || mlp->hidden > 512 || Q < 4096) return false;
    const char* e = getenv("KC_TRAIN_MODE");
    return !(e && e[0] == 's');
}
constexpr size_t KC_TCW_BYTES = (size_t)4 * (2 * 128 * 32 * 4 + 32768 + 16384);   // weight images of kc_tc_prep_weights_kernel

template <typename T>
static int mlp_fwd_tc(const kc_mlp*, int64_t, const void*, void*, cudaStream_t) { return 1; }
template <>
int mlp_fwd_tc<float>(const kc_mlp* mlp, int64_t Q, const void* x, void* out, cudaStream_t st) {
    if (!kc_tc_forward_ok(mlp, Q) || ((uintptr_t)x & 15) != 0) return 1;   // 1: not taken
    float* tcw = (float*)kc_tc_scratch(KC_TCW_BYTES + 256 * sizeof(double));
    if (!tcw) { kc_set_error("cudaMalloc of the tensor-core scratch failed"); return KC_ECUDA; }
    // forward-only mode: out[Q][25] = W2 ELU(W1 x + b1) + b2, x rows 28 floats apart
    return kc_tc_launch_mode(1, mlp, 1.f, Q, 2, 1, (const float*)x, nullptr, nullptr, tcw, tcw + 4 * 2 * 128 * 32, nullptr, 0,
                             (double*)((unsigned char*)tcw + KC_TCW_BYTES), (float*)out, nullptr, kc_train_tc_grid(Q), st, 28);
}

template <typename T>
static int mlp_fwd_typed(const kc_mlp* mlp, int64_t Q, const void* x, void* out, cudaStream_t st) {
    {
        const int rc = mlp_fwd_tc<T>(mlp, Q, x, out, st);
        if (rc <= 0) return rc;
    }
    // packed weights in a scratch buffer owned by the library (this entry point takes no workspace)
    static void* buf = nullptr;
    static size_t cap = 0;
    const size_t bytes = (size_t)mlp->hidden * (((mlp->in_dim + 3) & ~3) + 32) * sizeof(T);
    if (bytes > cap) {
        if (buf) cudaFree(buf);
        if (cudaMalloc(&buf, bytes) != cudaSuccess) { buf = nullptr; cap = 0; kc_set_error("cudaMalloc failed"); return KC_ECUDA; }
        cap = bytes;
    }
    MlpC<T> M;
    int rc = kc_pack_mlp<T>(mlp, (T*)buf, M, st);
    if (rc) return rc;
    if (Q == 0) return KC_OK;
    const unsigned grid = (unsigned)((Q + 127) / 128);
    if (mlp->in_dim == 28) kc_mlp_fwd_kernel<T, 28><<<grid, 128, 0, st>>>(M, Q, (const T*)x, (T*)out);
    else kc_mlp_fwd_kernel<T, 53><<<grid, 128, 0, st>>>(M, Q, (const T*)x, (T*)out);
    KC_CHECK_LAUNCH("kc_mlp_fwd_kernel");
    return KC_OK;
}

extern "C" int kc_mlp_fwd(int dtype, const kc_mlp* mlp, int64_t Q, const void* x, void* out, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(mlp, "kc_mlp is NULL");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    KC_CHECK_ARG(Q >= 0 && (Q == 0 || (x && out)), "bad Q or NULL data pointer");
    if (dtype == KC_F32) return mlp_fwd_typed<float>(mlp, Q, x, out, (cudaStream_t)stream);
    return mlp_fwd_typed<double>(mlp, Q, x, out, (cudaStream_t)stream);
}

template <typename T>
static int mlp_bwd_typed(const kc_mlp* mlp, int64_t Q, const void* x, const void* g_out, void* g_x, void* gW1, void* gb1,
                         void* gW2, void* gb2, void* workspace, cudaStream_t st) {
    const OdeBwdWs w = ode_bwd_ws(sizeof(T) == 4 ? KC_F32 : KC_F64, mlp, Q);
    unsigned char* ws = (unsigned char*)workspace;
    const bool want_params = gW1 || gb1 || gW2 || gb2;
    const int in_dim = mlp->in_dim;
    if (Q > 0 && kc_tc_bwd_ok(mlp, Q, sizeof(T) == 4) && ((uintptr_t)x & 15) == 0) {   // (16-byte loads of x rows)
        // tensor cores straight on the caller's arrays: x rows 28 apart, g_out rows 25 apart, g_x rows 28 apart
        if (g_x) {
            const int rc = kc_input_grads_tc<T>(mlp, Q, (const T*)x, 28, (const T*)g_out, 25, (T*)g_x, 28, w.t, ws, st);
            if (rc) return rc < 0 ? rc : KC_ECUDA;
        }
        if (want_params) {
            T* part = (T*)(ws + w.part);
            int slices = 0;
            kc_param_grads_tc<T>(mlp, Q, (const T*)x, (const T*)g_out, part, w.t, ws, st, slices, 28, 25);
            if (slices <= 0) return KC_ECUDA;
            kc_train_reduce_kernel<T><<<(unsigned)((w.t.NP + 63) / 64), 256, 0, st>>>(part, slices, w.t.NP, mlp->hidden, in_dim,
                                                                                      (T*)gW1, (T*)gb1, (T*)gW2, (T*)gb2, nullptr, 0, nullptr);
            KC_CHECK_LAUNCH("kc_train_reduce_kernel");
        }
        return KC_OK;
    }
    MlpC<T> M;
    int rc = kc_pack_mlp<T>(mlp, (T*)(ws + w.wp), M, st);
    if (rc) return rc;
    T* X = want_params ? (T*)(ws + w.X) : nullptr;
    T* dO = want_params ? (T*)(ws + w.dO) : nullptr;
    if (Q > 0) {
        const unsigned grid = (unsigned)((Q + 127) / 128);
        if (in_dim == 28) kc_mlp_bwd_kernel<T, 28><<<grid, 128, 0, st>>>(M, Q, (const T*)x, (const T*)g_out, (T*)g_x, X, w.t.XPG, dO);
        else kc_mlp_bwd_kernel<T, 53><<<grid, 128, 0, st>>>(M, Q, (const T*)x, (const T*)g_out, (T*)g_x, X, w.t.XPG, dO);
        KC_CHECK_LAUNCH("kc_mlp_bwd_kernel");
    }
    if (want_params) {
        T* part = (T*)(ws + w.part);
        int tc_slices = 0;
        if (Q > 0 && kc_param_grads_tc<T>(mlp, Q, X, dO, part, w.t, ws, st, tc_slices)) {
            if (tc_slices < 0) return KC_ECUDA;
        } else if (Q > 0) {
            const size_t smem = bwd_smem_bytes(in_dim, sizeof(T));
            dim3 grid((unsigned)w.t.chunks, (unsigned)w.t.splits);
            if (in_dim == 28) {
                auto k = kc_train_bwd_kernel<T, 28>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.t.NP, w.t.tps);
            } else {
                auto k = kc_train_bwd_kernel<T, 53>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.t.NP, w.t.tps);
            }
            KC_CHECK_LAUNCH("kc_train_bwd_kernel");
        }
        kc_train_reduce_kernel<T><<<(unsigned)((w.t.NP + 63) / 64), 256, 0, st>>>(part, Q > 0 ? (tc_slices > 0 ? tc_slices : w.t.splits) : 0, w.t.NP, mlp->hidden, in_dim,
                                                                                  (T*)gW1, (T*)gb1, (T*)gW2, (T*)gb2, nullptr, 0, nullptr);
        KC_CHECK_LAUNCH("kc_train_reduce_kernel");
    }
    return KC_OK;
}

extern "C" int kc_mlp_bwd(int dtype, const kc_mlp* mlp, int64_t Q, const void* x, const void* g_out, void* g_x, void* gW1,
                          void* gb1, void* gW2, void* gb2, void* workspace, int64_t workspace_bytes, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(mlp, "kc_mlp is NULL");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    KC_CHECK_ARG(Q >= 0 && (Q == 0 || (x && g_out)), "bad Q or NULL data pointer");
    const int64_t need = kc_ode_bwd_workspace_bytes(dtype, mlp, Q);
    if (workspace_bytes < need || !workspace) {
        kc_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return KC_ENOSPACE;
    }
    if (dtype == KC_F32) return mlp_bwd_typed<float>(mlp, Q, x, g_out, g_x, gW1, gb1, gW2, gb2, workspace, (cudaStream_t)stream);
    return mlp_bwd_typed<double>(mlp, Q, x, g_out, g_x, gW1, gb1, gW2, gb2, workspace, (cudaStream_t)stream);
}
