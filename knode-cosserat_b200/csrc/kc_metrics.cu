// kc_metrics.cu — the evaluation metrics of the training drivers on the GPU:
//   * DTW of the tip position of a predicted rollout against a reference rollout — physics_train.py:159,
//     physics_multitrain.py:211: fastdtw(trajectory[:, :3, 9], tip_pos)[0].  fastdtw (not installed here, third party,
//     version unpinned by the reference) approximates the exact dynamic-time-warping distance with the L1 point distance;
//     this is the EXACT distance, evaluated one anti-diagonal at a time (the cells of a diagonal depend only on the two
//     previous diagonals) by one CTA per (prediction, reference) pair, in fp64 and in the same order of operations as the
//     host function it replaces (_train.dtw_l1): bit-identical results.
//   * position + Euler-angle MSE x 1000 — physics_multitrain.py:213-222: squared position errors of every node and squared
//     differences of scipy's Rotation.from_quat(q, scalar_first=True).as_euler('zyx') angles, concatenated and averaged.
#include <cuda_runtime.h>
#include <math.h>
#include "kc_common.cuh"

// extrinsic z-y-x angles of a (w,x,y,z) quaternion as scipy returns them: R = Rx(c) Ry(b) Rz(a), out = [a, b, c]
__device__ __forceinline__ void euler_zyx(const double qin[4], double e[3]) {
    const double n = 1.0 / sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
    const double w = qin[0] * n, x = qin[1] * n, y = qin[2] * n, z = qin[3] * n;
    const double r00 = 1.0 - 2.0 * (y * y + z * z), r01 = 2.0 * (x * y - w * z), r02 = 2.0 * (x * z + w * y);
    const double r12 = 2.0 * (y * z - w * x), r22 = 1.0 - 2.0 * (x * x + y * y);
    e[0] = atan2(-r01, r00);
    e[1] = asin(fmin(1.0, fmax(-1.0, r02)));
    e[2] = atan2(-r12, r22);
}

template <typename T>
__global__ void __launch_bounds__(256)
kc_eval_metrics_kernel(int Ta, int Tb, int rows, int N, int node, const T* __restrict__ pred, const T* __restrict__ ref,
                       double* __restrict__ dtw_out, double* __restrict__ mse_out) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    __shared__ double red[8];
    double* a = reinterpret_cast<double*>(kc_smem);      // [Ta][3] tip positions of the prediction
    double* b = a + (size_t)3 * Ta;                       // [Tb][3] of the reference
    double* d0 = b + (size_t)3 * Tb;                      // three anti-diagonals, indexed by i = 0..Ta
    double* d1 = d0 + (Ta + 1);
    double* d2 = d1 + (Ta + 1);
    const int e = blockIdx.x, tid = threadIdx.x;
    const T* P = pred + (size_t)e * Ta * rows * N;
    const T* R = ref + (size_t)e * Tb * rows * N;
    for (int i = tid; i < 3 * Ta; i += blockDim.x) a[i] = (double)P[((size_t)(i / 3) * rows + (i % 3)) * N + node];
    for (int i = tid; i < 3 * Tb; i += blockDim.x) b[i] = (double)R[((size_t)(i / 3) * rows + (i % 3)) * N + node];
    const double INF = 1.0 / 0.0;
    for (int i = tid; i <= Ta; i += blockDim.x) { d0[i] = i == 0 ? 0.0 : INF; d1[i] = INF; d2[i] = INF; }   // d0: diagonal 0
    __syncthreads();
    // diagonal dg holds acc[i][j], i + j == dg (1-based cells, acc[0][0] = 0, the rest of row/column 0 = inf)
    double* pm2 = d0;   // diagonal dg-2
    double* pm1 = d1;   // diagonal dg-1 (diagonal 1 = acc[0][1], acc[1][0] = inf)
    double* cur = d2;
    for (int dg = 2; dg <= Ta + Tb; ++dg) {
        const int ilo = max(1, dg - Tb), ihi = min(Ta, dg - 1);
        for (int i = ilo + tid; i <= ihi; i += blockDim.x) {
            const int j = dg - i;
            const double* pa = a + 3 * (i - 1);
            const double* pb = b + 3 * (j - 1);
            const double cost = (fabs(pa[0] - pb[0]) + fabs(pa[1] - pb[1])) + fabs(pa[2] - pb[2]);
            // acc[i-1][j] = pm1[i-1], acc[i][j-1] = pm1[i], acc[i-1][j-1] = pm2[i-1]
            cur[i] = cost + fmin(fmin(pm1[i - 1], pm1[i]), pm2[i - 1]);
        }
        // cells of this diagonal that lie on row 0 / column 0 are infinite
        if (tid == 0) { if (ilo == 1) cur[0] = INF; }
        if (tid == 1 && ihi + 1 <= Ta) cur[ihi + 1] = INF;
        __syncthreads();
        double* t = pm2; pm2 = pm1; pm1 = cur; cur = t;
    }
    if (tid == 0) dtw_out[e] = pm1[Ta];
    // position + Euler MSE (needs equally long trajectories)
    if (mse_out) {
        double acc = 0.0;
        if (Ta == Tb) {
            for (int i = tid; i < Ta * N; i += blockDim.x) {
                const int t = i / N, j = i - t * N;
                const T* p = P + (size_t)t * rows * N + j;
                const T* r = R + (size_t)t * rows * N + j;
                double qp[4], qr[4], ep[3], er[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { const double dlt = (double)p[c * N] - (double)r[c * N]; acc += dlt * dlt; }
#pragma unroll
                for (int c = 0; c < 4; ++c) { qp[c] = (double)p[(3 + c) * N]; qr[c] = (double)r[(3 + c) * N]; }
                euler_zyx(qp, ep);
                euler_zyx(qr, er);
#pragma unroll
                for (int c = 0; c < 3; ++c) { const double dlt = ep[c] - er[c]; acc += dlt * dlt; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            mse_out[e] = Ta == Tb ? s / (6.0 * Ta * N) * 1000.0 : nan("");
        }
    }
}

extern "C" int kc_eval_metrics(int dtype, int64_t E, int64_t Ta, int64_t Tb, int32_t rows, int32_t N, int32_t node,
                               const void* pred, const void* ref, double* dtw, double* mse, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(E >= 0 && Ta >= 1 && Tb >= 1 && N >= 1 && rows >= 7 && node >= 0 && node < N, "bad shape");
    KC_CHECK_ARG(dtw && (E == 0 || (pred && ref)), "NULL pointer");
    const size_t smem = ((size_t)3 * (Ta + Tb) + 3 * (Ta + 1)) * sizeof(double);
    KC_CHECK_ARG(smem <= 200 * 1024, "trajectories too long for the shared-memory wavefront (%zu B)", smem);
    if (E == 0) return KC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32) {
        auto k = kc_eval_metrics_kernel<float>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(unsigned)E, 256, smem, st>>>((int)Ta, (int)Tb, rows, N, node, (const float*)pred, (const float*)ref, dtw, mse);
    } else {
        auto k = kc_eval_metrics_kernel<double>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<(unsigned)E, 256, smem, st>>>((int)Ta, (int)Tb, rows, N, node, (const double*)pred, (const double*)ref, dtw, mse);
    }
    KC_CHECK_LAUNCH("kc_eval_metrics_kernel");
    return KC_OK;
}
