// kc_host.cu — knode.simulate for HOST callers (knode.py:55-102): tensions in host memory in, trajectory in host memory
// out, the device work and the PCIe transfers pipelined inside the call.
//
// A rod-node-step costs ~1 ns of compute on the GPU but leaves 100 bytes of trajectory to bring back: end to end the path
// is bound by the device-to-host copy.  The rollout is therefore run in a few TIME RANGES (kc_rollout_fwd_range — the
// narrow and wide-lin kernels resume from the trajectory plus a few words of solver state) and every finished range is
// copied back on a second stream while the next one is being solved.
#include <cuda_runtime.h>
#include "kc_common.cuh"

#define KC_CHECK_CUDA(call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            kc_set_error("%s: %s", #call, cudaGetErrorString(e__));                          \
            return KC_ECUDA;                                                                 \
        }                                                                                    \
    } while (0)

static inline size_t h_align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HostBuf {
    size_t tens, traj, G, iters, ws, total;
};
static HostBuf host_buf(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_, int rows) {
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    HostBuf h;
    size_t off = 0;
    h.tens = off; off += h_align256((size_t)B * T_ * 4 * sz);
    h.traj = off; off += h_align256((size_t)B * T_ * rows * P->N * sz);
    h.G = off; off += h_align256((size_t)B * T_ * 6 * sz);
    h.iters = off; off += h_align256((size_t)B * T_ * 4);
    h.ws = off; off += (size_t)kc_rollout_workspace_bytes(dtype, P, mlp, B, T_);
    h.total = off;
    return h;
}

extern "C" int64_t kc_rollout_host_device_bytes(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B,
                                                int64_t T_, int32_t rows) {
    if (!P || P->N < 2 || B < 0 || T_ < 1 || (dtype != KC_F32 && dtype != KC_F64) || (rows != 25 && rows != 50))
        return KC_EINVAL;
    return (int64_t)host_buf(dtype, P, mlp, B, T_, rows).total;
}

// one copy stream per device, created on first use (never destroyed: lives as long as the library)
static cudaStream_t g_copy_stream[64] = {};

extern "C" int kc_rollout_host(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                               const void* tensions_host, const void* y0, const void* z0, double tol, int32_t max_iter,
                               int32_t rows, void* traj_host, void* G_host, int32_t* iters_host, void* device_buf,
                               int64_t device_buf_bytes, int32_t segments, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(B >= 0 && T_ >= 1, "B must be >= 0 and T >= 1");
    KC_CHECK_ARG(rows == 25 || rows == 50, "rows must be 25 or 50");
    if (B == 0) return KC_OK;
    KC_CHECK_ARG(tensions_host && traj_host && device_buf, "NULL tensions_host/traj_host/device_buf");
    const HostBuf h = host_buf(dtype, P, mlp, B, T_, rows);
    if (device_buf_bytes < (int64_t)h.total) {
        kc_set_error("device buffer too small: %lld < %lld", (long long)device_buf_bytes, (long long)h.total);
        return KC_ENOSPACE;
    }
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    unsigned char* d = (unsigned char*)device_buf;
    void* tens_d = d + h.tens;
    unsigned char* traj_d = d + h.traj;
    void* G_d = G_host ? d + h.G : nullptr;
    int32_t* iters_d = iters_host ? (int32_t*)(d + h.iters) : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    KC_CHECK_CUDA(cudaGetDevice(&dev));
    KC_CHECK_ARG(dev >= 0 && dev < 64, "device index out of range");
    if (!g_copy_stream[dev]) KC_CHECK_CUDA(cudaStreamCreateWithFlags(&g_copy_stream[dev], cudaStreamNonBlocking));
    cudaStream_t cs = g_copy_stream[dev];

    // number of time ranges: enough that the first copy starts early, few enough that a range still fills the chip for
    // a while (each range is one launch) — about 64 MB of trajectory per range, at most 8
    const size_t slice = (size_t)rows * P->N * sz;                  // one time index of one rod
    const size_t total_bytes = (size_t)B * T_ * slice;
    int nseg = segments > 0 ? segments : (int)(total_bytes / ((size_t)64 << 20));
    if (nseg > 8) nseg = 8;
    if (nseg > T_ - 1) nseg = (int)(T_ - 1);
    if (nseg < 1) nseg = 1;
    if (nseg > 1) {
        const int r = kc_rollout_resumable(dtype, P, mlp, B, T_, rows);
        if (r < 0) return r;
        if (r == 0) nseg = 1;
    }
    KC_CHECK_CUDA(cudaMemcpyAsync(tens_d, tensions_host, (size_t)B * T_ * 4 * sz, cudaMemcpyHostToDevice, st));
    cudaEvent_t ev[8] = {};
    int rc = KC_OK;
    for (int s = 0; s < nseg && rc == KC_OK; ++s) {
        // steps [t0, t1): the copy engine is ~6x slower than the solve, so only the FIRST range's solve is exposed — it is
        // kept short (a quarter of an even share); the other ranges share the rest evenly
        const int64_t first = nseg > 1 ? ((T_ - 1) / (4 * nseg) > 0 ? (T_ - 1) / (4 * nseg) : 1) : T_ - 1;
        const int64_t t0 = s == 0 ? 0 : first + (T_ - 1 - first) * (s - 1) / (nseg - 1);
        const int64_t t1 = s == 0 ? first : first + (T_ - 1 - first) * s / (nseg - 1);
        rc = kc_rollout_fwd_range(dtype, P, mlp, B, T_, tens_d, y0, z0, tol, max_iter, rows, traj_d, G_d, iters_d,
                                  d + h.ws, (int64_t)(h.total - h.ws), t0, t1, stream);
        if (rc != KC_OK) break;
        cudaError_t e = cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(ev[s], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ev[s], 0);
        const int64_t i0 = s == 0 ? 0 : t0 + 1, ni = t1 - i0 + 1;             // time indices this range produced
        if (e == cudaSuccess && ni > 0) {
            if (nseg == 1)
                e = cudaMemcpyAsync(traj_host, traj_d, total_bytes, cudaMemcpyDeviceToHost, cs);
            else
                e = cudaMemcpy2DAsync((unsigned char*)traj_host + (size_t)i0 * slice, (size_t)T_ * slice,
                                      traj_d + (size_t)i0 * slice, (size_t)T_ * slice, (size_t)ni * slice, (size_t)B,
                                      cudaMemcpyDeviceToHost, cs);
        }
        if (e != cudaSuccess) {
            kc_set_error("kc_rollout_host: %s", cudaGetErrorString(e));
            rc = KC_ECUDA;
        }
    }
    if (rc == KC_OK) {
        cudaError_t e = cudaSuccess;
        if (G_host) e = cudaMemcpyAsync(G_host, G_d, (size_t)B * T_ * 6 * sz, cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && iters_host)
            e = cudaMemcpyAsync(iters_host, iters_d, (size_t)B * T_ * 4, cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) {
            kc_set_error("kc_rollout_host: %s", cudaGetErrorString(e));
            rc = KC_ECUDA;
        }
    }
    // a host API: the result is in host memory when the call returns
    cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(st);
    for (int s = 0; s < 8; ++s)
        if (ev[s]) cudaEventDestroy(ev[s]);
    if (rc == KC_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        kc_set_error("kc_rollout_host: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = KC_ECUDA;
    }
    return rc;
}
