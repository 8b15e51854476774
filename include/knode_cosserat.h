/* knode_cosserat.h — C ABI of libknode_cosserat_b200.so (hand-written sm_100a kernels).
 *
 * This is the drop-in boundary for the one hot path of hsiehScalAR/KNODE-Cosserat: batched Cosserat-rod ODE
 * evaluations, shooting marches, time rollouts and the KNODE (physics + MLP residual) teacher-forced training
 * step.  The reference has no FFI of its own — its boundary is the Python surface of
 * knode_cosserat/cosserat_ode_torch.py, cosserat_ode.py and knode.py — so every entry point below names the
 * reference function (file:line, relative to the reference repo root) whose arithmetic it replaces.  The
 * Python classes in knode-cosserat_b200/ bind these with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless its name ends in _host;
 *   - `dtype` selects the arithmetic type of every data pointer of the call: KC_F32 or KC_F64;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous and
 *     never allocate: scratch is caller-provided (`workspace`, sized by the matching *_workspace_bytes call);
 *   - return value 0 = launched; <0 = error (KC_E*), text via kc_last_error() (thread-local);
 *   - the entry points that mirror reference methods taking nothing but tensors (kc_ode_fwd, kc_mlp_fwd, kc_march,
 *     kc_segment_fwd) have no workspace argument: they keep a library-owned device scratch (packed weights, tensor-core
 *     operand images) that is grown on demand and never shrunk — call them from one host thread at a time, as the
 *     reference does (physics_train.py:179 pins torch to one thread), and not while a CUDA graph is being captured;
 *   - state layout (cosserat_ode_torch.py:141-152): y[19] = p(0:3) h(3:7,wxyz) n(7:10) m(10:13) q(13:16)
 *     w(16:19); z[6] = v(0:3) u(3:6); a rod is [25][N] row-major (rows = y then z, columns = nodes).
 */
#ifndef KNODE_COSSERAT_H
#define KNODE_COSSERAT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KC_F32 0
#define KC_F64 1

#define KC_OK 0
#define KC_EINVAL (-1)   /* bad argument (NULL pointer, size, dtype, N < 2, unsupported MLP shape) */
#define KC_ECUDA (-2)    /* CUDA runtime error at launch */
#define KC_ENOSPACE (-3) /* workspace too small */

#define KC_MARCH_EULER 0
#define KC_MARCH_RK4 1

/* Derived rod constants, exactly the attributes CosseratRodTorch.compute_intermediate_terms() produces
 * (cosserat_ode_torch.py:108-129 == cosserat_ode.py:58-78), plus boundary conditions and tendon geometry
 * (cosserat_ode_torch.py:25-45).  Always double; converted once per call to the arithmetic type. Matrices are
 * row-major 3x3.  The host side refreshes this struct from the Python attributes on every call because callers
 * mutate them (knode.setup_robot, knode.py:11-53). */
typedef struct kc_rod_params {
    int32_t N;        /* number of nodes (robot.N) */
    int32_t reserved; /* must be 0 */
    double ds, c0, c1, c2, rhoA;
    double Kse_c0Bse_inv[9]; /* (Kse + c0*Bse)^-1 */
    double Kbt_c0Bbt_inv[9]; /* (Kbt + c0*Bbt)^-1 */
    double Bse[9], Bbt[9], rhoJ[9];
    double Kse_vstar[3], rhoAg[3], C[3], F_tip[3], M_tip[3];
    double p0[3], h0[4], q0[3], w0[3];
    double tendon_dirs[12]; /* [4][3]; tendon_force = tensions[4] @ tendon_dirs (cosserat_ode_torch.py:334) */
} kc_rod_params;

/* The KNODE residual MLP: Linear(in_dim,hidden) -> ELU -> Linear(hidden,25) (cosserat_ode_torch.py:60-62).
 * Weights in torch's nn.Linear layout, dtype of the call.  in_dim 28 = [y;z;tf], 53 = [y;yh;z;zh;tf]
 * (nn_input_history, :194-197).  Pass a NULL kc_mlp* for physics only (use_nn == False). */
typedef struct kc_mlp {
    int32_t in_dim, hidden, out_dim, reserved;
    const void *W1, *b1, *W2, *b2; /* W1[hidden][in_dim], b1[hidden], W2[25][hidden], b2[25] */
} kc_mlp;

int kc_version(void);
const char *kc_last_error(void);

/* CosseratRodTorch.ODE_parallel (cosserat_ode_torch.py:217-322), CosseratRodTorch.ODE (:137-214) and
 * CosseratRod.ODE (cosserat_ode.py:114-186): Q independent node evaluations.
 * y[Q][19], yh[Q][19], zh[Q][6], tf[Q][3] -> ys[Q][19], z[Q][6]. */
int kc_ode_fwd(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t Q, const void *y, const void *yh,
               const void *zh, const void *tf, void *ys, void *z, void *stream);

/* Reverse mode of kc_ode_fwd (what torch.autograd does through ODE_parallel): cotangents g_ys[Q][19], g_z[Q][6]
 * -> g_y[Q][19], g_yh[Q][19], g_zh[Q][6], g_tf[Q][3] (each may be NULL = not wanted) and, when mlp != NULL,
 * parameter cotangents gW1[hidden][in], gb1[hidden], gW2[25][hidden], gb2[25] which are OVERWRITTEN (may be
 * NULL).  workspace: kc_ode_bwd_workspace_bytes(). */
int64_t kc_ode_bwd_workspace_bytes(int dtype, const kc_mlp *mlp, int64_t Q);
int kc_ode_bwd(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t Q, const void *y, const void *yh,
               const void *zh, const void *tf, const void *g_ys, const void *g_z, void *g_y, void *g_yh,
               void *g_zh, void *g_tf, void *gW1, void *gb1, void *gW2, void *gb2, void *workspace,
               int64_t workspace_bytes, void *stream);

/* The KNODE MLP alone: CosseratRodTorch.forward (cosserat_ode_torch.py:131-134) and CosseratRod.get_nn_output
 * (cosserat_ode.py:90-112): x[Q][in_dim] -> out[Q][25]; kc_mlp_bwd is its reverse mode (g_out[Q][25] -> g_x[Q][in_dim]
 * and/or OVERWRITTEN parameter cotangents; any of them may be NULL).  workspace: kc_ode_bwd_workspace_bytes(). */
int kc_mlp_fwd(int dtype, const kc_mlp *mlp, int64_t Q, const void *x, void *out, void *stream);
int kc_mlp_bwd(int dtype, const kc_mlp *mlp, int64_t Q, const void *x, const void *g_out, void *g_x, void *gW1,
               void *gb1, void *gW2, void *gb2, void *workspace, int64_t workspace_bytes, void *stream);

/* Shooting residual + spatial march for B rods: CosseratRod.getResidualEuler (cosserat_ode.py:188-213),
 * CosseratRod.getResidualRK4 (:215-255), CosseratRodTorch.getResidualEuler (cosserat_ode_torch.py:325-367).
 * G[B][6], y[B][19][N] and z[B][6][N] are updated IN PLACE like the numpy reference does (column 0 of y is
 * replaced by [p0,h0,G,q0,w0]; z[:, N-1] is left untouched), yh[B][19][N], zh[B][6][N], tensions[B][4]
 * -> res[B][6] = [F_tip - n(L), M_tip - m(L)]. */
int kc_march(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int method, int64_t B, const void *G, void *y,
             void *z, const void *yh, const void *zh, const void *tensions, void *res, void *stream);

/* Teacher-forced one-step prediction: CosseratRodTorch.parallelGetNextSegmentEuler (cosserat_ode_torch.py:401-437)
 * when K > 0 (key_idx_host[K] = node indices k, the ODE is evaluated at node k-1), and
 * CosseratRodTorch.getNextSegmentEuler (:370-399) when K == 0 (all nodes; out column 0 = Gs column 0).
 * Gs[S][25][N], yh[S][19][N], zh[S][6][N], tensions[S][4] -> out[S][25][K] (K>0) or out[S][25][N] (K==0). */
int kc_segment_fwd(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t S, int32_t K,
                   const int32_t *key_idx_host, const void *Gs, const void *yh, const void *zh,
                   const void *tensions, void *out, void *stream);

/* Time rollout of B independent rods: knode.simulate (knode.py:55-102) = BDF2 history (:74-77) + shooting solve
 * of the 6 base reactions (:88-89, here quasi-Newton instead of MINPACK hybrd, converged to `tol`) + Euler march.
 * tensions[B][T][4]; y0[B][19][N], z0[B][6][N] initial state or NULL for the straight rod of knode.py:58-64.
 * traj[B][T][rows][N], rows = 25 ([y;z]) or 50 ([y;z;yh;zh], the reference's layout); index 0 is the initial
 * state and the step driven by tensions[:, T-1] is not computed (knode.py:102 drops it).
 * G_out[B][T][6] (may be NULL): converged base reactions; iters[B][T] int32 (may be NULL): marches used, negative
 * if the solve did not reach tol in max_iter marches.
 * rows == 0 skips the reference-layout output (traj may be NULL; the device-layout trajectory stays in the
 * workspace) — used by bench.py to time the rollout kernel alone.
 * tol <= 0 selects the default (1e-11 fp64, 2e-6 fp32; relative to max(1,|G|)). */
int64_t kc_rollout_workspace_bytes(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T);
int kc_rollout_fwd(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T,
                   const void *tensions, const void *y0, const void *z0, double tol, int32_t max_iter,
                   int32_t rows, void *traj, void *G_out, int32_t *iters, void *workspace,
                   int64_t workspace_bytes, void *stream);

/* The same rollout with the reference's RK4 spatial march (CosseratRod.getResidualRK4, cosserat_ode.py:215-255, with the
 * mid-point histories of knode.py:80-81) in place of the explicit-Euler march inside the shooting solve — what
 * knode.simulate computes when its residual is getResidualRK4.  Same arguments, workspace and outputs
 * as kc_rollout_fwd; one rod per thread (no wide / warp-cooperative mode). */
int kc_rollout_fwd_rk4(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T,
                       const void *tensions, const void *y0, const void *z0, double tol, int32_t max_iter,
                       int32_t rows, void *traj, void *G_out, int32_t *iters, void *workspace,
                       int64_t workspace_bytes, void *stream);

/* The same rollout in TIME RANGES: solve steps t in [t_begin, t_end) only, i.e. write time indices t_begin+1 .. t_end
 * (and index 0 when t_begin == 0) of traj / G_out / iters.  Ranges must be issued in order on one stream with the same
 * buffers and workspace: a later range resumes from the trajectory already written plus the solver state kept in the
 * workspace.  Used to overlap the device-to-host copy of finished ranges with the solve of the next one.
 * kc_rollout_resumable() -> 1 if the kernel this (dtype, MLP, B, T, rows) selects can run in ranges (one-rod-per-lane
 * Broyden and wide-lin do), 0 if only the full range [0, T-1) is accepted, <0 on bad arguments. */
int kc_rollout_fwd_range(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T,
                         const void *tensions, const void *y0, const void *z0, double tol, int32_t max_iter,
                         int32_t rows, void *traj, void *G_out, int32_t *iters, void *workspace,
                         int64_t workspace_bytes, int64_t t_begin, int64_t t_end, void *stream);
int kc_rollout_resumable(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T, int32_t rows);

/* knode.simulate for HOST callers (knode.py:55-102, batched): tensions_host[B][T][4] in host memory ->
 * traj_host[B][T][rows][N] (+ G_host[B][T][6], iters_host[B][T] if non-NULL) in host memory (pinned memory makes the
 * copies asynchronous; pageable memory works, slower).  The call stages the tensions, runs the rollout in `segments`
 * time ranges (<= 0: chosen from the trajectory size, 1..8) and copies every finished range back on an internal second
 * stream while the next one is solved; it RETURNS WHEN THE RESULT IS IN HOST MEMORY (it synchronises `stream`).
 * rows = 25 | 50.  y0/z0 (optional) and the MLP weights are device pointers as everywhere else.  device_buf: device
 * scratch of kc_rollout_host_device_bytes() bytes (holds the device copies of tensions/traj/G/iters + the workspace). */
int64_t kc_rollout_host_device_bytes(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T,
                                     int32_t rows);
int kc_rollout_host(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T,
                    const void *tensions_host, const void *y0, const void *z0, double tol, int32_t max_iter,
                    int32_t rows, void *traj_host, void *G_host, int32_t *iters_host, void *device_buf,
                    int64_t device_buf_bytes, int32_t segments, void *stream);

/* Reverse mode THROUGH kc_rollout_fwd (back-propagation through time; BASELINE.json north_star subsystem 3 — an extension:
 * the reference never differentiates a rollout).  traj[B][T][25][N] is the forward result (rows = 25),
 * g_traj[B][T][25][N] = dL/dtraj -> g_tensions[B][T][4] (may be NULL) and, with an MLP, the OVERWRITTEN parameter
 * cotangents gW1, gb1, gW2, gb2 (may be NULL).  The shooting solve is differentiated by the implicit function theorem. */
int64_t kc_rollout_bwd_workspace_bytes(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T);
int kc_rollout_bwd(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T, const void *tensions,
                   const void *traj, const void *g_traj, void *g_tensions, void *gW1, void *gb1, void *gW2, void *gb2,
                   void *workspace, int64_t workspace_bytes, void *stream);

/* The training loss of physics_train.py:345-352 (p | n,m,q,w | euler(h) at the key nodes k, z at node k-1; every block a
 * mean over its entries, the K key nodes and the T-1 steps, Utils/transformations.py:3-31 for the Euler angles) applied to a
 * ROLLOUT against a target trajectory, fused with its cotangent — the loss of the north-star rollout training step (C3 ii;
 * the reference only ever applies it to teacher-forced one-step predictions).  traj, target: [B][T][25][N];
 * key_idx_host[K]: distinct node indices in [1, N-1] -> *loss (device, float64 always) = scale * sum over trajectories, and
 * g_traj[B][T][25][N] = d loss / d traj (OVERWRITTEN, dense), ready for kc_rollout_bwd. */
int kc_rollout_loss(int dtype, int64_t B, int64_t T, int32_t N, int32_t K, const int32_t *key_idx_host, const void *traj,
                    const void *target, double scale, double *loss, void *g_traj, void *stream);

/* One teacher-forced training step of physics_train.py's fast path (:313-368) == slow path (:215-267) restricted
 * to its key nodes == train_segment.py:140-185: for every trajectory b, step t in [0, T-2] and key node k:
 * ODE+MLP at node k-1 of the NEXT ground-truth state, Euler step, 4-term MSE loss (p | n,m,q,w | euler(h) | z,
 * physics_train.py:345-352, Utils/transformations.py:3-31), summed and divided by (T-1); reverse mode to the four
 * MLP tensors.  traj[B][T][25][N], controls[B][T][4], key_idx_host[K]
 * -> loss[1] (float64 always), gW1[hidden][in], gb1[hidden], gW2[25][hidden], gb2[25] (OVERWRITTEN), and
 * pred[B][T-1][25][K] if pred != NULL. */
int64_t kc_train_step_workspace_bytes(int dtype, const kc_mlp *mlp, int64_t B, int64_t T, int32_t K);
int kc_train_step(int dtype, const kc_rod_params *P, const kc_mlp *mlp, int64_t B, int64_t T, int32_t K,
                  const int32_t *key_idx_host, const void *traj, const void *controls, double *loss, void *gW1,
                  void *gb1, void *gW2, void *gb2, void *pred, void *workspace, int64_t workspace_bytes,
                  void *stream);

/* torch.optim.Adam step (L2 weight decay folded into the gradient) followed by the reference's non-negative clamp of
 * the Linear weights (physics_train.py:199,296-304): n elements of param/grad/exp_avg/exp_avg_sq, step >= 1. */
int kc_adam_clamp(int dtype, int64_t n, void *param, const void *grad, void *exp_avg, void *exp_avg_sq,
                  int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                  int32_t clamp_min_zero, void *stream);

/* The same update for ALL parameter tensors in one launch, with the step count and the learning rate resident on the
 * device — nothing of the launch depends on host state, so a whole training step (kc_train_step, the gradient
 * all-reduce, this call) can be captured once in a CUDA graph and replayed (physics_train.py:289-304 per epoch).
 * *step_dev = number of updates applied so far (int64, advanced by the kernel), *lr_dev = learning rate (double,
 * rewritten by the host when ReduceLROnPlateau fires), ticket_dev = one zero-initialised int32 of scratch. */
typedef struct kc_adam_tensor {
    void *param; const void *grad; void *exp_avg, *exp_avg_sq; /* device pointers, dtype of the call */
    int64_t n;
    int32_t clamp_min_zero, reserved;
} kc_adam_tensor;
int kc_adam_clamp_multi(int dtype, int32_t n_tensors, const kc_adam_tensor *tensors_host, int64_t *step_dev,
                        const double *lr_dev, double beta1, double beta2, double eps, double weight_decay,
                        int32_t *ticket_dev, void *stream);

/* torch.optim.lr_scheduler.ReduceLROnPlateau('min', patience, factor) (physics_train.py:206,297) on the device:
 * *loss_dev (dtype of the call) is this epoch's loss; state_dev[6] = {best (init +inf), bad epochs (init 0), patience, factor,
 * relative threshold (1e-4), eps (1e-8)}; *lr_dev is the learning rate the optimiser kernels read.  Launch it after the
 * optimiser; it needs no host synchronisation, so a captured training step keeps its scheduler. */
int kc_plateau_step(int dtype, const void *loss_dev, double *state_dev, double *lr_dev, void *stream);

/* The gradient all-reduce of the training step (an all-reduce(sum) of the flat [gW1|gb1|gW2|gb2|loss] buffer, the
 * only collective of the path) fused with the optimiser, over NVLink peer memory instead of a collective library call:
 * regions_host[world] = device-visible pointers to every rank's region of kc_peer_region_bytes() bytes of symmetric
 * (peer-mapped) memory, zero-initialised once (torch.distributed._symmetric_memory in the Python host).
 *   kc_peer_publish     copies flat[n_flat] into the own region and signals every peer;
 *   kc_peer_gather_adam waits for all ranks' signals of this step, sums the ranks' copies in rank order (bitwise identical
 *                       on every rank), stores the sums back into flat and applies kc_adam_clamp_multi's update; the
 *                       tensors' gradients must be the consecutive leading views of `flat`; it advances *step_dev.
 * Both take *step_dev (updates applied so far) as the step number of the exchange, so the pair can be replayed from a CUDA
 * graph.  ticket_dev: two zero-initialised int32 (one per call).  world <= 8; world == 1 degenerates to Adam alone. */
int64_t kc_peer_region_bytes(int dtype, int64_t n_flat);
int kc_peer_publish(int dtype, int32_t world, int32_t rank, void *const *regions_host, int64_t n_flat, const void *flat,
                    const int64_t *step_dev, int32_t *ticket_dev, void *stream);
int kc_peer_gather_adam(int dtype, int32_t world, int32_t rank, void *const *regions_host, int64_t n_flat, void *flat,
                        int32_t n_tensors, const kc_adam_tensor *tensors_host, int64_t *step_dev, const double *lr_dev,
                        double beta1, double beta2, double eps, double weight_decay, int32_t *ticket_dev, void *stream);

/* Evaluation metrics of the training drivers for E (prediction, reference) pairs of rollouts, on the device:
 * dtw[E] = exact dynamic-time-warping distance with the L1 point distance between pred[e][:, 0:3, node] (Ta points) and
 * ref[e][:, 0:3, node] (Tb points) — what fastdtw(trajectory[:, :3, 9], tip_pos)[0] approximates (physics_train.py:159,
 * physics_multitrain.py:211; fastdtw is a third-party package the reference does not pin); and, if mse != NULL,
 * mse[E] = 1000 * mean of the squared position errors and squared 'zyx' Euler-angle differences over all nodes and time
 * indices (physics_multitrain.py:213-222, scipy Rotation.from_quat(q, scalar_first=True).as_euler('zyx')); NaN if Ta != Tb.
 * pred[E][Ta][rows][N], ref[E][Tb][rows][N] (rows >= 7: 25 or 50); outputs are float64 on the device. */
int kc_eval_metrics(int dtype, int64_t E, int64_t Ta, int64_t Tb, int32_t rows, int32_t N, int32_t node, const void *pred,
                    const void *ref, double *dtw, double *mse, void *stream);

/* State estimation from measurements (replaces estimate_state(data, tensions, robot),
 * knode_cosserat_realworld/estimate_state.py:158-242 with its helpers :11-156): data[B][T][7][N] (positions 0:3 and
 * quaternions 3:7 (w,x,y,z) on the full grid), tensions[B][T][4] -> est[B][T][25][N] in the state layout of the path.
 * L and del_t are robot.L and robot.del_t (the function reads them directly: arc lengths linspace(0, L, N), step L/N of the
 * n/m recursion, numpy.gradient spacing).  Bug-compatible with the reference for N != 10 (its recursion skips the write at
 * loop index 9 whatever N is).  T >= 3 (numpy.gradient with edge_order=2 needs three samples).  No workspace. */
int kc_estimate_state(int dtype, const kc_rod_params *P, double L, double del_t, int64_t B, int64_t T, const void *data,
                      const void *tensions, void *est, void *stream);

/* FMA-pipe micro-benchmark used as the compute-roofline denominator by bench.py: runs `iters` dependent-chain
 * FMAs x 8 chains per thread on a full grid and returns the number of FLOPs executed in *flops_host; time it with
 * CUDA events on `stream`. */
int kc_fma_peak(int dtype, int64_t iters, double *flops_host, void *scratch, void *stream);

/* Tensor-core self-test (tcgen05.mma kind::tf32, accumulators in TMEM): D[128][N] = A[128][K] * B[N][K]^T in fp32 storage,
 * one CTA.  Pins the operand layout / descriptor / TMEM read-back conventions that the fused KNODE MLP kernels rely on.
 * N % 16 == 0 in [16, 256], K % 8 == 0.  mode (mn_major) 0: tf32 operands, K-major; 1: tf32, MN-major (not a legal
 * no-swizzle layout for tf32 - kept to document the failure); 2: bf16 operands, MN-major, K % 16 == 0; 3: bf16, A written to
 * TMEM with tcgen05.st as packed pairs (".ts" MMA, the activation operand of the march and training kernels), B K-major,
 * K % 32 == 0; 4: as 3 with B staged as the K-major image of B^T and read as an MN-major operand (LBO = (N/8)*128, SBO = 128).
 * Modes 0, 2 and 3 are the layouts the fused kernels use. */
int kc_umma_selftest(const void *A, const void *B, void *D, int32_t N, int32_t K, int32_t mn_major, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KNODE_COSSERAT_H */
