"""Shared engine of the drop-in training scripts (physics_train.py, physics_multitrain.py, train_segment.py).

The reference trains with a Python loop over trajectories x time steps x nodes of eager torch ops and autograd
(physics_train.py:209-304 slow, :306-408 fast; train_segment.py:132-214).  Mathematically every epoch is ONE full-batch
loss over all (trajectory, step, key node) samples, so here an epoch is: kc_train_step (fused forward + 4-term loss +
reverse mode to the MLP weights) -> all-reduce of dW across ranks (only if torch.distributed is initialised) ->
kc_adam_clamp per parameter tensor.  ReduceLROnPlateau and the bookkeeping stay on the host exactly as in the reference.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

import _dist
import _ops


class PlateauLR:
    """torch.optim.lr_scheduler.ReduceLROnPlateau('min', patience, factor) with its defaults (threshold 1e-4 rel,
    cooldown 0, min_lr 0, eps 1e-8) — physics_train.py:206."""

    def __init__(self, lr, patience=80, factor=0.5, threshold=1e-4, eps=1e-8):
        self.lr, self.patience, self.factor, self.threshold, self.eps = lr, patience, factor, threshold, eps
        self.best, self.bad = float("inf"), 0

    def step(self, metric):
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            new = self.lr * self.factor
            if self.lr - new > self.eps:
                self.lr = new
            self.bad = 0
        return self.lr

    def get_last_lr(self):
        return [self.lr]


class TeacherForcedTrainer:
    """Full-batch teacher-forced KNODE training on the GPU.

    trajs: list of [T,25,N] tensors (or one [B,T,25,N]), controls: list of [T,4] (or [B,T,4]); all trajectories must
    share T (they do in the reference: train_len).  Under torch.distributed each rank keeps only its shard of the
    trajectories (presharded=True: the trajectories passed ARE this rank's shard); weights stay bitwise identical on all ranks because every rank applies the same all-reduced gradient.
    """

    def __init__(self, robot, trajs, controls, key_pt_idx, lr=1e-2, weight_decay=0.0, clamp_weight=True,
                 patience=80, factor=0.5, fused=True, use_graph=True, presharded=False):
        self.robot = robot
        self.fused, self.use_graph = fused, use_graph
        self.rank, self.world = _dist.world_info()
        traj = torch.stack(list(trajs)) if not torch.is_tensor(trajs) else trajs
        ctl = torch.stack(list(controls)) if not torch.is_tensor(controls) else controls
        if presharded:   # the caller hands every rank its own trajectories (weak scaling: the global batch grows with the ranks)
            self.n_total = traj.shape[0] * self.world
            lo, hi = 0, traj.shape[0]
        else:
            self.n_total = traj.shape[0]
            lo, hi = _dist.shard_range(self.n_total, self.rank, self.world)
        self.traj = traj[lo:hi].detach().contiguous()
        self.ctl = ctl[lo:hi].detach().to(self.traj.dtype).contiguous()
        self.key = np.asarray(key_pt_idx).reshape(-1)
        self.params = [p for p in robot.nn_models.parameters()]
        self.exp_avg = [torch.zeros_like(p.data) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p.data) for p in self.params]
        self.step_no = 0
        self.weight_decay = weight_decay
        self.clamp_weight = clamp_weight
        self.sched = PlateauLR(lr, patience, factor)
        self.loss_arr = []
        # physics_train.py:301-304: every parameter whose name contains 'weight' is clamped (both Linear layers)
        self.is_weight = ['weight' in n and 'layer1' not in n for n, _ in robot.nn_models.named_parameters()]

    # -- fused step: kc_train_step -> ONE all-reduce of the flat [gradients | loss] buffer -> kc_adam_clamp_multi --------
    def _fused_setup(self):
        weights = [p.data for p in self.params]
        self.plan = _ops.TrainStepPlan(self.robot._params(), weights, self.traj, self.ctl, self.key)
        self.adam = _ops.AdamClampMulti(weights, self.plan.grads,
                                        [self.clamp_weight and isw for isw in self.is_weight],
                                        lr=self.sched.get_last_lr()[0], weight_decay=self.weight_decay,
                                        exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq, step=self.step_no)
        self._lr_on_device = self.sched.get_last_lr()[0]
        self.graph = None
        self._eager_fused_steps = 0
        # ReduceLROnPlateau on the device (kc_plateau_step), stepped inside the graph on every epoch's loss: the host
        # scheduler object is replaced by a view of the device state
        self.device_sched = _ops.PlateauDevice(self.adam.lr_dev, patience=self.sched.patience, factor=self.sched.factor)
        self.device_sched.state[0] = self.sched.best
        self.device_sched.state[1] = float(self.sched.bad)
        self.sched = self.device_sched
        # the all-reduce fused with the update over NVLink peer memory (kc_peer_publish / kc_peer_gather_adam);
        # KC_PEER_ALLREDUCE=0 keeps the NCCL all-reduce + kc_adam_clamp_multi
        self.peer = False
        if self.world > 1 and os.environ.get("KC_PEER_ALLREDUCE", "1") != "0":
            self.peer = self.adam.enable_peer_allreduce(self.plan.flat)

    def _fused_body(self, train):
        self.plan.run()
        if train and self.peer:
            self.adam.run_peer()                     # publish -> wait for all ranks -> sum in rank order -> Adam + clamp
        else:
            if self.world > 1:
                dist.all_reduce(self.plan.flat)      # the only collective: gradients + loss, one launch
            if train:
                self.adam.run()
        if train:
            self.device_sched.run(self.plan.flat[-1:])   # scheduler.step(loss) (physics_train.py:297), on the device

    def fused_step(self, train=True, sync=True):
        """The same epoch as `step` with nothing but kernel launches on the critical path: outputs and workspace are
        allocated once, Adam runs as one launch with its step count / learning rate on the device, ReduceLROnPlateau is a
        device kernel too (kc_plateau_step), and from the third call on the whole step (kernels + all-reduce + Adam +
        scheduler) is ONE CUDA-graph launch.  sync=False skips the host read of the loss (loss_arr then misses this epoch;
        the scheduler has seen it all the same)."""
        if not hasattr(self, "plan"):
            self._fused_setup()
        P = self.robot._params()
        if P is not self.plan.P:                      # rod constants changed (setup_robot ...): a captured graph is stale
            self.plan.P = P
            self.graph = None
            self._eager_fused_steps = 0
        if not train:
            self._fused_body(False)
        elif self.graph is not None:
            self.graph.replay()
        elif self._eager_fused_steps < 2 or not self.use_graph:
            self._fused_body(True)                   # warm-up (also initialises NCCL) before capturing
            self._eager_fused_steps += 1
        else:
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._fused_body(True)
                self.graph = g                        # capture does not execute: run this epoch now
                self.graph.replay()
            except Exception as e:                    # e.g. a collective backend that cannot be captured
                self.use_graph = False
                self.graph = None
                torch.cuda.synchronize()
                print(f"[knode-cosserat_b200] CUDA-graph capture of the training step failed ({e}); running eagerly")
                self._fused_body(True)
        if train:
            self.step_no += 1
            for p, g in zip(self.params, self.plan.grads):
                p.grad = g
        if not sync:
            return None
        loss_val = float(self.plan.flat[-1].item())
        self.loss_arr.append(loss_val)
        return loss_val

    def close(self):
        """Release the captured graph (it holds NCCL kernels: do this before torch.distributed.destroy_process_group)."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize()
            self.graph = None

    def loss_and_grads(self):
        if self.traj.shape[0] > 0:
            loss, grads, _ = self.robot.teacher_forced_step(self.traj, self.ctl, self.key)
            grads = [g for g in grads]
            loss = loss.to(torch.float32) if self.world > 1 else loss
        else:  # a rank without trajectories contributes zeros
            grads = [torch.zeros_like(p.data) for p in self.params]
            loss = torch.zeros(1, dtype=torch.float32, device=self.params[0].device)
        if self.world > 1:
            _dist.allreduce_sum_(grads + [loss])
        return loss, grads

    def step(self, train=True):
        """One epoch of the reference loop: loss -> backward -> Adam -> scheduler -> clamp (physics_train.py:266-304)."""
        if self.fused:   # also on a rank whose shard is empty: kc_train_step returns zero gradients for B = 0, and every rank
            return self.fused_step(train)   # must apply the SAME Adam kernel for the weights to stay bitwise identical
        loss, grads = self.loss_and_grads()
        loss_val = float(loss.item())
        self.loss_arr.append(loss_val)
        if train:
            self.step_no += 1
            lr = self.sched.get_last_lr()[0]
            for p, g, m, v, isw in zip(self.params, grads, self.exp_avg, self.exp_avg_sq, self.is_weight):
                _ops.adam_clamp(p.data, g, m, v, self.step_no, lr=lr, weight_decay=self.weight_decay,
                                clamp=self.clamp_weight and isw)
                p.grad = g
            self.sched.step(loss_val)
        return loss_val

    def optim_state_dict(self):
        """Same shape as torch.optim.Adam.state_dict() so that checkpoints stay loadable by the reference's tooling."""
        return {"state": {i: {"step": torch.tensor(float(self.step_no)), "exp_avg": m, "exp_avg_sq": v}
                          for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))},
                "param_groups": [{"lr": self.sched.get_last_lr()[0], "betas": (0.9, 0.999), "eps": 1e-8,
                                  "weight_decay": self.weight_decay, "params": list(range(len(self.params)))}]}


class BpttTrainer:
    """Rollout training (north-star C3 ii): every step rolls the KNODE model out from the straight rod under the recorded
    tensions, compares the rollout with the target trajectories at the key nodes (the reference's 4-term loss), and
    back-propagates THROUGH the rollout (kc_rollout_bwd) into the MLP weights; then the same all-reduce / Adam / clamp as
    the teacher-forced trainer.  The reference has no such loop (it never differentiates a rollout); the
    optimiser settings are physics_train.py's (:198-207,289-304).

    targets [B,T,25,N], tensions [B,T,4]; under torch.distributed every rank keeps its shard (presharded=True: the tensors
    passed ARE the shard).  The loss is the mean over the GLOBAL batch, so the all-reduced gradient is the single-GPU one."""

    def __init__(self, robot, targets, tensions, key_pt_idx, lr=1e-3, weight_decay=0.0, clamp_weight=True, patience=80,
                 factor=0.5, presharded=False):
        self.robot = robot
        self.rank, self.world = _dist.world_info()
        n = targets.shape[0]
        if presharded:
            self.n_total, lo, hi = n * self.world, 0, n
        else:
            self.n_total = n
            lo, hi = _dist.shard_range(n, self.rank, self.world)
        self.params = [p for p in robot.nn_models.parameters()]
        weights = [p.data for p in self.params]
        self.is_weight = ['weight' in nm and 'layer1' not in nm for nm, _ in robot.nn_models.named_parameters()]
        self.sched = PlateauLR(lr, patience, factor)
        self.plan = _ops.BpttStepPlan(robot._params(), weights, tensions[lo:hi].detach(), targets[lo:hi].detach(),
                                      key_pt_idx, scale=1.0 / max(self.n_total, 1))
        self.adam = _ops.AdamClampMulti(weights, self.plan.grads, [clamp_weight and w for w in self.is_weight],
                                        lr=lr, weight_decay=weight_decay)
        self._lr_on_device = lr
        self.step_no = 0
        self.loss_arr = []
        self.peer = False
        if self.world > 1 and os.environ.get("KC_PEER_ALLREDUCE", "1") != "0":
            self.peer = self.adam.enable_peer_allreduce(self.plan.flat)

    def step(self, train=True, sync=True):
        self.plan.P = self.robot._params()
        lr = self.sched.get_last_lr()[0]
        if lr != self._lr_on_device:
            self.adam.set_lr(lr)
            self._lr_on_device = lr
        self.plan.run()
        if train and self.peer:
            self.adam.run_peer()
        elif self.world > 1:
            dist.all_reduce(self.plan.flat)
        if train and not self.peer:
            self.adam.run()
        if train:
            self.step_no += 1
            for p, g in zip(self.params, self.plan.grads):
                p.grad = g
        if not sync:
            return None
        loss_val = float(self.plan.flat[-1].item())
        self.loss_arr.append(loss_val)
        if train:
            self.sched.step(loss_val)
        return loss_val

    def converged(self):
        return bool(int(self.plan.fwd.iters.min()) >= 0)


def dtw_l1(a, b):
    """Exact dynamic-time-warping distance with the L1 point distance — what fastdtw(a, b)[0] approximates with its
    default radius=1 (physics_train.py:159, physics_multitrain.py:211).  a[Ta,d], b[Tb,d].
    The recurrence acc[i,j] = cost[i,j] + min(acc[i-1,j], acc[i,j-1], acc[i-1,j-1]) is evaluated one ANTI-DIAGONAL at a time
    (all cells of a diagonal depend only on the two previous ones), i.e. Ta+Tb-1 vectorised steps instead of Ta*Tb."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        a, b = a[:, None], b[:, None]
    cost = np.abs(a[:, None, :] - b[None, :, :]).sum(-1)
    Ta, Tb = cost.shape
    acc = np.full((Ta + 1, Tb + 1), np.inf)
    acc[0, 0] = 0.0
    for d in range(2, Ta + Tb + 1):                 # cells (i, j), 1-based, with i + j == d
        i = np.arange(max(1, d - Tb), min(Ta, d - 1) + 1)
        j = d - i
        acc[i, j] = cost[i - 1, j - 1] + np.minimum(np.minimum(acc[i - 1, j], acc[i, j - 1]), acc[i - 1, j - 1])
    return float(acc[Ta, Tb])


def transplant(np_robot, torch_robot):
    """Force the numpy-flavoured rod to use the torch MLP (physics_train.py:103-110,137-144)."""
    nn_model = torch_robot.nn_models
    np_robot.nn_model = nn_model
    np_robot.param_ls = [layer.detach().cpu().numpy() for _, layer in nn_model.state_dict().items()]
    np_robot.nn_path = 'whatever'  # Force the robot to use nn
