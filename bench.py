#!/usr/bin/env python
"""bench.py — the benchmark contract of this repo.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): batched forward rollout of the physics-only Cosserat rod,
4096 rods per GPU x N=10 nodes x T=100 time indices (99 solved steps), setup_robot parameters, fp32, synthetic
tensions (half sine, half random).  One "step" = one rollout of the whole batch through kc_rollout_fwd (rollout kernel +
layout transpose), inputs resident in HBM.  metric = rod-node-steps/s, whole job (all ranks).  Rods are independent:
ranks share nothing on the data path (weak scaling, no collective); the KNODE training step reported under "train" is
the only path with a collective (NCCL all-reduce of the 27,673 MLP gradients).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "knode-cosserat_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

B_PER_GPU, T_STEPS, N_NODES = 4096, 100, 10
F_ODE = 449.0                      # FLOP per physics node evaluation + Euler update (SURVEY §8 header)
E_REF = 15.0                       # nominal residual evaluations per time step (reference fsolve mean)
FLOP_PER_RNS = E_REF * (N_NODES - 1) / N_NODES * F_ODE   # 6.06 kFLOP per rod-node-step, physics only
BYTES_PER_RNS = 25 * 4 + 16.0 / N_NODES                  # 101.6 B per rod-node-step (fp32 trajectory + tensions)
TRAIN_B, TRAIN_T, TRAIN_H, TRAIN_KEYS = 1024, 30, 512, [3, 5, 7, 9]
FLOP_PER_TRAIN_SAMPLE = F_ODE + 106 * TRAIN_H + 156 * TRAIN_H  # 134.6 kFLOP


def _cpu_rollout_worker(args):
    """One rod of the workload on one host core with the reference's algorithm (oracle port of knode.simulate: numpy
    march + scipy fsolve)."""
    ctl, = args
    from oracle import rod_oracle as O
    P = O.setup_params(O.RodParams())
    t0 = time.perf_counter()
    O.rollout_fsolve(P, ctl)
    return time.perf_counter() - t0


def cpu_reference_sample(ctl_rods, cores):
    """Roll `len(ctl_rods)` rods out on `cores` host processes; returns (rod-node-steps/s, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_rollout_worker, [(c,) for c in ctl_rods], chunksize=1)
    wall = time.perf_counter() - t0
    n = sum((c.shape[0] - 1) * N_NODES for c in ctl_rods)
    return n / wall, wall


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) during the timed region."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x10: "sync_boost",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz, self.ok = index, False, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)   # the timed region is ~13 ms: several samples must fall inside it

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: numpy + scipy fsolve, one rod
    per host process) on a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from oracle import rod_oracle as O
    from physics_controls import synthetic_tensions
    P = O.setup_params(O.RodParams())
    cores = min(host_cores(), 64)
    t_sample = 26  # 25 solved steps per rod per bench step (~1-2 s of CPU work per core)
    ctl = synthetic_tensions(B_PER_GPU, T_STEPS, P.del_t, seed=0, dtype=np.float64)
    pick = np.linspace(0, B_PER_GPU - 1, cores).astype(int)
    rods = [ctl[i, :t_sample] for i in pick]
    for _ in range(args.warmup):
        cpu_reference_sample(rods[:cores], cores)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        cpu_reference_sample(rods, cores)
        n += len(rods) * (t_sample - 1) * N_NODES
    wall = time.perf_counter() - t0
    val = n / wall
    sample = f"{len(rods)} rods x {t_sample - 1} solved steps per bench step, one rod per host process"
    print(json.dumps({
        "impl": "reference", "metric": "rod-node-steps/sec", "value": val, "unit": "rod-node-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 batched forward rollout, physics-only rod, 4096 rods x 10 nodes x 100 time indices "
                               "(bounded sample of it per step)", "params": "setup_robot", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "rod-node-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rod-node-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"],
                    help="c2 (default): BASELINE configs[1], the headline; c5: configs[4], 8192 rods/GPU x 20 nodes x 500 time "
                         "indices, class-default rod (train_segment.py's configuration), physics-only and KNODE")
    ap.add_argument("--cpu-sample", type=int, nargs=2, default=None, help=argparse.SUPPRESS)  # T_sample cores
    args = ap.parse_args()
    if args.cpu_sample is not None:  # child process of the cpu_baseline leg (keeps fork() away from the CUDA context)
        from oracle import rod_oracle as O
        from physics_controls import synthetic_tensions
        t_sample, cores = args.cpu_sample
        P = O.setup_params(O.RodParams())
        c = synthetic_tensions(B_PER_GPU, T_STEPS, P.del_t, seed=0, dtype=np.float64)
        pick = np.linspace(0, B_PER_GPU - 1, cores).astype(int)
        val, wall = cpu_reference_sample([c[i, :t_sample] for i in pick], cores)
        print(json.dumps({"value": val, "wall": wall, "rods": len(pick)}))
        return
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import _kc
    import _ops
    from cosserat_ode import CosseratRod
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot, simulate
    from physics_controls import synthetic_tensions

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this repo has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = None
    if world > 1:
        import _dist
        numa_node = _dist.bind_to_gpu_numa_node(local_rank)   # pinned buffers land on the GPU's own socket
    if world > 1:
        # NCCL prints its version banner on stdout during initialisation; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config == "c5":
        run_c5(args, rank, world, dev, barrier, max_over_ranks)
        finish(world, None, args)
        return

    # ---------------- workload ----------------
    robot = CosseratRod(use_fsolve=True)
    setup_robot(robot)
    P = _kc.rod_params(robot)
    B, T = B_PER_GPU, T_STEPS
    ctl_host = synthetic_tensions(B, T, robot.del_t, seed=rank, dtype=np.float32)
    ctl = torch.tensor(ctl_host, device=dev)
    plan = _ops.RolloutPlan(P, None, B, T, torch.float32, dev, rows=25)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    for _ in range(args.warmup):
        plan.run(ctl)
    torch.cuda.synchronize()
    its = plan.iters.cpu().numpy()
    assert its.min() >= 0, "a rod failed to converge during warm-up"
    marches_mean = float(np.abs(its[:, 1:]).mean())
    rpw = 4   # the wide kernel (B <= 4736 per GPU) packs 4 rods per warp
    marches_warp = float(np.abs(its[:, 1:]).reshape(B // rpw, rpw, T - 1).max(1).mean())

    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        plan.run(ctl)
        ev[i][1].record()
    barrier()
    sampler.stop_flag = True
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    total_ms = max_over_ranks(total_ms)
    rns_per_step = B * N_NODES * (T - 1)
    value = world * rns_per_step * args.steps / (total_ms * 1e-3)

    # a step IS one launch of the rollout kernel (it writes the reference layout itself): the events around the step
    # bracket exactly that kernel, on this rank
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    fp32_peak = _ops.fma_peak(torch.float32, 40000, dev)
    achieved = rns_per_step * FLOP_PER_RNS / (kern_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_achieved = rns_per_step * BYTES_PER_RNS / (kern_ms * 1e-3) / 1e9
    executed_flop = rns_per_step / N_NODES * (N_NODES - 1) * marches_warp * 7 * 337.0  # 7 lanes x (120 FFMA*2 + 71 FMUL + 26 FADD) per node

    # ---------------- end to end through the public API (host buffers in, host buffers out) ----------------
    pinned = torch.empty((B, T, 25, N_NODES), dtype=torch.float32).pin_memory()
    ctl_pinned = torch.from_numpy(ctl_host).pin_memory()      # the step's inputs live in pinned host memory (fp32)
    for _ in range(2):
        simulate(robot, ctl_pinned, dtype=np.float32, rows=25, pinned_out=pinned)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        simulate(robot, ctl_pinned, dtype=np.float32, rows=25, pinned_out=pinned)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * rns_per_step * args.steps / e2e_s

    # ---------------- end to end for a caller that only reads the tip positions (physics_train.py:159) ----------------------
    tip_pinned = torch.empty((B, T, 3, 1), dtype=torch.float32).pin_memory()
    tip_sel = ([0, 1, 2], [N_NODES - 1])
    for _ in range(2):
        simulate(robot, ctl_pinned, dtype=np.float32, rows=25, pinned_out=tip_pinned, select=tip_sel)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        simulate(robot, ctl_pinned, dtype=np.float32, rows=25, pinned_out=tip_pinned, select=tip_sel)
    barrier()
    tip_s = max_over_ranks(time.perf_counter() - t0)
    e2e_tip = {"value": world * rns_per_step * args.steps / tip_s, "unit": "rod-node-steps/s", "ms_per_step": tip_s / args.steps * 1e3,
               "api": "knode.simulate(..., select=([0,1,2],[N-1])): same rollout, only the tip positions cross PCIe",
               "h2d_bytes_per_step": int(ctl_host.nbytes), "d2h_bytes_per_step": int(tip_pinned.numel() * 4)}

    # ---------------- the same API under the REFERENCE's contract: float64 host list/array in -> fresh float64 [B,T,50,N] ----
    e2e_ref = None
    if not args.no_train:
        ctl64 = ctl_host.astype(np.float64)
        simulate(robot, ctl64)
        barrier()
        t0 = time.perf_counter()
        n_rc = max(2, min(args.steps, 3))
        for _ in range(n_rc):
            out64 = simulate(robot, ctl64)
        barrier()
        rc_s = max_over_ranks(time.perf_counter() - t0) / n_rc
        e2e_ref = {"value": world * rns_per_step / rc_s, "unit": "rod-node-steps/s", "ms_per_step": rc_s * 1e3,
                   "api": "knode.simulate(robot, ctl[B,T,4] float64) -> fresh float64 ndarray [B,T,50,N] (the reference's own "
                          "contract, knode.py:55-102: fp64 arithmetic, rows 25:50 = yh, zh, pageable output)",
                   "h2d_bytes_per_step": int(ctl64.nbytes), "d2h_bytes_per_step": int(out64.nbytes), "dtype": "f64"}
        del out64

    # ---------------- fp64 rollout (the reference's only rollout arithmetic is fp64: knode.py:55-102, cosserat_ode.py) ------
    f64 = None
    if not args.no_train:
        plan64 = _ops.RolloutPlan(P, None, B, T, torch.float64, dev, rows=25)
        ctl64_dev = ctl.double()
        for _ in range(2):
            plan64.run(ctl64_dev)
        barrier()
        ms64 = []
        for _ in range(max(3, min(args.steps, 5))):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan64.run(ctl64_dev)
            e1.record()
            e1.synchronize()
            ms64.append(e0.elapsed_time(e1))
        ms64_med = max_over_ranks(float(np.median(ms64)))
        fp64_peak = _ops.fma_peak(torch.float64, 20000, dev)
        f64 = {"metric": "rod-node-steps/sec, fp64", "value": world * rns_per_step / (ms64_med * 1e-3), "unit": "rod-node-steps/s",
               "ms_per_step": ms64_med, "dtype": "f64", "all_converged": bool(int(plan64.iters.min()) >= 0),
               "marches_per_step_mean": float(plan64.iters[:, 1:].abs().float().mean().item()), "tol": 1e-11,
               "roofline": {"bound": "fp64", "kernel": "kc_rollout_wide_kernel<double> (8 lanes per rod, Newton) + layout transpose",
                            "achieved": rns_per_step * FLOP_PER_RNS / (ms64_med * 1e-3) / 1e12, "peak": fp64_peak / 1e12,
                            "unit": "TFLOP/s", "frac": rns_per_step * FLOP_PER_RNS / (ms64_med * 1e-3) / fp64_peak,
                            "peak_source": "kc_fma_peak(fp64) micro-benchmark measured live in this run",
                            "normalisation": "6.06 kFLOP per rod-node-step (as the fp32 line)"}}
        del plan64, ctl64_dev

    # ---------------- C1 (BASELINE configs[0]): ONE class-default rod, 200 steps, through knode.simulate itself ------------
    c1 = None
    if not args.no_train and rank == 0:
        from physics_controls import calc_controls
        r1 = CosseratRod(use_fsolve=True)                      # class-default parameters, no setup_robot
        ctl1 = np.array(calc_controls('sine', 1.0, r1.del_t, 200))
        simulate(r1, ctl1)
        t0 = time.perf_counter()
        for _ in range(5):
            tr1 = simulate(r1, ctl1)
        c1_s = (time.perf_counter() - t0) / 5
        c1 = {"metric": "rod-node-steps/sec, single rod", "value": 199 * N_NODES / c1_s, "unit": "rod-node-steps/s",
              "ms_per_call": c1_s * 1e3, "dtype": "f64", "tip_xyz_index_29": [float(v) for v in tr1[29, :3, -1]],
              "api": "knode.simulate(CosseratRod(use_fsolve=True), calc_controls('sine', 1.0, 0.005, 200)) -> float64 [200,50,10]; "
                     "one rod is one dependent chain of 199 shooting solves: latency, not throughput",
              "cpu_reference_published": "361 rod-node-steps/s (the reference itself, one core, SURVEY 6)"}

    # ---------------- KNODE training step (C3/C4): teacher-forced fwd + loss + bwd + allreduce + Adam + clamp -----------
    train = None
    if not args.no_train:
        torch.manual_seed(0)
        trobot = CosseratRodTorch(str(dev), TRAIN_H)
        setup_robot(trobot)
        # C4: the 1024 trajectories are sharded across the ranks by the trainer itself (rank r keeps shard r)
        from _train import TeacherForcedTrainer
        trainer = TeacherForcedTrainer(trobot, plan.traj[:TRAIN_B, :TRAIN_T].contiguous(),
                                       ctl[:TRAIN_B, :TRAIN_T].contiguous(), TRAIN_KEYS, lr=1e-2)

        def train_step(sync):
            # kc_train_step -> all-reduce of [gradients | loss] -> kc_adam_clamp_multi; one CUDA-graph launch per step
            # from the third call on.  ReduceLROnPlateau (physics_train.py:206,297) is stepped on every epoch's loss by a
            # device kernel inside the same graph (kc_plateau_step); sync=True additionally reads the loss back to the
            # host every step (what a caller that logs every epoch pays), sync=False leaves it on the device.
            trainer.fused_step(train=True, sync=sync)

        def time_train(sync):
            for _ in range(args.warmup):
                train_step(sync)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                train_step(sync)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1)) / args.steps

        tms = time_train(False)
        tms_hostread = time_train(True)
        q = TRAIN_B * (TRAIN_T - 1) * len(TRAIN_KEYS)
        identical = None
        if world > 1:   # C4: every rank applied the same all-reduced gradient -> weights bitwise identical on all ranks
            wflat = torch.cat([p_.data.reshape(-1) for p_ in trobot.nn_models.parameters()])
            wlo, whi = wflat.clone(), wflat.clone()
            dist.all_reduce(wlo, op=dist.ReduceOp.MIN)
            dist.all_reduce(whi, op=dist.ReduceOp.MAX)
            identical = bool(torch.equal(wlo, whi))
        train = {"metric": "KNODE train steps/sec", "value": 1e3 / tms, "unit": "steps/s", "ms_per_step": tms,
                 "global_batch_trajectories": TRAIN_B, "samples_per_step": q, "scaling": "strong",
                 "semantics": "teacher-forced (physics_train.py), fwd+loss+bwd+allreduce+Adam+clamp+plateau scheduler through "
                              "_train.TeacherForcedTrainer.fused_step (CUDA graph: %s; all-reduce: %s)"
                              % (trainer.graph is not None, "peer memory, fused with Adam" if trainer.peer else
                                 ("NCCL" if world > 1 else "none (1 GPU)")),
                 "useful_tflops": q * FLOP_PER_TRAIN_SAMPLE / (tms * 1e-3) / 1e12, "hidden": TRAIN_H,
                 "roofline": {"bound": "tensor", "kernel": "kc_train_tc3_kernel (tcgen05 + TMEM, transposed backward: activations stay in TMEM; ~78 % of the step)",
                              "achieved": q * FLOP_PER_TRAIN_SAMPLE / world / (tms * 1e-3) / 1e12, "unit": "TFLOP/s",
                              "note": "useful fp32-accurate FLOP/s per GPU of the WHOLE step (134.6 kFLOP per sample, SURVEY "
                                      "8d); every contraction runs 3 tensor-core passes (bf16 hi-lo split) to keep "
                                      "fp32 accuracy, so the executed tensor FLOP/s are 3x this",
                              "frac_of_fp32_fma_peak": None, "frac_of_bf16_tensor_peak_executed": None},
                 "loss": float(trainer.plan.flat[-1].item()), "kernels_per_step": 4 + (2 if trainer.peer else 1) + 1,
                 "scheduler": "ReduceLROnPlateau(patience 80, factor 0.5) stepped on every epoch's loss by kc_plateau_step inside "
                              "the captured step (physics_train.py:206,297); learning rate now %g" % trainer.sched.get_last_lr()[0],
                 "train_hostread": {"value": 1e3 / tms_hostread, "unit": "steps/s", "ms_per_step": tms_hostread,
                                    "note": "same step plus a host read of the loss after every step (loss logging every epoch)"},
                 "weights_bitwise_identical_across_ranks": identical}

    # ---------------- the same training step, weak scaling: 1024 trajectories PER GPU (global batch grows with N) -------
    train_weak = None
    if train is not None and world > 1:
        torch.manual_seed(0)
        wrobot = CosseratRodTorch(str(dev), TRAIN_H)
        setup_robot(wrobot)
        wtrainer = TeacherForcedTrainer(wrobot, plan.traj[:TRAIN_B, :TRAIN_T].contiguous(),
                                        ctl[:TRAIN_B, :TRAIN_T].contiguous(), TRAIN_KEYS, lr=1e-2, presharded=True)
        for _ in range(args.warmup):
            wtrainer.fused_step(train=True, sync=False)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            wtrainer.fused_step(train=True, sync=False)
        e1.record()
        barrier()
        wms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        train_weak = {"metric": "KNODE train steps/sec, 1024 trajectories per GPU", "value": 1e3 / wms, "unit": "steps/s",
                      "ms_per_step": wms, "global_batch_trajectories": TRAIN_B * world, "scaling": "weak",
                      "trajectory_steps_per_s": TRAIN_B * world * 1e3 / wms,
                      "useful_tflops": world * q * FLOP_PER_TRAIN_SAMPLE / (wms * 1e-3) / 1e12,
                      "semantics": "same fused step as `train` (kc_train_step -> one all-reduce -> kc_adam_clamp_multi, CUDA "
                                   "graph) with every rank holding its own 1024 trajectories; compare trajectory_steps_per_s "
                                   "with 1024 x train.value at N = 1"}

    # ---------------- KNODE rollout + rollout training step (C3 ii, north-star subsystems 2 and 3) ----------------------
    # MLP of the march on tcgen05 (csrc/kc_knode_tc.cu).  Executed tensor work per node evaluation of a 128-row CTA:
    # forward 2 GEMMs x 128 x 32 x 512 MAC, backward 3 GEMMs, each in 3 bf16 hi/lo passes.
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0)) * 1e12
    MMA_FWD = 2 * 128 * 32 * 512 * 2 * 3.0
    MMA_BWD = 3 * 128 * 32 * 512 * 2 * 3.0
    F_MLP = 106.0 * TRAIN_H
    FLOP_PER_RNS_KNODE = E_REF * (N_NODES - 1) / N_NODES * (F_ODE + F_MLP)      # 738.7 kFLOP

    def knode_model(seed):
        torch.manual_seed(seed)
        r_ = CosseratRodTorch(str(dev), TRAIN_H)
        setup_robot(r_, "youngs")
        with torch.no_grad():   # a trained-size residual (the reference's random init is unstable inside a rollout)
            r_.nn_models[2].weight.mul_(0.02)
            r_.nn_models[2].bias.mul_(0.02)
        return r_

    def group_marches(iters_bt):
        """Marches the tensor-core forward kernel executes (each is one 128-row MLP per node and CTA): a CTA's rods march
        until the slowest one has converged.  Mirrors the launcher's choice (csrc/kc_knode_tc.cu): 16 rods x 8 rows per CTA,
        or — beyond 3 rounds of such CTAs — one row per rod."""
        a = iters_bt[:, 1:].abs().float()
        nb = a.shape[0]
        if nb > 3 * 148 * 16:
            rounds = -(-nb // (128 * 148))
            rpc = min(128, max(32, -(-(-(-nb // (rounds * 148))) // 32) * 32))
            lo = 1
        else:
            rpc, lo = 16, 2
        pad = (-nb) % rpc
        if pad:
            a = torch.cat([a, a[-1:].expand(pad, -1)])
        return float(a.view(-1, rpc, a.shape[1]).amax(1).clamp(min=lo).sum().item()), rpc

    knode = None
    bptt = None
    bptt_weak = None
    if not args.no_train:
        # (a) forward KNODE rollout at a batch that fills the chip: 148 CTAs x 128 rows = 18 944 rods per GPU
        KB, KT = 148 * 128, TRAIN_T
        krobot = knode_model(1)
        ksd = krobot.nn_models.state_dict()
        kmlp = _ops.Mlp(ksd["0.weight"], ksd["0.bias"], ksd["2.weight"], ksd["2.bias"])
        ktens = torch.tensor(synthetic_tensions(KB, KT, krobot.del_t, seed=100 + rank), device=dev)
        kplan = _ops.RolloutPlan(krobot._params(), kmlp, KB, KT, torch.float32, dev, rows=25)
        for _ in range(2):
            kplan.run(ktens)
        barrier()
        kms = []
        for _ in range(max(3, min(args.steps, 5))):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            kplan.run(ktens)
            e1.record()
            e1.synchronize()
            kms.append(e0.elapsed_time(e1))
        kms_med = max_over_ranks(float(np.median(kms)))
        k_rns = KB * N_NODES * (KT - 1)
        k_marches, k_rpc = group_marches(kplan.iters)
        k_exec = k_marches * (N_NODES - 1) * MMA_FWD
        knode = {"metric": "KNODE rod-node-steps/sec (forward rollout, MLP in the march)", "value": world * k_rns / (kms_med * 1e-3),
                 "unit": "rod-node-steps/s", "ms_per_rollout": kms_med, "rods_per_gpu": KB, "time_indices": KT, "hidden": TRAIN_H,
                 "dtype": "f32", "all_converged": bool(int(kplan.iters.min()) >= 0),
                 "marches_per_step_mean": float(kplan.iters[:, 1:].abs().float().mean().item()),
                 "roofline": {"bound": "tensor", "kernel": "kc_knode_tc_fwd1_kernel (tcgen05 kind::f16, activation tile in TMEM, "
                              ".ts MMA; one row per rod, %d rods per CTA, Broyden in lock step)" % k_rpc,
                              "achieved": k_exec / (kms_med * 1e-3) / 1e12, "peak": bf16_peak / 1e12, "unit": "TFLOP/s",
                              "frac": k_exec / (kms_med * 1e-3) / bf16_peak,
                              "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst; the kernel is timed alone)",
                              "what": "EXECUTED tensor FLOP/s: every march evaluates the MLP for a 128-row tile per CTA in 3 bf16 "
                                      "hi/lo passes (fp32-grade products), K = 32 inputs / N = 32 outputs padded; a CTA "
                                      "marches until its slowest rod has converged",
                              "useful_tflops": k_rns * FLOP_PER_RNS_KNODE / (kms_med * 1e-3) / 1e12,
                              "useful_frac": k_rns * FLOP_PER_RNS_KNODE / (kms_med * 1e-3) / bf16_peak,
                              "useful_normalisation": "738.7 kFLOP per rod-node-step = 15 nominal evaluations x 9/10 x (449 + "
                                                      "106 H) FLOP (SURVEY 8d)"}}
        del kplan, ktens

        # (b) rollout training: 29-step KNODE rollout from the straight rod + fused 4-term loss + BPTT + all-reduce + Adam
        from _train import BpttTrainer

        def bptt_leg(presharded):
            brobot = knode_model(1)
            nb = TRAIN_B if presharded else TRAIN_B // world
            lo = 0 if presharded else rank * nb
            tr = BpttTrainer(brobot, plan.traj[lo:lo + nb, :TRAIN_T].contiguous(), ctl[lo:lo + nb, :TRAIN_T].contiguous(),
                             TRAIN_KEYS, lr=1e-4, presharded=True)
            for _ in range(3):
                tr.step(sync=False)
            barrier()
            n_ = max(3, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_):
                tr.step(sync=True)       # the loss is read every step: the plateau scheduler runs as in physics_train.py
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1)) / n_
            its_ = tr.plan.fwd.iters
            exec_flop = group_marches(its_)[0] * (N_NODES - 1) * MMA_FWD + \
                ((nb + 15) // 16) * (TRAIN_T - 1) * (N_NODES - 1) * MMA_BWD
            identical = None
            if world > 1:
                wflat = torch.cat([p_.data.reshape(-1) for p_ in brobot.nn_models.parameters()])
                wlo, whi = wflat.clone(), wflat.clone()
                dist.all_reduce(wlo, op=dist.ReduceOp.MIN)
                dist.all_reduce(whi, op=dist.ReduceOp.MAX)
                identical = bool(torch.equal(wlo, whi))
            gb = nb * world
            return {"metric": "KNODE rollout-BPTT train steps/sec", "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms,
                    "global_batch_trajectories": gb, "trajectories_per_gpu": nb, "rollout_steps": TRAIN_T - 1,
                    "hidden": TRAIN_H, "steps": n_, "scaling": "weak" if presharded else "strong",
                    "trajectory_steps_per_s": gb * 1e3 / ms,
                    "semantics": "29-step KNODE rollout from the straight rod (kc_rollout_fwd, MLP on tcgen05) + fused 4-term "
                                 "loss and cotangent at the key nodes (kc_rollout_loss) + BPTT (kc_rollout_bwd: joint adjoint "
                                 "march, MLP input-VJP on tcgen05, weight gradients by the tensor-core sample reduction) + "
                                 "one all-reduce of [gradients | loss] + kc_adam_clamp_multi; loss read every step",
                    "loss": tr.loss_arr[-1], "all_converged": tr.converged(),
                    "rod_node_steps_per_s_fwd_bwd": gb * N_NODES * (TRAIN_T - 1) / (ms * 1e-3),
                    "weights_bitwise_identical_across_ranks": identical,
                    "roofline": {"bound": "tensor", "kernel": "kc_knode_tc_fwd_kernel + kc_knode_tc_bwd_kernel",
                                 "achieved": exec_flop / (ms * 1e-3) / 1e12, "peak": bf16_peak / 1e12, "unit": "TFLOP/s",
                                 "frac": exec_flop / (ms * 1e-3) / bf16_peak,
                                 "what": "executed tensor FLOP/s per GPU of the WHOLE step (3 bf16 hi/lo passes, 8 rows per "
                                         "rod); at %d rods per GPU only %d of 148 SMs hold a CTA — the step is bound by the "
                                         "dependent chain of node evaluations, not by the tensor pipe" % (nb, (nb + 15) // 16)}}

        bptt = bptt_leg(False)
        if world > 1:
            bptt_weak = bptt_leg(True)

    # ---------------- state estimation: estimate_state on recordings resident in HBM ----------------
    estimate = None
    if not args.no_train:
        EB, ET = 64, 6000                     # recordings per GPU x time steps (60 s at 100 Hz), class-default robot, fp64
        erobot = CosseratRod()
        s_ = torch.linspace(0, erobot.L, N_NODES, device=dev, dtype=torch.float64)[None, None, :]
        t_ = (torch.arange(ET, device=dev, dtype=torch.float64) * erobot.del_t)[None, :, None]
        om = 2 * np.pi / (torch.rand(EB, 1, 1, device=dev, dtype=torch.float64) * 25 + 15) / erobot.del_t
        one = torch.ones(EB, ET, N_NODES, device=dev, dtype=torch.float64)
        meas = torch.stack([0.3 * torch.sin(om * t_) * s_ ** 2, 0.2 * torch.cos(1.3 * om * t_) * s_ ** 2, s_ * one, one,
                            0.8 * s_ * torch.sin(om * t_), 0.6 * s_ * torch.cos(0.7 * om * t_),
                            0.3 * s_ * torch.sin(0.4 * om * t_ + 1)], 2).contiguous()
        etens = 5 + 5 * torch.rand(EB, ET, 4, device=dev, dtype=torch.float64)
        eP = _kc.rod_params(erobot)
        for _ in range(3):
            est = _ops.estimate_state(eP, erobot.L, erobot.del_t, meas, etens)
        barrier()
        ems = []
        for _ in range(max(args.steps, 5)):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            est = _ops.estimate_state(eP, erobot.L, erobot.del_t, meas, etens)
            e1.record()
            e1.synchronize()
            ems.append(e0.elapsed_time(e1))
        ems_med = max_over_ranks(float(np.median(ems)))
        ebytes = EB * ET * (32 * N_NODES + 4) * 8     # 7 rows in + 25 rows out per node, 4 tensions per step
        hbm = float(peaks.get("hbm_gbs", 6544.7))
        estimate = {"metric": "estimate_state rod-node-steps/sec", "value": world * EB * ET * N_NODES / (ems_med * 1e-3),
                    "unit": "rod-node-steps/s", "ms_per_call": ems_med, "recordings_per_gpu": EB, "time_steps": ET,
                    "dtype": "f64", "finite": bool(torch.isfinite(est).all().item()),
                    "roofline": {"bound": "hbm", "kernel": "kc_estimate_kernel<double,10>", "achieved": ebytes / (ems_med * 1e-3) / 1e9,
                                 "peak": hbm, "unit": "GB/s", "frac": ebytes / (ems_med * 1e-3) / 1e9 / hbm,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)",
                                 "algorithmic_bytes_per_rod_node_step": (32 * N_NODES + 4) * 8 / N_NODES,
                                 "algorithmic_bytes": ebytes, "traffic": 4.048e9 * EB / 256,
                                 "traffic_source": "ncu dram__bytes_read+write of one launch at 256 recordings in the round-2 "
                                                   "build (1.032 GB + 3.015 GB, profiles/r02_ncu_prof_estimate.csv), scaled "
                                                   "to this batch"}}

    # ---------------- config 5 (BASELINE configs[4]): this rank's 8192-rod shard of the 65 536-rod sweep -----------------------
    config5 = None
    if not args.no_train:
        del plan
        torch.cuda.empty_cache()
        config5 = measure_c5(args, rank, world, dev, barrier, max_over_ranks, 2)
        config5.pop("e2e", None)
        config5.pop("cpu_baseline", None)
        config5["note"] = ("every rank rolls its own 8192-rod shard out (no collective): at --gpus 8 this IS configuration 5 "
                           "(65 536 rods x 20 nodes x 500 time indices); `python bench.py --config c5` prints it as a contract line")

    # ---------------- CPU baseline (rank 0, bounded sample, the reference's algorithm on the host cores) -----------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        t_sample = 51
        import subprocess
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-sample", str(t_sample), str(cores)],
                           capture_output=True, text=True, check=True)
        cj = json.loads(r.stdout.strip().splitlines()[-1])
        cval, cwall, nrods = cj["value"], cj["wall"], cj["rods"]
        cpu = {"value": cval, "unit": "rod-node-steps/s", "cores": cores, "kind": "port",
               "sample": f"{nrods} rods x {t_sample - 1} solved steps of the same workload, oracle port of "
                         f"knode.simulate (numpy + scipy fsolve), one rod per host process, {cwall:.1f} s wall"}

    if cpu is not None and estimate is not None:
        # the same checker leg for the state estimator: oracle port of estimate_state (numpy, vectorised over time, one
        # core) on a bounded sample of one recording
        from oracle import estimate_oracle as EO
        from oracle import rod_oracle as RO
        n_s = 1200
        m_np, c_np = meas[0, :n_s].cpu().numpy(), etens[0, :n_s].cpu().numpy()
        t0 = time.perf_counter()
        EO.estimate_state(RO.RodParams(), m_np, c_np)
        dt_cpu = time.perf_counter() - t0
        estimate["cpu_baseline"] = {"value": n_s * N_NODES / dt_cpu, "unit": "rod-node-steps/s", "cores": 1, "kind": "port",
                                    "sample": f"{n_s} time steps of one recording, oracle port of estimate_state (numpy), "
                                              f"{dt_cpu:.2f} s wall"}
    if train is not None:
        ach = train["roofline"]["achieved"] * 1e12
        train["roofline"]["frac_of_fp32_fma_peak"] = ach / fp32_peak
        bf16_peak = float(peaks.get("bf16_tflops_sustained", 1397.8))
        train["roofline"]["frac_of_bf16_tensor_peak_executed"] = 3.0 * ach / (bf16_peak * 1e12)
        # plain roofline fields for the stated pass mode: executed tensor FLOP/s (3 bf16 hi/lo passes per contraction) of the
        # whole step against the sustained dense-bf16 tensor peak (the kernel runs inside a long captured step)
        train["roofline"]["pass_mode"] = "3 x bf16 (hi*hi + lo*hi + hi*lo), fp32 accumulate"
        train["roofline"]["executed"] = 3.0 * ach / 1e12
        train["roofline"]["peak"] = bf16_peak
        train["roofline"]["frac"] = 3.0 * ach / (bf16_peak * 1e12)
        train["roofline"]["peak_source"] = ("MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks
                                            else "fallback 1397.8 TFLOP/s (B200_PROFILING.md)")
    if rank == 0:
        out = {
            "metric": "rod-node-steps/sec", "value": value, "unit": "rod-node-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 batched forward rollout, physics-only Cosserat rod: 4096 rods/GPU x 10 nodes x "
                                   "100 time indices (99 solved steps), setup_robot params, half sine / half random "
                                   "tensions", "rods_per_gpu": B, "nodes": N_NODES, "time_indices": T,
                       "output": "traj[B,T,25,N] fp32 in the reference layout, resident in HBM",
                       "l2": "410 MB written per step (> 126 MB L2) plus an explicit 256 MB flush between timed "
                             "iterations", "parallelism": f"rods sharded over {world} rank(s), no collective",
                       "solver": {"marches_per_step_mean": marches_mean, "marches_per_step_warp": marches_warp,
                                  "nominal_evals_per_step": E_REF}},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "rod-node-steps/s", "h2d_bytes_per_step": int(ctl_host.nbytes),
                    "d2h_bytes_per_step": int(pinned.numel() * 4), "api": "knode.simulate(robot, ctl[B,T,4], "
                    "dtype=float32, rows=25) pinned host tensions in -> pinned host trajectory out = one kc_rollout_host C-ABI call (H2D, "
                    "rollout in time ranges, D2H of each finished range overlapped with the next)",
                    "ms_per_step": e2e_s / args.steps * 1e3, "numa_node_rank0": numa_node},
            "gpu_launches": 1 * args.steps,
            "roofline": {"bound": "fp32", "kernel": "kc_rollout_wide_lin_kernel<float,diag,physics,N=10> (8 lanes per rod: "
                         "Newton + per-march FD Jacobian + linearised final correction; writes traj[B,T,25,N] directly)",
                         "achieved": achieved / 1e12,
                         "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "peak_source": "kc_fma_peak micro-benchmark measured live in this run (MEASURED_PEAKS.json has "
                                        "no FP32-pipe figure)", "normalisation": "6.06 kFLOP per rod-node-step = 15 "
                         "nominal residual evaluations x 9/10 x 449 FLOP (SURVEY 8d)",
                         "kernel_ms": kern_ms, "executed_tflops": executed_flop / (kern_ms * 1e-3) / 1e12,
                         "traffic": 364.9e6, "traffic_source": "ncu --set full of this kernel in the round-2 build (7.5 MB read + 357.4 MB "
                         "written per launch), profiles/r02_ncu_prof_rollout_lin.csv (algorithmic 416 MB: 410 MB trajectory written once + "
                         "6.5 MB tensions read; the tail of the trajectory is still in L2 when the kernel ends)",
                         "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak}},
            "cpu_baseline": cpu, "train": train, "train_weak": train_weak, "knode_rollout": knode, "train_bptt": bptt,
            "config5": config5, "rollout_f64": f64, "c1_single_rod": c1, "e2e_reference_contract": e2e_ref, "e2e_tip_only": e2e_tip,
            "train_bptt_weak": bptt_weak, "estimate_state": estimate}
        print(json.dumps(out))
    sys.stdout.flush()
    finish(world, None if args.no_train else trainer, args)


def finish(world, trainer, args):
    import torch
    import torch.distributed as dist
    if world > 1:
        # a CUDA graph that captured NCCL kernels must be released before its communicator is torn down (otherwise
        # destroy_process_group can block forever); a watchdog turns any remaining teardown hang into a clean exit
        import gc
        import threading
        if trainer is not None:
            trainer.close()
        gc.collect()
        torch.cuda.synchronize()
        wd = threading.Timer(30.0, lambda: os._exit(0))
        wd.daemon = True
        wd.start()
        dist.barrier()
        dist.destroy_process_group()
        wd.cancel()


def run_c5(args, rank, world, dev, barrier, max_over_ranks):
    out = measure_c5(args, rank, world, dev, barrier, max_over_ranks, max(2, min(args.steps, 5)))
    if rank == 0:
        print(json.dumps(out))
    sys.stdout.flush()


def measure_c5(args, rank, world, dev, barrier, max_over_ranks, steps):
    """BASELINE configs[4]: 65 536 rods x 20 nodes x 500 time indices over 8 GPUs = 8 192 rods per GPU, the
    configuration of train_segment.py (class-default physical parameters: it never calls setup_robot; dt = 0.005; H = 512).
    Physics-only rollout (the contract line's value) and the KNODE rollout (MLP in the march on tcgen05)."""
    import torch
    import _kc
    import _ops
    from cosserat_ode import CosseratRod
    from cosserat_ode_torch import CosseratRodTorch
    from physics_controls import synthetic_tensions
    B, T, N = 8192, 500, 20
    robot = CosseratRod(use_fsolve=True)
    robot.N = N
    robot.compute_intermediate_terms()
    P = _kc.rod_params(robot)
    ctl = torch.tensor(synthetic_tensions(B, T, robot.del_t, seed=rank, dtype=np.float32), device=dev)
    flop_rns = E_REF * (N - 1) / N * F_ODE
    sampler = ClockSampler(dev.index)
    plan = _ops.RolloutPlan(P, None, B, T, torch.float32, dev, rows=25)
    for _ in range(2):
        plan.run(ctl)
    barrier()
    sampler.start()
    ev = []
    for _ in range(steps):     # 8.2 GB written per step: far beyond L2, no flush needed
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(ctl)
        e1.record()
        ev.append((e0, e1))
    barrier()
    sampler.stop_flag = True
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev)) / steps
    its = plan.iters
    ok = bool(int(its.min()) >= 0)
    marches = float(its[:, 1:].abs().float().mean().item())
    rns = B * N * (T - 1)
    fp32_peak = _ops.fma_peak(torch.float32, 40000, dev)
    del plan
    torch.cuda.empty_cache()
    # KNODE on: H = 512 (train_segment.py:13), a trained-size residual
    torch.manual_seed(1)
    kr = CosseratRodTorch(str(dev), 512)
    kr.N = N
    kr.compute_intermediate_terms()
    with torch.no_grad():
        kr.nn_models[2].weight.mul_(0.02)
        kr.nn_models[2].bias.mul_(0.02)
    sd = kr.nn_models.state_dict()
    mlp = _ops.Mlp(sd["0.weight"], sd["0.bias"], sd["2.weight"], sd["2.bias"])
    kplan = _ops.RolloutPlan(kr._params(), mlp, B, T, torch.float32, dev, rows=25)
    kplan.run(ctl)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kplan.run(ctl)
    e1.record()
    barrier()
    kms = max_over_ranks(e0.elapsed_time(e1))
    kits = kplan.iters
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0)) * 1e12
    a = kits[:, 1:].abs().float()
    rounds = -(-B // (128 * 148))
    rpc = min(128, max(32, -(-(-(-B // (rounds * 148))) // 32) * 32))      # one row per rod (csrc/kc_knode_tc.cu: tc_fwd1_rpc)
    joint = float(a.view(-1, rpc, a.shape[1]).amax(1).sum().item())
    k_exec = joint * (N - 1) * 2 * 128 * 32 * 512 * 2 * 3.0
    del kplan
    torch.cuda.empty_cache()
    return ({
            "metric": "rod-node-steps/sec", "value": world * rns / (ms * 1e-3), "unit": "rod-node-steps/s", "n_gpus": world,
            "steps": steps, "warmup": 2, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 large sweep (BASELINE configs[4]): 8192 rods/GPU x 20 nodes x 500 time indices (65 536 "
                                   "rods on 8 GPUs), class-default rod of train_segment.py, physics-only, half sine / half "
                                   "random tensions", "rods_per_gpu": B, "nodes": N, "time_indices": T,
                       "l2": "8.2 GB of trajectory written per step (>> 126 MB L2)",
                       "parallelism": f"rods sharded over {world} rank(s), no collective",
                       "solver": {"marches_per_step_mean": marches, "all_converged": ok}},
            "clocks": sampler.summary(), "gpu_launches": 2 * steps,
            "roofline": {"bound": "fp32", "kernel": "kc_rollout_kernel<float,diag,physics> (one rod per lane, Broyden) + layout "
                         "transpose", "achieved": rns * flop_rns / (ms * 1e-3) / 1e12, "peak": fp32_peak / 1e12,
                         "unit": "TFLOP/s", "frac": rns * flop_rns / (ms * 1e-3) / fp32_peak,
                         "peak_source": "kc_fma_peak micro-benchmark measured live in this run",
                         "normalisation": "6.40 kFLOP per rod-node-step = 15 nominal evaluations x 19/20 x 449 FLOP (SURVEY 8d)",
                         "traffic": None},
            "e2e": None, "cpu_baseline": None,
            "knode": {"metric": "KNODE rod-node-steps/sec", "value": world * rns / (kms * 1e-3), "unit": "rod-node-steps/s",
                      "ms_per_rollout": kms, "hidden": 512, "all_converged": bool(int(kits.min()) >= 0),
                      "marches_per_step_mean": float(a.mean().item()),
                      "roofline": {"bound": "tensor", "kernel": "kc_knode_tc_fwd1_kernel (one row per rod, %d rods per CTA)" % rpc,
                                   "achieved": k_exec / (kms * 1e-3) / 1e12,
                                   "peak": bf16_peak / 1e12, "unit": "TFLOP/s", "frac": k_exec / (kms * 1e-3) / bf16_peak,
                                   "what": "executed tensor FLOP/s (128-row tiles, 3 bf16 hi/lo passes)",
                                   "useful_tflops": rns * E_REF * (N - 1) / N * (F_ODE + 106 * 512) / (kms * 1e-3) / 1e12}}})


if __name__ == "__main__":
    main()
