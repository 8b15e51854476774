"""Rollout kernel time vs batch size (latency- or issue-bound?): python tools/b_sweep.py [B ...]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
from physics_controls import synthetic_tensions
T = 100
P = _kc.rod_params(_robot)
for B in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]:
    ctl = torch.tensor(synthetic_tensions(B, T, 0.05, seed=0, dtype=np.float32), device="cuda")
    plan = _ops.RolloutPlan(P, None, B, T, torch.float32, "cuda", rows=25)
    for _ in range(3):
        plan.run(ctl)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); plan.run(ctl); b.record()
    torch.cuda.synchronize()
    ms = np.mean([a.elapsed_time(b) for a, b in ev])
    it = np.abs(plan.iters.cpu().numpy()[:, 1:])
    print("B %6d: %.3f ms  %.3g rod-node-steps/s  marches mean %.2f" % (B, ms, B * 10 * (T - 1) / ms * 1e3, it.mean()))
