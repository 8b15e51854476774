import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/knode-cosserat_b200")
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
from physics_controls import synthetic_tensions
P = _kc.rod_params(_robot)
for B in (1, 64, 1024, 1776, 4096):
    T = 100
    ctl = torch.tensor(synthetic_tensions(B, T, 0.05, seed=0, dtype=np.float64), device="cuda")
    for lin in ("0", None):
        if lin is None: os.environ.pop("KC_ROLLOUT_LIN", None)
        else: os.environ["KC_ROLLOUT_LIN"] = lin
        plan = _ops.RolloutPlan(P, None, B, T, torch.float64, "cuda", rows=50)
        for _ in range(2): plan.run(ctl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): plan.run(ctl)
        e1.record(); e1.synchronize()
        it = plan.iters.cpu().numpy()
        print("B %5d lin=%s: %.3f ms  marches %.2f  ok %s" % (B, lin, e0.elapsed_time(e1) / 5, np.abs(it[:, 1:]).mean(), bool((it >= 0).all())))
