"""A few teacher-forced training steps at the C3 size for ncu: python tools/prof_train.py [B] [reps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from cosserat_ode_torch import CosseratRodTorch
from knode import setup_robot
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
robot = CosseratRod(use_fsolve=True); setup_robot(robot)
ctl = torch.tensor(synthetic_tensions(B, 30, robot.del_t, seed=0), device="cuda")
traj, _, _ = _ops.rollout(_kc.rod_params(robot), None, ctl)
torch.manual_seed(0)
tr = CosseratRodTorch("cuda", 512); setup_robot(tr)
for _ in range(reps):
    loss, grads, _ = tr.teacher_forced_step(traj, ctl, [3, 5, 7, 9])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    loss, grads, _ = tr.teacher_forced_step(traj, ctl, [3, 5, 7, 9])
e1.record(); e1.synchronize()
print("ok loss %.6f  %.3f ms/step (kc_train_step only)" % (float(loss.item()), e0.elapsed_time(e1) / reps))
