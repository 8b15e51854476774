"""One rollout launch for ncu: python tools/prof_rollout.py [B] [T] [f32|f64] [reps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
P = _robot
rng = np.random.default_rng(0)
ctl = np.empty((B, T, 4), np.float32)
i = np.arange(1, T + 1)[None, :, None]
per = rng.uniform(0.5, 3.0, (B // 2, 1, 1)) / P.del_t
ph = rng.uniform(0, 2 * np.pi, (B // 2, 1, 1))
ctl[:B // 2] = 6 + np.sin(2 * np.pi * i / per + ph + np.arange(4)[None, None, :] * np.pi / 2)
ctl[B // 2:] = 5 + 5 * rng.random((B - B // 2, T, 4))
ctl = torch.tensor(ctl, dtype=dt, device="cuda")
plan = _ops.RolloutPlan(_kc.rod_params(P), None, B, T, dt, "cuda")
for _ in range(reps):
    plan.run(ctl)
torch.cuda.synchronize()
it = plan.iters.cpu().numpy()
print("ok marches mean %.2f; per-warp max mean %.2f" % (np.abs(it[:, 1:]).mean(), np.abs(it[:, 1:]).reshape(B // 32, 32, T - 1).max(1).mean()))
