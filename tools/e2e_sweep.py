"""Time kc_rollout_host (C2 shape) for several segment counts: python tools/e2e_sweep.py"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
from physics_controls import synthetic_tensions
B, T = 4096, 100
P = _kc.rod_params(_robot)
ctl = torch.tensor(synthetic_tensions(B, T, 0.05, seed=0, dtype=np.float32)).pin_memory()
out = torch.empty((B, T, 25, 10), dtype=torch.float32).pin_memory()
dev = torch.device("cuda", 0)
d = torch.empty((B, T, 25, 10), dtype=torch.float32, device=dev)
for _ in range(2):
    out.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out.copy_(d, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("plain D2H 410 MB: %.2f ms  %.1f GB/s" % (dt * 1e3, out.numel() * 4 / dt / 1e9))
for seg in (1, 2, 4, 6, 8):
    hp = _ops.HostRolloutPlan(P, None, B, T, torch.float32, dev, rows=25, segments=seg)
    for _ in range(2):
        hp.run(ctl, out=out)
    t0 = time.perf_counter()
    for _ in range(5):
        hp.run(ctl, out=out)
    dt = (time.perf_counter() - t0) / 5
    print("segments %d: %.2f ms/step  %.3g rod-node-steps/s" % (seg, dt * 1e3, B * 10 * (T - 1) / dt))
