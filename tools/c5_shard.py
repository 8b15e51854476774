"""BASELINE config 5, one GPU's shard (8192 rods x 20 nodes x 500 time indices, class-default parameters, fp32): kernel
time with CUDA events, rod-node-steps/s and nominal-FLOP roofline fraction.  python tools/c5_shard.py [B] [N] [T]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
robot = CosseratRod(use_fsolve=True); robot.N = N; robot.compute_intermediate_terms()
P = _kc.rod_params(robot)
ctl = torch.tensor(synthetic_tensions(B, T, robot.del_t, seed=0, dtype=np.float32), device="cuda")
plan = _ops.RolloutPlan(P, None, B, T, torch.float32, torch.device("cuda", 0), rows=25)
for _ in range(2): plan.run(ctl)
torch.cuda.synchronize()
its = plan.iters.cpu().numpy()
assert its.min() >= 0
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run(ctl); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
rns = B * N * (T - 1)
flop = 15.0 * (N - 1) / N * 449.0
print("C5 shard B %d N %d T %d fp32: %.2f ms  %.3e rod-node-steps/s  %.1f TFLOP/s nominal  marches/step %.2f  output %.1f GB -> %.0f GB/s"
      % (B, N, T, ms, rns / ms * 1e3, rns * flop / ms / 1e9, float(np.abs(its[:, 1:]).mean()), plan.traj.numel() * 4 / 1e9,
         plan.traj.numel() * 4 / ms / 1e6))
