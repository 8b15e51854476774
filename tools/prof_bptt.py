"""One KNODE rollout + BPTT step at the C3 size for ncu: python tools/prof_bptt.py [B] [T] [H]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode_torch import CosseratRodTorch
from knode import setup_robot
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
torch.manual_seed(1)
robot = CosseratRodTorch("cuda", H); setup_robot(robot, "youngs")
with torch.no_grad():
    robot.nn_models[2].weight.mul_(0.02); robot.nn_models[2].bias.mul_(0.02)
ctl = torch.tensor(synthetic_tensions(B, T, robot.del_t, seed=0), device="cuda")
P = robot._params(); mlp = robot._mlp()
for rep in range(int(os.environ.get('REPS', '2'))):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    traj, _, iters = _ops.rollout(P, mlp, ctl, rows=25)
    e[1].record()
    g = torch.ones_like(traj) * 1e-3
    out = _ops.rollout_bwd(P, mlp, ctl, traj, g)
    e[2].record(); torch.cuda.synchronize()
    if os.environ.get('VERBOSE'): print('rep %d fwd %.2f bwd %.2f' % (rep, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
it = iters[:, 1:].abs().float()
print("ok fwd %.2f ms  bwd %.2f ms  converged %s  |gW1| %.3e  marches/step mean %.2f max %d" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), bool((iters >= 0).all()), float(out[1].abs().max()), float(it.mean()), int(it.max())))
