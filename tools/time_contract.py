"""dev: knode.simulate through the reference's own contract (fp64 in, fresh fp64 [B,T,50,N] ndarray out): python tools/time_contract.py [B]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
from cosserat_ode import CosseratRod
from knode import setup_robot, simulate
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
robot = CosseratRod(use_fsolve=True); setup_robot(robot)
ctl = synthetic_tensions(B, 100, robot.del_t, seed=0).astype(np.float64)
out = simulate(robot, ctl)
ref = simulate(robot, torch.tensor(ctl, device="cuda"))      # device path -> .cpu(): the plain copy
print("shape", out.shape, out.dtype, "identical to the device path:", bool(np.array_equal(out, ref)))
del ref
for _ in range(3):
    t0 = time.perf_counter(); out = simulate(robot, ctl); t1 = time.perf_counter()
    print("%.1f ms  (%.3g rod-node-steps/s)" % ((t1 - t0) * 1e3, B * 10 * 99 / (t1 - t0)))
    del out
