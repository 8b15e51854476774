"""RK4-march rollout (kc_rollout_fwd_rk4) next to the Euler-march rollout on the same inputs: class-default rod (the
configuration the reference's RK4 residual is stable for), fp32 and fp64.  python tools/rk4_rollout.py [B] [T]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from physics_controls import synthetic_tensions
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
robot = CosseratRod(use_fsolve=True)
P = _kc.rod_params(robot)
N = robot.N
for dt in (torch.float32, torch.float64):
    ctl = torch.tensor(synthetic_tensions(B, T, robot.del_t, seed=0, dtype=np.float64), device="cuda", dtype=dt)
    for method in ("euler", "rk4"):
        plan = _ops.RolloutPlan(P, None, B, T, dt, torch.device("cuda", 0), rows=25, method=method)
        for _ in range(2): plan.run(ctl)
        torch.cuda.synchronize()
        its = plan.iters.cpu().numpy()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.run(ctl); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        evals = 4 if method == "rk4" else 1
        print("%s %-5s B %d T %d: %.2f ms  %.3e rod-node-steps/s  marches/step %.2f  converged %s  (%d ODE evaluations per node per march)"
              % (str(dt).split(".")[-1], method, B, T, ms, B * N * (T - 1) / ms * 1e3, float(np.abs(its[:, 1:]).mean()),
                 bool(its.min() >= 0), evals))
