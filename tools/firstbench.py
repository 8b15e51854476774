"""Scratch timing of the rollout kernel (first GPU visit). Not the bench contract."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
P = _robot
pc = _kc.rod_params(P)
def make_ctl(B, T, seed=0):
    rng = np.random.default_rng(seed)
    ctl = np.empty((B, T, 4), np.float32)
    i = np.arange(1, T + 1)[None, :, None]
    per = rng.uniform(0.5, 3.0, (B // 2, 1, 1)) / P.del_t
    ph = rng.uniform(0, 2 * np.pi, (B // 2, 1, 1))
    ctl[:B // 2] = 6 + np.sin(2 * np.pi * i / per + ph + np.arange(4)[None, None, :] * np.pi / 2)
    ctl[B // 2:] = 5 + 5 * rng.random((B - B // 2, T, 4))
    return ctl
print("fma peak fp32 %.1f TF  fp64 %.1f TF" % (_ops.fma_peak(torch.float32, 20000, "cuda") / 1e12, _ops.fma_peak(torch.float64, 20000, "cuda") / 1e12))
for dt in (torch.float32, torch.float64):
    for B in (4096, 16384, 65536):
        T = 100
        ctl = torch.tensor(make_ctl(B, T), dtype=dt, device="cuda")
        plan = _ops.RolloutPlan(pc, None, B, T, dt, "cuda")
        for _ in range(2): plan.run(ctl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.run(ctl)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 3
        it = plan.iters.cpu().numpy()
        print(f"{dt} B={B} T={T}: {ms:.3f} ms  {B*10*T/ms*1e3:.3e} rod-node-steps/s  marches mean {np.abs(it[:,1:]).mean():.2f} max {np.abs(it).max()} fails {(it<0).sum()}")
