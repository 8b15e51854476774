"""ODE_parallel forward / backward at the C3 sample count: python tools/prof_ode.py [Q] [H]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "knode-cosserat_b200"))
import numpy as np, torch
import _kc, _ops
from cosserat_ode import CosseratRod
from knode import setup_robot
_robot = CosseratRod(use_fsolve=True); setup_robot(_robot)
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 118784
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
P = _kc.rod_params(_robot)
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
mlp = _ops.Mlp(r(H, 28) * 0.1, r(H) * 0.1, r(25, H) * 0.02, r(25) * 0.01)
y = r(Q, 19); y[:, 3] += 2.0; yh = r(Q, 19); zh = r(Q, 6); tf = r(Q, 3)
def timeit(f, n=5):
    for _ in range(2): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print("Q %d H %d" % (Q, H))
print("ode_fwd physics  %.3f ms" % timeit(lambda: _ops.ode_fwd(P, None, y, yh, zh, tf)))
print("ode_fwd knode    %.3f ms" % timeit(lambda: _ops.ode_fwd(P, mlp, y, yh, zh, tf)))
gys, gz = r(Q, 19), r(Q, 6)
print("ode_bwd knode    %.3f ms" % timeit(lambda: _ops.ode_bwd(P, mlp, y, yh, zh, tf, gys, gz)))
x = r(Q, 28); go = r(Q, 25)
print("mlp_fwd          %.3f ms" % timeit(lambda: _ops.mlp_fwd(mlp, x)))
print("mlp_bwd          %.3f ms" % timeit(lambda: _ops.mlp_bwd(mlp, x, go)))
