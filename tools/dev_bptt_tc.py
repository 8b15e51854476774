"""dev: tensor-core BPTT vs fp64 / SIMT fp32 kernels, per-gradient and per-step errors"""
import os, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "knode-cosserat_b200"); sys.path.insert(0, "tests")
from oracle import rod_oracle as O
import _kc, _ops
from test_gpu_knode_tc import _case, _dev, _params, PK

def grads(P, ctl, mlp, Cw, dt, tcmode):
    os.environ["KC_ROLLOUT_TC"] = tcmode
    m = _ops.Mlp(*[_dev(mlp[k], dt) for k in PK])
    traj, _, iters = _ops.rollout(_params(P), m, _dev(ctl, dt), rows=25, tol=1e-13 if dt == torch.float64 else 0.0)
    out = _ops.rollout_bwd(_params(P), m, _dev(ctl, dt), traj, _dev(Cw, dt))
    torch.cuda.synchronize()
    return [g.cpu().numpy().astype(np.float64) for g in out]

for (H, B, T) in [(64, 5, 12), (64, 1, 3), (512, 21, 7)]:
    P, ctl, mlp = _case(H, B, T, seed=3 * H + B)
    Cw = np.random.default_rng(H).standard_normal((B, T, 25, P.N))
    g64 = grads(P, ctl, mlp, Cw, torch.float64, "0")
    gs = grads(P, ctl, mlp, Cw, torch.float32, "0")
    gt = grads(P, ctl, mlp, Cw, torch.float32, "1")
    print("case", H, B, T)
    for name, a, s, b in zip(("tensions",) + PK, gt, gs, g64):
        print(f"  {name:9s} tc err {np.max(np.abs(a-b))/np.abs(b).max():.3e}   simt err {np.max(np.abs(s-b))/np.abs(b).max():.3e}  scale {np.abs(b).max():.3e}")
    e = np.abs(gt[0] - g64[0]).max(axis=(0, 2)) / np.abs(g64[0]).max()
    print("  tension err per step:", np.array2string(e, precision=2))
    print("  tc   gten[0,:,0]", np.array2string(gt[0][0, :, 0], precision=4))
    print("  f64  gten[0,:,0]", np.array2string(g64[0][0, :, 0], precision=4))
