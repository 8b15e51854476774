"""Timing of the KNODE rollout / BPTT kernels (tensor-core march vs the SIMT warp-cooperative kernels).
usage: python tools/time_knode.py [B ...]   (env KC_TIME_SIMT=1 also times the SIMT kernels)"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "knode-cosserat_b200")
import _kc, _ops
from cosserat_ode_torch import CosseratRodTorch
from knode import setup_robot
from physics_controls import synthetic_tensions

def time_it(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    Bs = [int(a) for a in sys.argv[1:]] or [1024]
    T, H = int(os.environ.get("KC_TIME_T", 30)), int(os.environ.get("KC_TIME_H", 512))
    torch.manual_seed(1)
    robot = CosseratRodTorch("cuda", H)
    setup_robot(robot, "youngs")
    with torch.no_grad():
        robot.nn_models[2].weight.mul_(0.02); robot.nn_models[2].bias.mul_(0.02)
    P = robot._params()
    sd = robot.nn_models.state_dict()
    mlp = _ops.Mlp(sd["0.weight"], sd["0.bias"], sd["2.weight"], sd["2.bias"])
    out = []
    for B in Bs:
        tens = torch.tensor(synthetic_tensions(B, T, robot.del_t, seed=0), device="cuda", dtype=torch.float32)
        for mode in (["1", "0"] if os.environ.get("KC_TIME_SIMT") else ["1"]):
            os.environ["KC_ROLLOUT_TC"] = mode
            plan = _ops.RolloutPlan(P, mlp, B, T, torch.float32, tens.device, rows=25)
            fwd = time_it(lambda: plan.run(tens))
            its = plan.iters
            traj = plan.traj.clone()
            g = torch.randn_like(traj)
            bwd = time_it(lambda: _ops.rollout_bwd(P, mlp, tens, traj, g), n=3, warm=1)
            r = {"B": B, "T": T, "H": H, "tc": mode, "fwd_ms": fwd, "bwd_ms": bwd, "marches_per_step": float(its[:, 1:].abs().float().mean()),
                 "converged": bool(int(its.min()) >= 0), "rod_node_steps_per_s_fwd": B * int(P.N) * (T - 1) / (fwd * 1e-3)}
            print(json.dumps(r)); out.append(r)
    return out

if __name__ == "__main__":
    main()
