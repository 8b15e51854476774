"""setup_robot / simulate — drop-in for knode_cosserat/knode.py, with the time rollout on the GPU.

`simulate(robot, ctl)` keeps the reference signature and return layout (float64 [T,50,N]; index 0 = initial state, the
step driven by the last control is dropped, rows 25:50 = yh, zh — knode.py:55-102) but runs ONE launch of the batched
rollout kernel instead of T calls of scipy.optimize.fsolve around a Python march; a 3-D `ctl` [B,T,4] rolls out B
independent rods at once and returns [B,T,50,N].  The shooting solve is a quasi-Newton iteration converged to 1e-11
(fp64), i.e. tighter than the reference's hybrd default (xtol 1.49e-8); `use_fsolve` is accepted and irrelevant.
"""
import numpy as np
import torch

import _kc
import _ops
from cosserat_ode import CosseratRod


def setup_robot(robot, mod=None, original=False):
    """Set up robot based on experimental parameters (knode.py:6-53)."""
    if original:
        raise Exception("--original parameter no longer supported")
    # Measured on the robot (:11-20)
    robot.del_t = 0.05
    robot.L = 0.635
    robot.tendon_offset = 0.04445
    robot.r = 0.003175
    robot.rho = 1411.6751
    robot.E = 2.757903e9
    Bbt = 3e-2
    if mod is None:
        pass
    elif mod == 'noair':
        if isinstance(robot, CosseratRod):
            robot.C = np.array([0, 0, 0])
        else:
            robot.C = torch.tensor([0, 0, 0], device=robot.device)
    elif mod == 'nsw':
        if isinstance(robot, CosseratRod):
            robot.g = np.array([0, 0, 0])
        else:
            robot.g = torch.tensor([0, 0, 0], device=robot.device)
    elif mod == 'short':
        robot.L = 0.4
    elif mod == 'damping':
        Bbt = 0.2
    elif mod == 'dampstiff':
        Bbt = 0.2
        robot.E = 10e9
    elif mod == 'lengthstiff':
        robot.L = 0.4
        robot.E = 10e9
    elif mod == 'youngs':
        robot.E = 10e9
    else:
        raise Exception('Unknown mod ' + mod)
    if isinstance(robot, CosseratRod):
        robot.Bbt = np.diag([Bbt, Bbt, Bbt])
    else:
        robot.Bbt = torch.diag(torch.tensor([Bbt, Bbt, Bbt], device=robot.device))
    robot.compute_intermediate_terms()


_PLANS = {}   # output/workspace buffers reused across calls of the same shape (they are overwritten by every call)


def _initial_state(robot_reference, B):
    """Straight rod (knode.py:58-64), float64 host arrays [B,19,N], [B,6,N]."""
    N = robot_reference.N
    y = np.vstack([np.zeros((2, N)), np.linspace(0, robot_reference.L, N), np.ones((1, N)), np.zeros((15, N))])
    z = np.vstack([np.zeros((2, N)), np.ones((1, N)), np.zeros((3, N))])
    return np.broadcast_to(y, (B, 19, N)).copy(), np.broadcast_to(z, (B, 6, N)).copy()


def _robot_mlp(robot, dtype):
    if isinstance(robot, CosseratRod):
        return robot._mlp(dtype)
    return robot._mlp()  # CosseratRodTorch


def simulate(robot, ctl, robot_reference=None, *, dtype=np.float64, rows=50, tol=0.0, max_iter=0, return_info=False,
             pinned_out=None, method="euler", device_out=False, select=None):
    """Roll the rod out under the tendon tensions `ctl` ([T,4] -> [T,rows,N]; [B,T,4] -> [B,T,rows,N]).

    Keyword-only extensions (not in the reference): dtype (np.float64 | np.float32 arithmetic and output), rows (50 =
    reference layout, 25 = [y;z] only), tol / max_iter of the shooting solve, return_info -> (traj, G[..,T,6],
    marches[..,T]), pinned_out = a pinned host torch tensor to receive the result (avoids an allocation per call),
    method = "euler" (getResidualEuler, what the reference's loop calls, knode.py:89) or "rk4" (the reference's
    getResidualRK4, cosserat_ode.py:215-255, as the residual of the same loop), device_out = True with CUDA tensions:
    return the trajectory as a CUDA tensor (no device-to-host copy; for callers that reduce it on the GPU, e.g. the
    evaluation metrics of physics_train / physics_multitrain).

    select = (row_indices, node_indices): return only traj[..., rows, :][..., nodes] — e.g. ([0, 1, 2], [N-1]) for the tip
    positions that physics_train.py:159 / physics_multitrain.py:211 actually read.  The rollout is unchanged; the selection
    happens on the device, so the device-to-host copy (the bulk of an end-to-end call: 410 MB at BASELINE config 2) shrinks
    to the selected entries.

    Shooting solve: Newton / Broyden on the 6 base reactions to `tol` (default 1e-11 fp64, 2e-6 fp32, relative to
    max(1, |G|)) instead of MINPACK hybrd.  In the kernels with the linearised final correction (small batches) the LAST
    Newton step of a time step is not re-marched: once the step dG is so small that its estimated second-order remainder
    4 C |dG|^2 is below tol, the state at G + dG is formed from the finite-difference marches of the same iteration.  Such
    steps are reported in `marches` like any converged step; KC_ROLLOUT_LIN=0 in the environment switches the correction
    off (every accepted state is then a marched one) — the parity tests run both.
    """
    if robot_reference is None:
        robot_reference = robot
    if not torch.cuda.is_available():
        raise RuntimeError("knode-cosserat_b200 has no CPU fallback: simulate() needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device())
    if select is not None:
        # solve on the device, gather the selected rows / nodes there, copy only those back
        rows_sel, nodes_sel = select
        if torch.is_tensor(ctl):
            ctl_dev = ctl.to(dev, non_blocking=True)
        else:
            ctl_np_ = np.asarray(ctl)
            ctl_dev = torch.as_tensor(ctl_np_ if ctl_np_.dtype.kind == 'f' else ctl_np_.astype(np.float64)).to(dev, non_blocking=True)
        full = simulate(robot, ctl_dev, robot_reference, dtype=dtype, rows=rows, tol=tol, max_iter=max_iter, method=method,
                        device_out=True)
        ri = torch.as_tensor(np.asarray(rows_sel).reshape(-1), device=dev, dtype=torch.long)
        ni = torch.as_tensor(np.asarray(nodes_sel).reshape(-1), device=dev, dtype=torch.long)
        sel = full.index_select(-2, ri).index_select(-1, ni)
        if device_out:
            return sel
        if pinned_out is not None:
            pinned_out.copy_(sel, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return pinned_out.numpy()
        return sel.cpu().numpy()
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    on_device = torch.is_tensor(ctl) and ctl.is_cuda
    if method != "euler" and not on_device:     # the pipelined host entry point runs the Euler march only
        ctl_t = ctl if torch.is_tensor(ctl) else torch.as_tensor(np.asarray(ctl, dtype=np.float64))
        out = simulate(robot, ctl_t.to(dev), robot_reference, dtype=dtype, rows=rows, tol=tol, max_iter=max_iter,
                       return_info=return_info, method=method)
        if pinned_out is not None and not return_info:
            pinned_out.copy_(torch.from_numpy(out))
        return out
    if on_device:
        tens = ctl.to(dev, tdt)
    else:
        ctl_np = ctl.detach().numpy() if torch.is_tensor(ctl) else np.asarray(ctl)
        if ctl_np.dtype.kind != 'f':
            ctl_np = ctl_np.astype(np.float64)                 # lists / ints go through float64 like knode.py:71
        # the cast to the arithmetic type happens while staging into pinned memory (HostRolloutPlan.run), in one pass
        tens = torch.from_numpy(np.ascontiguousarray(ctl_np))
    single = tens.ndim == 2
    if single:
        tens = tens[None]
    B, T, _ = tens.shape
    # the reference leaves the last applied tensions on the robot (knode.py:71)
    if T > 0 and B > 0:
        last = ctl[-1] if single else ctl[-1][-1]
        robot.tendon_tensions = np.asarray(last.detach().cpu().numpy() if torch.is_tensor(last) else last,
                                           dtype=np.float64).copy()
    P = _kc.rod_params(robot)
    mlp = _robot_mlp(robot, tdt)
    # initial state: the kernel builds the straight rod of knode.py:58-64 itself; an explicit y0/z0 is only needed when
    # another robot defines the geometry (robot_reference) or in fp64, where linspace's exact end point is reproduced
    y0 = z0 = None
    if robot_reference is not robot or tdt == torch.float64:
        y0n, z0n = _initial_state(robot_reference, B)
        y0, z0 = torch.as_tensor(y0n).to(dev, tdt), torch.as_tensor(z0n).to(dev, tdt)
    key = (B, T, rows, tdt, dev, int(P.N), None if mlp is None else (mlp.in_dim, mlp.hidden), bool(return_info),
           on_device, method)
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= 4:
            _PLANS.clear()
        if on_device:
            plan = _ops.RolloutPlan(P, mlp, B, T, tdt, dev, rows, want_G=return_info, method=method)
        else:
            plan = _ops.HostRolloutPlan(P, mlp, B, T, tdt, dev, rows, want_G=return_info)
        _PLANS[key] = plan
    plan.rebind(P, mlp)
    if on_device:
        plan.run(tens, y0, z0, tol=tol, max_iter=max_iter)
        traj, G, iters = plan.traj, plan.G, plan.iters
        if device_out:
            out_t = traj.clone()
            if single:
                out_t = out_t[0]
            if return_info:
                return (out_t, G[0].clone(), iters[0].clone()) if single else (out_t, G.clone(), iters.clone())
            return out_t
        if pinned_out is not None:
            pinned_out.copy_(traj, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            out = pinned_out.numpy()
        else:
            out = traj.cpu().numpy()   # a fresh host array: the device buffers of the cached plan are reused by later calls
        if return_info:
            Gn, itn = G.cpu().numpy(), iters.cpu().numpy()
    else:
        # host data in, host data out: ONE C-ABI call (kc_rollout_host) that stages the tensions, solves the rollout in
        # time ranges and copies each finished range back while the next one is computed
        dst = pinned_out if pinned_out is not None else torch.empty((B, T, rows, int(P.N)), dtype=tdt)
        out = plan.run(tens, y0, z0, tol=tol, max_iter=max_iter, out=dst).numpy()
        if return_info:
            Gn, itn = plan.G_h.numpy().copy(), plan.iters_h.numpy().copy()
    if single:
        out = out[0]
    if return_info:
        return (out, Gn[0], itn[0]) if single else (out, Gn, itn)
    return out
