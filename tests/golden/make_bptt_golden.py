#!/usr/bin/env python
"""Golden gradients THROUGH a KNODE rollout, composed only of the unmodified reference's pieces.

The reference never differentiates a rollout, so there is nothing to call directly.  This script builds the rollout of
knode.simulate (knode.py:55-102: BDF2 history :74-77, warm-started shooting solve :88-89, Euler march of
cosserat_ode.py:188-213) out of CosseratRodTorch.ODE_parallel (cosserat_ode_torch.py:217-322, the reference's only fully
differentiable node function; `ODE` detaches R and hs_mat, :158-161,185-188) under torch.set_default_dtype(float64) and
lets torch.autograd differentiate it.  The shooting solve is differentiated by the implicit function theorem, written as
one extra Newton step from the converged root with a detached Jacobian: G = G* - J(G*)^-1 F(G*; theta) has the value G*
(F(G*) ~ 1e-14) and the derivative -J^-1 dF/dtheta.

    python tests/golden/make_bptt_golden.py      ->  tests/golden/bptt.npz

loss = sum(Cw * traj[B,T,25,N]) with a fixed random Cw; gradients w.r.t. W1, b1, W2, b2 and the tensions.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/knode_cosserat"
sys.path.insert(0, REF)
torch.set_default_dtype(torch.float64)

from cosserat_ode_torch import CosseratRodTorch  # noqa: E402
from knode import setup_robot  # noqa: E402
from physics_controls import calc_controls  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def march(robot, G, y, z, yh, zh, tf):
    """cosserat_ode.py:188-213 for a batch: base node [p0,h0,G,q0,w0], Euler march with ODE_parallel; z[:, :, N-1] keeps its
    previous value.  Returns residual [B,6], y_new [B,19,N], z_new [B,6,N]."""
    B, N = G.shape[0], robot.N
    rep = lambda v: v.reshape(1, -1).repeat(B, 1)
    yj = torch.cat([rep(robot.p0), rep(robot.h0), G, rep(robot.q0), rep(robot.w0)], dim=1)
    ys, zs = [yj], []
    for j in range(N - 1):
        dy, zj = robot.ODE_parallel(yj, yh[:, :, j], zh[:, :, j], tf)
        zs.append(zj)
        yj = yj + robot.ds * dy
        ys.append(yj)
    zs.append(z[:, :, N - 1])
    y_new, z_new = torch.stack(ys, dim=2), torch.stack(zs, dim=2)
    res = torch.cat([rep(robot.F_tip) - y_new[:, 7:10, -1], rep(robot.M_tip) - y_new[:, 10:13, -1]], dim=1)
    return res, y_new, z_new


def rollout(robot, ctl):
    """knode.py:55-102 for a batch, differentiable w.r.t. the MLP parameters and ctl.  -> traj [B,T,25,N]"""
    B, T, _ = ctl.shape
    N = robot.N
    y = torch.zeros(B, 19, N)
    y[:, 2, :] = torch.linspace(0, float(robot.L), N)
    y[:, 3, :] = 1.0
    z = torch.zeros(B, 6, N)
    z[:, 2, :] = 1.0
    y_prev, z_prev = y.clone(), z.clone()
    G = torch.zeros(B, 6)
    traj = [torch.cat([y, z], dim=1)]
    for t in range(T - 1):
        yh = robot.c1 * y + robot.c2 * y_prev
        zh = robot.c1 * z + robot.c2 * z_prev
        y_prev, z_prev = y, z
        tf = ctl[:, t] @ robot.tendon_dirs
        with torch.no_grad():   # the root itself: Newton on the detached problem, central-difference Jacobian
            for _ in range(30):
                F, _, _ = march(robot, G, y, z, yh, zh, tf)
                J = torch.zeros(B, 6, 6)
                for k in range(6):
                    e = 1e-6 * torch.clamp(G[:, k].abs(), min=1.0)
                    Gp, Gm = G.clone(), G.clone()
                    Gp[:, k] += e
                    Gm[:, k] -= e
                    J[:, :, k] = (march(robot, Gp, y, z, yh, zh, tf)[0] - march(robot, Gm, y, z, yh, zh, tf)[0]) / (2 * e)[:, None]
                dG = torch.linalg.solve(J, -F.unsqueeze(2)).squeeze(2)
                G = G + dG
                if float(dG.abs().max()) < 1e-13 * max(1.0, float(G.abs().max())):
                    break
            # exact Jacobian at the root by autograd (for the implicit-function derivative)
        with torch.enable_grad():
            Gd = G.detach().clone().requires_grad_(True)
            Fd, _, _ = march(robot, Gd, y.detach(), z.detach(), yh.detach(), zh.detach(), tf.detach())
            Jx = torch.stack([torch.autograd.grad(Fd[:, i].sum(), Gd, retain_graph=True)[0] for i in range(6)], dim=1)
        F, _, _ = march(robot, G.detach(), y, z, yh, zh, tf)                   # depends on theta, ctl, history
        Gi = G.detach() - torch.linalg.solve(Jx.detach(), F.unsqueeze(2)).squeeze(2)
        _, y, z = march(robot, Gi, y, z, yh, zh, tf)
        G = Gi.detach()
        traj.append(torch.cat([y, z], dim=1))
    return torch.stack(traj, dim=1)


def main():
    out = {}
    rng = np.random.default_rng(21)
    for tag, H, B, T in [("h16", 16, 3, 7), ("h48", 48, 2, 5)]:
        torch.manual_seed(5)
        robot = CosseratRodTorch("cpu", H)
        setup_robot(robot, "youngs")
        with torch.no_grad():
            robot.nn_models[2].weight.mul_(0.3)
            robot.nn_models[2].bias.mul_(0.3)
        ctl = np.stack([np.array(calc_controls("sine", 0.7 + 0.2 * b, robot.del_t, T)) if b % 2 == 0
                        else 5 + 5 * rng.random((T, 4)) for b in range(B)])
        ctl_t = torch.tensor(ctl, requires_grad=True)
        Cw = rng.standard_normal((B, T, 25, robot.N))
        traj = rollout(robot, ctl_t)
        loss = (torch.tensor(Cw) * traj).sum()
        params = [robot.nn_models[0].weight, robot.nn_models[0].bias, robot.nn_models[2].weight, robot.nn_models[2].bias]
        grads = torch.autograd.grad(loss, params + [ctl_t])
        out[tag + "_ctl"], out[tag + "_Cw"], out[tag + "_traj"] = ctl, Cw, traj.detach().numpy()
        for k, p, g in zip(("W1", "b1", "W2", "b2"), params, grads[:4]):
            out[f"{tag}_{k}"] = p.detach().numpy().copy()
            out[f"{tag}_g{k}"] = g.numpy().copy()
        out[tag + "_gctl"] = grads[4].numpy().copy()
        print(tag, "loss", float(loss), {k: float(np.abs(out[f'{tag}_g{k}']).max()) for k in ("W1", "b1", "W2", "b2", "ctl")})
    np.savez_compressed(os.path.join(OUT, "bptt.npz"), **out)


if __name__ == "__main__":
    main()
