#!/usr/bin/env python
"""Generate golden vectors by importing the UNMODIFIED reference from /root/reference.

Run once in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Every array written here is an output of the reference's own functions on seeded inputs; the
loop bodies of the reference *scripts* (which cannot be imported because they need matplotlib /
fastdtw at module scope) are replicated verbatim around the imported classes, with the
reference line ranges cited next to each block.  Nothing in this file is used at test time
except the .npz files it writes.
"""
import os
import sys

import numpy as np
import scipy.optimize
import torch
import torch.nn as nn

REF = "/root/reference/knode_cosserat"
sys.path.insert(0, REF)

import knode as ref_knode  # noqa: E402
from cosserat_ode import CosseratRod  # noqa: E402
from cosserat_ode_torch import CosseratRodTorch  # noqa: E402
from knode import setup_robot, simulate  # noqa: E402
from physics_controls import calc_controls  # noqa: E402
from Utils.transformations import quaternion_to_euler  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MODS = [None, "noair", "nsw", "short", "damping", "dampstiff", "lengthstiff", "youngs"]


def tight_fsolve(func, x0, args=()):
    return scipy.optimize.fsolve(func, x0, args=args, xtol=1e-13)


def simulate_tight(robot, ctl):
    """knode.simulate (knode.py:55-102) with fsolve's xtol tightened to 1e-13."""
    saved = ref_knode.fsolve
    ref_knode.fsolve = tight_fsolve
    try:
        return simulate(robot, ctl)
    finally:
        ref_knode.fsolve = saved


def params_of(robot):
    """Snapshot of the derived constants (compute_intermediate_terms) as float64 arrays."""
    f = lambda x: np.asarray(x.detach().cpu().numpy() if torch.is_tensor(x) else x, dtype=np.float64)
    keys = ["L", "N", "E", "r", "rho", "del_t", "A", "G", "ds", "c0", "c1", "c2", "rhoA"]
    d = {k: np.float64(getattr(robot, k)) for k in keys}
    for k in ["vstar", "g", "Bse", "Bbt", "C", "F_tip", "M_tip", "tendon_dirs", "p0", "h0", "q0", "w0",
              "J", "Kse", "Kbt", "Kse_plus_c0_Bse_inv", "Kbt_plus_c0_Bbt_inv", "Kse_vstar", "rhoAg", "rhoJ"]:
        d[k] = f(getattr(robot, k))
    return d


def transplant(np_robot, torch_robot):
    """physics_train.py:103-110 — force the numpy rod to use the torch MLP."""
    nn_model = torch_robot.nn_models
    param_ls = []
    for _, layer in nn_model.state_dict().items():
        param_ls.append(layer.detach().cpu().numpy())
    np_robot.nn_model = nn_model
    np_robot.param_ls = param_ls
    np_robot.nn_path = "whatever"


def mlp_arrays(torch_robot):
    sd = torch_robot.nn_models.state_dict()
    return {"W1": sd["0.weight"].numpy().copy(), "b1": sd["0.bias"].numpy().copy(),
            "W2": sd["2.weight"].numpy().copy(), "b2": sd["2.bias"].numpy().copy()}


# --------------------------------------------------------------------------------------------
# 1. rollouts (knode.simulate)
# --------------------------------------------------------------------------------------------
def gen_rollouts():
    out = {}
    # C1-style: class-default params (cosserat_ode.py:15-29), sine 1.0, dt 0.005
    r = CosseratRod(use_fsolve=True)
    ctl = np.array(calc_controls("sine", 1.0, 0.005, 30))
    out["default_sine_ctl"] = ctl
    out["default_sine_traj_defaulttol"] = simulate(CosseratRod(use_fsolve=True), ctl)
    out["default_sine_traj"] = simulate_tight(r, ctl)
    for k, v in params_of(r).items():
        out["default_params_" + k] = v
    # longer default-param rollout (fp32 drift case of SURVEY §4)
    ctl = np.array(calc_controls("sine", 1.0, 0.005, 200))
    out["default_sine200_ctl"] = ctl
    out["default_sine200_traj"] = simulate_tight(CosseratRod(use_fsolve=True), ctl)
    # setup_robot params, every control family the reference generates
    for name, (ctype, carg, T) in {"setup_sine": ("sine", 1.0, 100), "setup_step": ("step", 1.0, 60),
                                   "setup_random": ("random", 0.0, 60)}.items():
        r = CosseratRod(use_fsolve=True)
        setup_robot(r)
        ctl = np.array(calc_controls(ctype, carg, r.del_t, T))
        out[name + "_ctl"] = ctl
        out[name + "_traj"] = simulate_tight(r, ctl)
        if name == "setup_sine":
            r2 = CosseratRod(use_fsolve=True)
            setup_robot(r2)
            out[name + "_traj_defaulttol"] = simulate(r2, ctl)
            for k, v in params_of(r).items():
                out["setup_params_" + k] = v
    # every ablation mod (knode.py:22-47), short rollouts
    for mod in MODS[1:]:
        r = CosseratRod(use_fsolve=True)
        setup_robot(r, mod)
        ctl = np.array(calc_controls("sine", 0.5, r.del_t, 20))
        out[f"mod_{mod}_ctl"] = ctl
        out[f"mod_{mod}_traj"] = simulate_tight(r, ctl)
    np.savez_compressed(os.path.join(OUT, "rollouts.npz"), **out)
    print("rollouts.npz", {k: v.shape for k, v in out.items() if k.endswith("traj")})


def gen_knode_rollouts():
    """KNODE rollout: numpy rod with a transplanted torch MLP (physics_train.py:103-116,136-158)."""
    out = {}
    # The reference's random init is unstable inside a rollout for H=512 and for history inputs (fsolve stops
    # converging, values reach 1e3..NaN), so those two nets are shrunk towards a "trained, small residual" regime.
    for tag, H, hist, s1, s2 in [("h64", 64, False, 1.0, 1.0), ("h512", 512, False, 1.0, 0.02),
                                 ("h32hist", 32, True, 0.02, 0.1)]:
        torch.manual_seed(3)
        tr = CosseratRodTorch("cpu", H, nn_input_history=hist)
        setup_robot(tr, "youngs")
        with torch.no_grad():
            tr.nn_models[0].weight.mul_(s1)
            tr.nn_models[2].weight.mul_(s2)
            tr.nn_models[2].bias.mul_(s2)
        r = CosseratRod(use_fsolve=True, nn_input_history=hist)
        setup_robot(r, "youngs")
        transplant(r, tr)
        ctl = np.array(calc_controls("sine", 1.25, r.del_t, 40))
        out[tag + "_ctl"] = ctl
        out[tag + "_traj"] = simulate_tight(r, ctl)
        for k, v in mlp_arrays(tr).items():
            out[tag + "_" + k] = v
    np.savez_compressed(os.path.join(OUT, "knode_rollouts.npz"), **out)
    print("knode_rollouts.npz", {k: v.shape for k, v in out.items() if k.endswith("traj")})


# --------------------------------------------------------------------------------------------
# 2. single ODE evaluations (ODE, ODE_parallel, numpy ODE), residual march
# --------------------------------------------------------------------------------------------
def gen_ode():
    out = {}
    rng = np.random.default_rng(11)
    r = CosseratRod(use_fsolve=True)
    setup_robot(r)
    ctl = np.array(calc_controls("random", 1.0, r.del_t, 24))
    traj = simulate_tight(r, ctl)  # [24,50,10]
    # samples: every (t>=2, node) of the trajectory, lightly perturbed so that nothing is at a root
    ys, yhs, zhs, tfs = [], [], [], []
    for t in range(2, 24):
        for j in range(10):
            ys.append(traj[t, :19, j] * (1 + 0.01 * rng.standard_normal(19)) + 1e-3 * rng.standard_normal(19))
            yhs.append(traj[t, 25:44, j])
            zhs.append(traj[t, 44:50, j])
            tfs.append(np.dot(ctl[t], r.tendon_dirs))
    ys, yhs, zhs, tfs = map(np.array, (ys, yhs, zhs, tfs))
    out.update(y=ys, yh=yhs, zh=zhs, tf=tfs)
    Q = ys.shape[0]

    for mod in MODS:
        tag = "none" if mod is None else mod
        rn = CosseratRod(use_fsolve=True)
        setup_robot(rn, mod)
        o = [rn.ODE(ys[i].copy(), yhs[i].copy(), zhs[i].copy(), tfs[i].copy()) for i in range(Q)]
        out[f"np_{tag}_ys"] = np.array([a for a, _ in o])
        out[f"np_{tag}_z"] = np.array([b for _, b in o])
    rn = CosseratRod(use_fsolve=True)  # class-default params
    o = [rn.ODE(ys[i].copy(), yhs[i].copy(), zhs[i].copy(), tfs[i].copy()) for i in range(Q)]
    out["np_default_ys"] = np.array([a for a, _ in o])
    out["np_default_z"] = np.array([b for _, b in o])

    # torch paths with the MLP: ODE (fp32, scalar) and ODE_parallel (fp32 and fp64)
    for tag, H, hist in [("h512", 512, False), ("h64hist", 64, True)]:
        torch.set_default_dtype(torch.float32)
        torch.manual_seed(5)
        tr = CosseratRodTorch("cpu", H, nn_input_history=hist)
        setup_robot(tr)
        for k, v in mlp_arrays(tr).items():
            out[f"{tag}_{k}"] = v
        t32 = [torch.tensor(a, dtype=torch.float32) for a in (ys, yhs, zhs, tfs)]
        with torch.no_grad():
            a, b = tr.ODE_parallel(*t32)
            out[f"{tag}_par32_ys"], out[f"{tag}_par32_z"] = a.numpy(), b.numpy()
            o = [tr.ODE(t32[0][i], t32[1][i], t32[2][i], t32[3][i]) for i in range(40)]
            out[f"{tag}_ode32_ys"] = np.array([a.numpy() for a, _ in o])
            out[f"{tag}_ode32_z"] = np.array([b.numpy() for _, b in o])
            tr.use_nn = False
            a, b = tr.ODE_parallel(*t32)
            out[f"{tag}_par32_nonn_ys"], out[f"{tag}_par32_nonn_z"] = a.numpy(), b.numpy()
            tr.use_nn = True
        # fp64: ODE_parallel is dtype-generic
        torch.set_default_dtype(torch.float64)
        torch.manual_seed(5)
        tr64 = CosseratRodTorch("cpu", H, nn_input_history=hist)
        setup_robot(tr64)
        tr64.nn_models.load_state_dict({k: v.double() for k, v in tr.nn_models.state_dict().items()})
        t64 = [torch.tensor(a, dtype=torch.float64, requires_grad=True) for a in (ys, yhs, zhs, tfs)]
        a, b = tr64.ODE_parallel(*t64)
        out[f"{tag}_par64_ys"], out[f"{tag}_par64_z"] = a.detach().numpy(), b.detach().numpy()
        # input + parameter gradients of a fixed scalar functional (seeds the ODE adjoint tests)
        ca = torch.tensor(rng.standard_normal(a.shape))
        cb = torch.tensor(rng.standard_normal(b.shape))
        out[f"{tag}_cot_ys"], out[f"{tag}_cot_z"] = ca.numpy(), cb.numpy()
        s = (a * ca).sum() + (b * cb).sum()
        gr = torch.autograd.grad(s, t64 + list(tr64.nn_models.parameters()))
        for nm, g in zip(["y", "yh", "zh", "tf", "W1", "b1", "W2", "b2"], gr):
            out[f"{tag}_grad64_{nm}"] = g.numpy()
        torch.set_default_dtype(torch.float32)

    # numpy shooting residual + in-place march (cosserat_ode.py:188-213)
    rn = CosseratRod(use_fsolve=True)
    setup_robot(rn)
    t = 12
    y = traj[t, :19].copy()
    z = traj[t, 19:25].copy()
    yh = traj[t + 1, 25:44].copy()
    zh = traj[t + 1, 44:50].copy()
    G = traj[t + 1, 7:13, 0] * 1.05 + 0.01
    rn.tendon_tensions = ctl[t + 1].astype(np.float64)
    res = rn.getResidualEuler(G, y, z, yh, None, zh, None)
    out.update(march_G=G, march_y0=traj[t, :19], march_z0=traj[t, 19:25], march_yh=yh, march_zh=zh,
               march_tensions=ctl[t + 1], march_res=res, march_y=y, march_z=z)
    rk = CosseratRod(use_fsolve=True)
    setup_robot(rk)
    rk.tendon_tensions = ctl[t + 1].astype(np.float64)
    y = traj[t, :19].copy()
    z = traj[t, 19:25].copy()
    yh_int = 0.5 * (yh[:, :-1] + yh[:, 1:])
    zh_int = 0.5 * (zh[:, :-1] + zh[:, 1:])
    res = rk.getResidualRK4(G, y, z, yh, yh_int, zh, zh_int)
    out.update(rk4_res=res, rk4_y=y, rk4_z=z)

    # torch shooting residual (cosserat_ode_torch.py:325-367), fp32, with MLP
    torch.manual_seed(5)
    tr = CosseratRodTorch("cpu", 512)
    setup_robot(tr)
    tr.y = torch.tensor(traj[t, :19]).float()
    tr.z = torch.tensor(traj[t, 19:25]).float()
    tr.residualArgs["yh"] = torch.tensor(yh).float()
    tr.residualArgs["zh"] = torch.tensor(zh).float()
    tr.tendon_tensions = torch.tensor(ctl[t + 1]).float()
    with torch.no_grad():
        total, full = tr.getResidualEuler(torch.tensor(G).float())
    out.update(tres32_total=total.numpy(), tres32_full=full.numpy(), tres32_y=tr.y.numpy())
    np.savez_compressed(os.path.join(OUT, "ode.npz"), **out)
    print("ode.npz Q =", Q)


# --------------------------------------------------------------------------------------------
# 3. teacher-forced segment step + training step (physics_train.py fast + slow, train_segment.py)
# --------------------------------------------------------------------------------------------
def four_term_loss(loss_func, grow_traj, nxt, key_np):
    """physics_train.py:345-352 (fast) == :252-259 (slow, after the [:, key] gather)."""
    return loss_func(grow_traj[:3], nxt[:3, key_np]) + \
        loss_func(grow_traj[7:19], nxt[7:19, key_np]) + \
        loss_func(quaternion_to_euler(grow_traj[3:7]), quaternion_to_euler(nxt[3:7, key_np])) + \
        loss_func(grow_traj[19:], nxt[19:, key_np - 1])


def gen_train():
    out = {}
    train_len = batch_len = 30
    ref = CosseratRod(use_fsolve=True)
    setup_robot(ref)
    trajs, ctls = [], []
    for ctype, carg in [("sine", 0.5), ("sine", 1.0), ("random", 0.0)]:
        c = np.array(calc_controls(ctype, carg, ref.del_t, train_len))
        trajs.append(simulate(ref, c)[:, :25])  # default tolerance, exactly as physics_train.py:116
        ctls.append(c)
    out["traj"] = np.array(trajs)
    out["controls"] = np.array(ctls)
    torch_traj_ls = [torch.tensor(t, requires_grad=True).float() for t in trajs]  # physics_train.py:126
    torch_controls_ls = [torch.tensor(c).float() for c in ctls]

    for mod in [None, "youngs"]:
        tag = "none" if mod is None else mod
        torch.manual_seed(0)
        robot = CosseratRodTorch("cpu", 512)
        setup_robot(robot, mod)
        robot.use_nn = True
        for k, v in mlp_arrays(robot).items():
            out[f"{tag}_init_{k}"] = v
        optimizer = torch.optim.Adam(robot.nn_models.parameters(), lr=1e-2, weight_decay=0)
        loss_func = nn.MSELoss()
        # ---- fast path: physics_train.py:313-368 ----
        grow_loss = 0
        for traj_idx in range(len(torch_traj_ls)):
            traj = torch_traj_ls[traj_idx]
            controls = torch_controls_ls[traj_idx]
            ys = traj[:batch_len - 1, 0:19, :]
            zs = traj[:batch_len - 1, 19:, :]
            y_prevs = torch.cat((ys[:1], ys[:-1]))
            z_prevs = torch.cat((zs[:1], zs[:-1]))
            Gs = traj[1:batch_len]
            key_pt_idx = np.array([3, 5, 7, 9])
            grow_trajs = robot.parallelGetNextSegmentEuler(Gs, key_pt_idx, {
                "yh": robot.c1 * ys + robot.c2 * y_prevs,
                "zh": robot.c1 * zs + robot.c2 * z_prevs,
                "tendon_tensions": controls[:batch_len - 1],
            })
            if traj_idx == 0:
                out[f"{tag}_fast_grow_trajs0"] = grow_trajs.detach().numpy()
            for batch_idx in range(batch_len - 1):
                grow_loss = grow_loss + four_term_loss(loss_func, grow_trajs[batch_idx], traj[batch_idx + 1],
                                                       key_pt_idx)
        total_loss = grow_loss / (batch_len - 1)
        optimizer.zero_grad()
        total_loss.backward()
        out[f"{tag}_fast_loss"] = np.float64(total_loss.item())
        for nm, p in zip(["W1", "b1", "W2", "b2"], robot.nn_models.parameters()):
            out[f"{tag}_fast_grad_{nm}"] = p.grad.numpy().copy()
        # one optimiser step + clamp: physics_train.py:400-408
        optimizer.step()
        for name, param in robot.nn_models.named_parameters():
            if "weight" in name and "layer1" not in name:
                with torch.no_grad():
                    param.clamp_(min=0)
        for k, v in mlp_arrays(robot).items():
            out[f"{tag}_fast_step1_{k}"] = v
        # a second step so Adam's moment/bias-correction handling is pinned too
        grow_loss = 0
        for traj_idx in range(len(torch_traj_ls)):
            traj = torch_traj_ls[traj_idx]
            controls = torch_controls_ls[traj_idx]
            ys = traj[:batch_len - 1, 0:19, :]
            zs = traj[:batch_len - 1, 19:, :]
            y_prevs = torch.cat((ys[:1], ys[:-1]))
            z_prevs = torch.cat((zs[:1], zs[:-1]))
            Gs = traj[1:batch_len]
            key_pt_idx = np.array([3, 5, 7, 9])
            grow_trajs = robot.parallelGetNextSegmentEuler(Gs, key_pt_idx, {
                "yh": robot.c1 * ys + robot.c2 * y_prevs, "zh": robot.c1 * zs + robot.c2 * z_prevs,
                "tendon_tensions": controls[:batch_len - 1]})
            for batch_idx in range(batch_len - 1):
                grow_loss = grow_loss + four_term_loss(loss_func, grow_trajs[batch_idx], traj[batch_idx + 1],
                                                       key_pt_idx)
        total_loss = grow_loss / (batch_len - 1)
        optimizer.zero_grad()
        total_loss.backward()
        optimizer.step()
        for name, param in robot.nn_models.named_parameters():
            if "weight" in name and "layer1" not in name:
                with torch.no_grad():
                    param.clamp_(min=0)
        out[f"{tag}_fast_loss2"] = np.float64(total_loss.item())
        for k, v in mlp_arrays(robot).items():
            out[f"{tag}_fast_step2_{k}"] = v

    # ---- slow path: physics_train.py:215-267 (key [2,6,9]) and train_segment.py:140-185 (key [1,3,6,9]) ----
    for tag, key_list, wd in [("slow", [2, 6, 9], 0.0), ("segment", [1, 3, 6, 9], 0.1)]:
        torch.manual_seed(0)
        robot = CosseratRodTorch("cpu", 512)
        if tag == "slow":
            setup_robot(robot)  # train_segment.py never calls setup_robot (class-default params)
        robot.use_nn = True
        for k, v in mlp_arrays(robot).items():
            out[f"{tag}_init_{k}"] = v
        optimizer = torch.optim.Adam(robot.nn_models.parameters(), lr=1e-2, weight_decay=wd)
        loss_func = nn.MSELoss()
        epoch = 0
        grow_loss = 0
        for traj_idx in range(2):
            traj = torch_traj_ls[traj_idx]
            controls = torch_controls_ls[traj_idx]
            for stp_idx in range(batch_len - 1):
                batch_idx = ((epoch % batch_len - 1) * batch_len + stp_idx + batch_len) % train_len
                if batch_idx >= train_len - 1:
                    break
                y = traj[batch_idx, 0:19, :]
                z = traj[batch_idx, 19:, :]
                if stp_idx == 0:
                    y_prev = y.clone().requires_grad_(True)
                    z_prev = z.clone().requires_grad_(True)
                else:
                    y_prev = traj[batch_idx - 1, 0:19, :]
                    z_prev = traj[batch_idx - 1, 19:, :]
                robot.y = y
                robot.z = z
                G = torch.cat((traj[batch_idx + 1, :19, :], traj[batch_idx + 1, 19:, :])).clone().requires_grad_(True)
                robot.tendon_tensions = controls[batch_idx]
                yh = robot.c1 * robot.y + robot.c2 * y_prev
                zh = robot.c1 * robot.z + robot.c2 * z_prev
                robot.residualArgs["yh"] = yh
                robot.residualArgs["zh"] = zh
                grow_traj = robot.getNextSegmentEuler(G)
                if traj_idx == 0 and stp_idx == 3:
                    out[f"{tag}_grow_traj_t3"] = grow_traj.detach().numpy()
                key_pt_idx = torch.tensor(key_list)
                grow_loss = grow_loss + four_term_loss(loss_func, grow_traj[:, key_pt_idx], traj[batch_idx + 1],
                                                       key_pt_idx)
        total_loss = grow_loss / (batch_len - 1)
        optimizer.zero_grad()
        total_loss.backward()
        out[f"{tag}_loss"] = np.float64(total_loss.item())
        for nm, p in zip(["W1", "b1", "W2", "b2"], robot.nn_models.parameters()):
            out[f"{tag}_grad_{nm}"] = p.grad.numpy().copy()
        optimizer.step()
        for name, param in robot.nn_models.named_parameters():
            if "weight" in name and "layer1" not in name:
                with torch.no_grad():
                    param.clamp_(min=0)
        for k, v in mlp_arrays(robot).items():
            out[f"{tag}_step1_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "train.npz"), **out)
    print("train.npz losses", {k: float(v) for k, v in out.items() if "loss" in k})


def gen_misc():
    out = {}
    rng = np.random.default_rng(2)
    q = rng.standard_normal((4, 50))
    q[:, :5] = np.array([[1, 0, 0, 0], [0.7071, 0.7071, 0, 0], [0.5, 0.5, 0.5, 0.5], [0.7071, 0, 0, 0.7072],
                         [1, 1e-4, -2e-4, 3e-4]]).T
    out["quat"] = q
    out["euler"] = quaternion_to_euler(torch.tensor(q)).numpy()
    for ctype, carg, dt, T in [("sine", 1.0, 0.05, 30), ("sine", 0.5, 0.005, 40), ("step", 1.0, 0.05, 40),
                               ("random", 0.0, 0.05, 30), ("random", 3.0, 0.05, 10)]:
        out[f"ctl_{ctype}_{carg}_{dt}_{T}"] = np.array(calc_controls(ctype, carg, dt, T))
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)
    print("misc.npz")


if __name__ == "__main__":
    torch.set_num_threads(1)
    which = sys.argv[1:] or ["rollouts", "knode", "ode", "train", "misc"]
    if "rollouts" in which:
        gen_rollouts()
    if "knode" in which:
        gen_knode_rollouts()
    if "ode" in which:
        gen_ode()
    if "train" in which:
        gen_train()
    if "misc" in which:
        gen_misc()
