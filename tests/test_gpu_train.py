"""GPU parity of the training-step kernels (kc_train_step, kc_adam_clamp, kc_ode_bwd) through the C ABI, against the
reference's own autograd results (tests/golden/train.npz, ode.npz) and the numpy oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import rod_oracle as O

pytestmark = pytest.mark.gpu
PK = ("W1", "b1", "W2", "b2")


@pytest.fixture(scope="module")
def ops():
    import _kc
    import _ops
    assert torch.cuda.is_available()
    _kc.lib()
    return _ops


def params(P):
    import _kc
    return _kc.rod_params(P)


def P_setup(mod=None):
    return O.setup_params(O.RodParams(), mod)


def dev(a, dt):
    return torch.tensor(np.asarray(a), dtype=dt, device="cuda")


CASES = [("none", None, [3, 5, 7, 9], 3, "none_fast"), ("youngs", "youngs", [3, 5, 7, 9], 3, "youngs_fast"),
         ("slow", None, [2, 6, 9], 2, "slow"), ("segment", "classdefault", [1, 3, 6, 9], 2, "segment")]


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag,mod,key,ntraj,prefix", CASES)
def test_train_step_vs_reference_autograd(ops, golden, dt, tag, mod, key, ntraj, prefix):
    """loss and dW of one full-batch step: the reference computed them in fp32 with torch.autograd."""
    d = golden["train"]
    traj = d["traj"].astype(np.float32)[:ntraj]
    ctl = d["controls"].astype(np.float32)[:ntraj]
    P = O.RodParams() if mod == "classdefault" else P_setup(mod)
    mlp = ops.Mlp(*[dev(d[f"{tag}_init_{k}"], dt) for k in PK])
    loss, grads, pred = ops.train_step(params(P), mlp, dev(traj, dt), dev(ctl, dt), key, want_pred=True)
    ref_loss = float(d[f"{prefix}_loss"])
    assert abs(float(loss.item()) - ref_loss) < 3e-5 * abs(ref_loss)
    for k, g in zip(PK, grads):
        ref = d[f"{prefix}_grad_{k}"]
        assert np.max(np.abs(g.cpu().numpy() - ref)) < 1e-4 * np.abs(ref).max(), k
    # and against the fp64 oracle on the same fp32-rounded inputs (tight in fp64)
    mlp_np = {k: d[f"{tag}_init_{k}"].astype(np.float64) for k in PK}
    o_loss, o_grads, o_pred = O.teacher_forced_loss_and_grads(P, traj.astype(np.float64), ctl.astype(np.float64), key, mlp_np)
    tol = 1e-9 if dt == torch.float64 else 1e-4
    assert abs(float(loss.item()) - o_loss) < tol * abs(o_loss)
    for k, g in zip(PK, grads):
        assert np.max(np.abs(g.cpu().numpy() - o_grads[k])) < tol * np.abs(o_grads[k]).max(), k
    pr = pred.cpu().numpy().astype(np.float64)
    scale = np.abs(o_pred).max((0, 1, 3), keepdims=True) + 1e-6
    assert np.max(np.abs(pr - o_pred) / scale) < tol
    if prefix.endswith("fast"):
        ref = d[f"{tag}_fast_grow_trajs0"].astype(np.float64)
        assert np.max(np.abs(pr[0] - ref) / scale[0]) < 1e-4


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_train_step_history_inputs_and_odd_hidden(ops, golden, dt):
    """nn_input_history (53 inputs) and a hidden width that is not a multiple of the CTA's 32-unit chunk."""
    rng = np.random.default_rng(3)
    P = P_setup()
    d = golden["train"]
    traj = d["traj"][1:3, :9].astype(np.float32).astype(np.float64)
    ctl = d["controls"][1:3, :9].astype(np.float32).astype(np.float64)
    for in_dim, H in ((53, 40), (28, 70)):
        mlp_np = {"W1": np.abs(rng.normal(0.01, 0.01, (H, in_dim))) * (0.05 if in_dim == 53 else 1.0),
                  "b1": rng.normal(0, 0.01, H), "W2": np.abs(rng.normal(0.01, 0.01, (25, H))), "b2": rng.normal(0, 0.01, 25)}
        o_loss, o_grads, _ = O.teacher_forced_loss_and_grads(P, traj, ctl, [1, 4, 9], mlp_np, nn_input_history=in_dim == 53)
        mlp = ops.Mlp(*[dev(mlp_np[k], dt) for k in PK])
        loss, grads, _ = ops.train_step(params(P), mlp, dev(traj, dt), dev(ctl, dt), [1, 4, 9])
        tol = 1e-9 if dt == torch.float64 else 2e-4
        assert abs(float(loss.item()) - o_loss) < tol * abs(o_loss)
        for k, g in zip(PK, grads):
            assert np.max(np.abs(g.cpu().numpy() - o_grads[k])) < tol * np.abs(o_grads[k]).max(), (in_dim, k)


def test_adam_clamp_two_steps_vs_torch_optim(ops, golden):
    """Two optimiser steps of physics_train.py:393-408 (Adam lr 1e-2 + clamp), weights compared with the reference's."""
    d = golden["train"]
    dt = torch.float32
    P = P_setup()
    traj, ctl = dev(d["traj"].astype(np.float32), dt), dev(d["controls"].astype(np.float32), dt)
    W = [dev(d[f"none_init_{k}"], dt) for k in PK]
    m = [torch.zeros_like(w) for w in W]
    v = [torch.zeros_like(w) for w in W]
    for step in (1, 2):
        loss, grads, _ = ops.train_step(params(P), ops.Mlp(*W), traj, ctl, [3, 5, 7, 9])
        for i, k in enumerate(PK):
            ops.adam_clamp(W[i], grads[i], m[i], v[i], step, lr=1e-2, clamp=k in ("W1", "W2"))
        for i, k in enumerate(PK):
            ref = d[f"none_fast_step{step}_{k}"]
            assert np.max(np.abs(W[i].cpu().numpy() - ref)) < 2e-4 * max(np.abs(ref).max(), 1e-2), (step, k)
    assert abs(float(loss.item()) - float(d["none_fast_loss2"])) < 5e-3 * float(d["none_fast_loss2"])
    # weight decay 0.1 (train_segment.py:118-120)
    W = [dev(d[f"segment_init_{k}"], dt) for k in PK]
    m = [torch.zeros_like(w) for w in W]
    v = [torch.zeros_like(w) for w in W]
    loss, grads, _ = ops.train_step(params(O.RodParams()), ops.Mlp(*W), traj[:2], ctl[:2], [1, 3, 6, 9])
    for i, k in enumerate(PK):
        ops.adam_clamp(W[i], grads[i], m[i], v[i], 1, lr=1e-2, weight_decay=0.1, clamp=k in ("W1", "W2"))
        ref = d[f"segment_step1_{k}"]
        assert np.max(np.abs(W[i].cpu().numpy() - ref)) < 2e-4 * max(np.abs(ref).max(), 1e-2), k


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag", ["h512", "h64hist"])
def test_ode_bwd_vs_reference_autograd(ops, golden, dt, tag):
    """kc_ode_bwd against torch.autograd through the reference's ODE_parallel (fp64), inputs and parameters."""
    d = golden["ode"]
    mlp = ops.Mlp(*[dev(d[f"{tag}_{k}"], dt) for k in PK])
    out = ops.ode_bwd(params(P_setup()), mlp, dev(d["y"], dt), dev(d["yh"], dt), dev(d["zh"], dt), dev(d["tf"], dt),
                      dev(d[f"{tag}_cot_ys"], dt), dev(d[f"{tag}_cot_z"], dt))
    tol = 1e-9 if dt == torch.float64 else 2e-4
    for name, g in zip(["y", "yh", "zh", "tf", "W1", "b1", "W2", "b2"], out):
        ref = d[f"{tag}_grad64_{name}"]
        g = g.cpu().numpy().astype(np.float64)
        if ref.ndim == 2 and name in ("y", "yh", "zh", "tf"):
            scale = np.abs(ref).max(0, keepdims=True) + 1e-6 * np.abs(ref).max()
            err = np.max(np.abs(g - ref) / scale)
        else:
            err = np.max(np.abs(g - ref)) / np.abs(ref).max()
        assert err < tol, (name, err)


def test_ode_bwd_physics_only_finite_difference(ops):
    """Physics-only adjoint (dense user matrices, so the non-diagonal code path) against central differences."""
    rng = np.random.default_rng(5)
    P = P_setup()
    P.Bse = 1e-3 * rng.standard_normal((3, 3))
    P.Bbt = P.Bbt + 1e-3 * rng.standard_normal((3, 3))
    P.compute_intermediate_terms()
    Q = 9
    y = rng.standard_normal((Q, 19)); y[:, 3] += 2.0
    yh, zh, tf = rng.standard_normal((Q, 19)), rng.standard_normal((Q, 6)), rng.standard_normal((Q, 3))
    cy, cz = rng.standard_normal((Q, 19)), rng.standard_normal((Q, 6))
    dt = torch.float64
    g = ops.ode_bwd(params(P), None, dev(y, dt), dev(yh, dt), dev(zh, dt), dev(tf, dt), dev(cy, dt), dev(cz, dt))

    def f(y_, yh_, zh_, tf_):
        a, b = O.ode(P, y_, yh_, zh_, tf_)
        return np.sum(a * cy, 1) + np.sum(b * cz, 1)
    args = [y, yh, zh, tf]
    for ai, got in enumerate(g[:4]):
        got = got.cpu().numpy()
        for k in range(args[ai].shape[1]):
            e = 1e-6
            ap = [a.copy() for a in args]; am = [a.copy() for a in args]
            ap[ai][:, k] += e; am[ai][:, k] -= e
            fd = (f(*ap) - f(*am)) / (2 * e)
            assert np.max(np.abs(fd - got[:, k])) < 1e-6 * (1 + np.abs(fd).max()), (ai, k)


def _golden_robot(d, tag="none"):
    """A CosseratRodTorch whose MLP holds the golden initial weights (hidden size from the fixture)."""
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    W = {k: d[f"{tag}_init_{k}"] for k in PK}
    robot = CosseratRodTorch("cuda", int(W["W1"].shape[0]))
    setup_robot(robot)
    with torch.no_grad():
        for p, k in zip(robot.nn_models.parameters(), PK):
            p.copy_(torch.tensor(W[k], device="cuda"))
    return robot


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_trainer_steps_vs_reference(ops, golden, use_graph):
    """TeacherForcedTrainer.fused_step (preallocated step + ONE kc_adam_clamp_multi launch, CUDA graph from the third
    epoch) against the reference's weights after two Adam steps, and against the unfused trainer over six epochs."""
    from _train import TeacherForcedTrainer
    d = golden["train"]
    traj = dev(d["traj"].astype(np.float32), torch.float32)
    ctl = dev(d["controls"].astype(np.float32), torch.float32)
    a = TeacherForcedTrainer(_golden_robot(d), traj, ctl, [3, 5, 7, 9], fused=True, use_graph=use_graph)
    b = TeacherForcedTrainer(_golden_robot(d), traj, ctl, [3, 5, 7, 9], fused=False)
    la, lb = [], []
    for epoch in range(6):
        la.append(a.step())
        lb.append(b.step())
        if epoch == 1:
            for p, k in zip(a.params, PK):
                ref = d[f"none_fast_step2_{k}"]
                assert np.max(np.abs(p.detach().cpu().numpy() - ref)) < 2e-4 * max(np.abs(ref).max(), 1e-2), k
    assert (a.graph is not None) == use_graph
    assert int(a.adam.step_dev.item()) == 6 and a.step_no == 6
    np.testing.assert_allclose(la, lb, rtol=2e-5)
    for pa, pb in zip(a.params, b.params):
        np.testing.assert_allclose(pa.detach().cpu().numpy(), pb.detach().cpu().numpy(), rtol=1e-4, atol=1e-6)
    # the learning rate lives on the device: a scheduler change reaches the (captured) update
    a.sched.lr = 0.0
    before = [p.detach().clone() for p in a.params]
    a.step()
    for p, q in zip(a.params, before):
        assert torch.equal(p.detach(), q)


def test_adam_clamp_multi_matches_single_tensor_kernels(ops):
    rng = np.random.default_rng(3)
    shapes = [(17, 28), (17,), (25, 17), (25,)]
    P1 = [dev(rng.normal(size=s), torch.float64) for s in shapes]
    P2 = [p.clone() for p in P1]
    m1 = [torch.zeros_like(p) for p in P1]
    v1 = [torch.zeros_like(p) for p in P1]
    grads = [torch.zeros_like(p) for p in P1]
    multi = ops.AdamClampMulti(P2, grads, [True, False, True, False], lr=3e-3, weight_decay=0.05)
    for step in (1, 2, 3):
        for g in grads:
            g.copy_(dev(rng.normal(size=tuple(g.shape)), torch.float64))
        for i in range(4):
            ops.adam_clamp(P1[i], grads[i], m1[i], v1[i], step, lr=3e-3, weight_decay=0.05, clamp=i % 2 == 0)
        multi.run()
        for a, b in zip(P1, P2):
            np.testing.assert_allclose(b.cpu().numpy(), a.cpu().numpy(), rtol=1e-12, atol=1e-14)
    assert int(multi.step_dev.item()) == 3 and int(multi.ticket.item()) == 0


@pytest.mark.parametrize("H", [512, 200])
def test_train_step_kernel_generations_agree(ops, monkeypatch, H):
    """The three generations of the tensor-core training kernel (KC_TRAIN_TC = 1 | 2 | default 3; kc_train_tc.cu,
    kc_train_tc2.cu, kc_train_tc3.cu) and the fused-prep variant of the third compute the same step: loss and gradients agree
    with the fp64 kernels at the fp32 tolerance, on a batch with several tiles per CTA and a ragged last tile."""
    import physics_controls
    P = P_setup()
    B, T, key = 200, 30, [3, 5, 7, 9]              # 23 200 samples: 182 tiles, the last one partly filled
    ctl = physics_controls.synthetic_tensions(B, T, P.del_t, seed=5, dtype=np.float32)
    traj32, _, iters = ops.rollout(params(P), None, dev(ctl, torch.float32))
    assert int(iters.min()) >= 0
    g = torch.Generator().manual_seed(H)
    W = [torch.normal(0.01, 0.02, (H, 28), generator=g), torch.normal(0.0, 0.01, (H,), generator=g),
         torch.normal(0.01, 0.02, (25, H), generator=g), torch.normal(0.0, 0.01, (25,), generator=g)]

    def run(dt, gen, fuse=None):
        for k, v in (("KC_TRAIN_TC", gen), ("KC_TRAIN_FUSE_PREP", fuse)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        mlp = ops.Mlp(*[w.to("cuda", dt) for w in W])
        loss, grads, _ = ops.train_step(params(P), mlp, traj32.to(dt), dev(ctl, dt), key)
        return float(loss.item()), [x.cpu().numpy().astype(np.float64) for x in grads]

    l64, g64 = run(torch.float64, None)
    for gen, fuse in (("1", None), ("2", None), (None, "0"), (None, "1")):
        l, gr = run(torch.float32, gen, fuse)
        assert abs(l - l64) < 1e-5 * abs(l64), (gen, fuse)
        for k, a, r in zip(PK, gr, g64):
            assert np.abs(a - r).max() < 1e-4 * np.abs(r).max(), (gen, fuse, k)
    monkeypatch.delenv("KC_TRAIN_TC", raising=False)
    monkeypatch.delenv("KC_TRAIN_FUSE_PREP", raising=False)


@pytest.mark.parametrize("H", [512, 200, 64])
def test_param_grads_tensor_core_vs_simt_and_fp64(ops, monkeypatch, H):
    """kc_mlp_bwd with enough samples takes the tcgen05 gradient kernel (kc_train_tc_kernel<2>); same numbers as the SIMT
    kernel and as fp64 (3-pass bf16/tf32 splits keep fp32 accuracy).  H = 200: a partly filled last 128-unit chunk."""
    rng = np.random.default_rng(H)
    Q = 6000 + 37                                   # not a multiple of the 128-sample tile
    W = [rng.normal(0, 0.3, (H, 28)), rng.normal(0, 0.1, H), rng.normal(0, 0.3, (25, H)), rng.normal(0, 0.1, 25)]
    x = rng.normal(0, 1.0, (Q, 28))
    go = rng.normal(0, 1.0, (Q, 25))

    def run(dt, mode):
        if mode:
            monkeypatch.setenv("KC_TRAIN_MODE", mode)
        else:
            monkeypatch.delenv("KC_TRAIN_MODE", raising=False)
        mlp = ops.Mlp(*[dev(w, dt) for w in W])
        out = ops.mlp_bwd(mlp, dev(x, dt), dev(go, dt))
        return [o.cpu().numpy().astype(np.float64) for o in out]

    ref = run(torch.float64, None)
    simt = run(torch.float32, "simt")
    tc = run(torch.float32, None)
    for name, r, a, b in zip(("gx",) + PK, ref, simt, tc):
        scale = np.abs(r).max()
        assert np.abs(a - r).max() < 1e-4 * scale, (name, "simt")
        assert np.abs(b - r).max() < 1e-4 * scale, (name, "tc", np.abs(b - r).max() / scale)


@pytest.mark.parametrize("H", [512, 200])
def test_forward_tensor_core_vs_simt_and_fp64(ops, monkeypatch, H):
    """kc_mlp_fwd / kc_ode_fwd with >= 4096 samples run the MLP contraction on tcgen05 (kc_train_tc_kernel<1>, physics on
    the SIMT pipes): same numbers as the SIMT kernels and as fp64."""
    rng = np.random.default_rng(100 + H)
    Q = 5000 + 61
    W = [rng.normal(0, 0.3, (H, 28)), rng.normal(0, 0.1, H), rng.normal(0, 0.05, (25, H)), rng.normal(0, 0.1, 25)]
    x = rng.normal(0, 1.0, (Q, 28))
    y = rng.normal(0, 1.0, (Q, 19)); y[:, 3] += 2.0
    yh, zh, tf = rng.normal(0, 1.0, (Q, 19)), rng.normal(0, 1.0, (Q, 6)), rng.normal(0, 1.0, (Q, 3))
    P = params(P_setup())

    def run(dt, mode):
        if mode:
            monkeypatch.setenv("KC_TRAIN_MODE", mode)
        else:
            monkeypatch.delenv("KC_TRAIN_MODE", raising=False)
        mlp = ops.Mlp(*[dev(w, dt) for w in W])
        o = ops.mlp_fwd(mlp, dev(x, dt))
        ys, z = ops.ode_fwd(P, mlp, dev(y, dt), dev(yh, dt), dev(zh, dt), dev(tf, dt))
        return [t.cpu().numpy().astype(np.float64) for t in (o, ys, z)]

    ref = run(torch.float64, None)
    simt = run(torch.float32, "simt")
    tc = run(torch.float32, None)
    for name, r, a, b in zip(("mlp", "ys", "z"), ref, simt, tc):
        assert col_err(a, r) < 1e-4, (name, "simt")
        assert col_err(b, r) < 1e-4, (name, "tc", col_err(b, r))


def col_err(a, b, floor=1e-3):
    scale = np.abs(b).reshape(-1, b.shape[-1]).max(0) + floor
    return float(np.max(np.abs(a - b) / scale))


@pytest.mark.parametrize("H", [512, 200])
def test_ode_bwd_tensor_core_vs_simt_and_fp64(ops, monkeypatch, H):
    """kc_ode_bwd / kc_mlp_bwd with >= 4096 samples: MLP input gradients (kc_train_tc_kernel<3>) and parameter gradients
    (<2>) on tcgen05 around the SIMT physics adjoint; same numbers as the all-SIMT path and as fp64."""
    rng = np.random.default_rng(300 + H)
    Q = 4500 + 13
    W = [rng.normal(0, 0.3, (H, 28)), rng.normal(0, 0.1, H), rng.normal(0, 0.05, (25, H)), rng.normal(0, 0.1, 25)]
    y = rng.normal(0, 1.0, (Q, 19)); y[:, 3] += 2.0
    yh, zh, tf = rng.normal(0, 1.0, (Q, 19)), rng.normal(0, 1.0, (Q, 6)), rng.normal(0, 1.0, (Q, 3))
    gys, gz = rng.normal(0, 1.0, (Q, 19)), rng.normal(0, 1.0, (Q, 6))
    P = params(P_setup())

    def run(dt, mode):
        if mode:
            monkeypatch.setenv("KC_TRAIN_MODE", mode)
        else:
            monkeypatch.delenv("KC_TRAIN_MODE", raising=False)
        mlp = ops.Mlp(*[dev(w, dt) for w in W])
        out = ops.ode_bwd(P, mlp, dev(y, dt), dev(yh, dt), dev(zh, dt), dev(tf, dt), dev(gys, dt), dev(gz, dt))
        return [t.cpu().numpy().astype(np.float64) for t in out]

    ref = run(torch.float64, None)
    simt = run(torch.float32, "simt")
    tc = run(torch.float32, None)
    names = ("g_y", "g_yh", "g_zh", "g_tf") + PK
    for name, r, a, b in zip(names, ref, simt, tc):
        scale = np.abs(r).max() + 1e-30
        assert np.abs(a - r).max() < 2e-4 * scale, (name, "simt", np.abs(a - r).max() / scale)
        assert np.abs(b - r).max() < 2e-4 * scale, (name, "tc", np.abs(b - r).max() / scale)


def test_train_step_config3_full_size_properties(ops, monkeypatch):
    """BASELINE config 3 at full size (1024 trajectories x 30 time indices, key nodes [3,5,7,9] -> 118 784 samples, H = 512,
    the reference's weight init): size-independent properties of kc_train_step.  (1) the fp32 tensor-core kernel, the fp32
    SIMT kernels and the fp64 kernels agree on loss and gradients; (2) loss and gradients are additive over a split of the
    batch (the loss is a sum over trajectories divided by a global constant, physics_train.py:266-267); (3) two runs are
    bitwise identical (no atomics); (4) a sampled sub-batch agrees with the numpy oracle."""
    import physics_controls
    P = P_setup()
    B, T, key = 1024, 30, [3, 5, 7, 9]
    ctl = physics_controls.synthetic_tensions(B, T, P.del_t, seed=3, dtype=np.float32)
    traj32, _, iters = ops.rollout(params(P), None, dev(ctl, torch.float32))
    assert int(iters.min()) >= 0
    g = torch.Generator().manual_seed(0)
    W = [torch.normal(0.01, 0.01, (512, 28), generator=g).abs(), torch.normal(0.0, 0.01, (512,), generator=g),
         torch.normal(0.01, 0.01, (25, 512), generator=g).abs(), torch.normal(0.0, 0.01, (25,), generator=g)]

    def run(dt, mode, sl=slice(None)):
        if mode:
            monkeypatch.setenv("KC_TRAIN_MODE", mode)
        else:
            monkeypatch.delenv("KC_TRAIN_MODE", raising=False)
        mlp = ops.Mlp(*[w.to("cuda", dt) for w in W])
        loss, grads, _ = ops.train_step(params(P), mlp, traj32[sl].to(dt), dev(ctl[sl], dt), key)
        return float(loss.item()), [x.cpu().numpy().astype(np.float64) for x in grads], grads

    l64, g64, _ = run(torch.float64, None)
    ltc, gtc, raw1 = run(torch.float32, None)
    lsi, gsi, _ = run(torch.float32, "simt")
    _, _, raw2 = run(torch.float32, None)
    assert np.isfinite(l64) and l64 > 0
    for name, l, gr in (("tc", ltc, gtc), ("simt", lsi, gsi)):
        assert abs(l - l64) < 1e-5 * l64, name
        for k, a, r in zip(PK, gr, g64):
            assert np.abs(a - r).max() < 1e-4 * np.abs(r).max(), (name, k)
    for a, b in zip(raw1, raw2):
        assert torch.equal(a, b)                                             # deterministic reduction
    # (5) the kernel that forms its own samples (default only below one tile per SM) and the stand-alone prep kernel feed
    # the same values to the same arithmetic: bitwise identical, at the full size and at a one-tile-per-CTA size
    for sl_ in (slice(None), slice(0, 100)):
        res = []
        for fuse in ("1", "0"):
            monkeypatch.setenv("KC_TRAIN_FUSE_PREP", fuse)
            lf, _, rawf = run(torch.float32, None, sl_)
            res.append((lf, rawf))
        monkeypatch.delenv("KC_TRAIN_FUSE_PREP", raising=False)
        assert res[0][0] == res[1][0]
        for a, b in zip(res[0][1], res[1][1]):
            assert torch.equal(a, b)
    la, ga, _ = run(torch.float64, None, slice(0, 300))
    lb, gb, _ = run(torch.float64, None, slice(300, B))
    assert abs(la + lb - l64) < 1e-11 * l64
    for k, a, b_, r in zip(PK, ga, gb, g64):
        assert np.abs(a + b_ - r).max() < 1e-11 * np.abs(r).max(), k
    lt1, gt1, _ = run(torch.float32, None, slice(0, 300))
    lt2, gt2, _ = run(torch.float32, None, slice(300, B))
    assert abs(lt1 + lt2 - ltc) < 1e-5 * ltc
    for k, a, b_, r in zip(PK, gt1, gt2, gtc):
        assert np.abs(a + b_ - r).max() < 1e-4 * np.abs(r).max(), k
    sel = [0, 511, 1023]
    tr = traj32[sel].cpu().numpy().astype(np.float64)
    mlp_np = {k: w.numpy().astype(np.float64) for k, w in zip(PK, W)}
    o_loss, o_grads, _ = O.teacher_forced_loss_and_grads(P, tr, ctl[sel].astype(np.float64), key, mlp_np)
    ls, gs, _ = run(torch.float32, None, sel)
    assert abs(ls - o_loss) < 1e-4 * abs(o_loss)
    for k, a in zip(PK, gs):
        assert np.abs(a - o_grads[k]).max() < 1e-4 * np.abs(o_grads[k]).max(), k


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_peer_allreduce_adam_emulated_ranks(dt):
    """kc_peer_publish / kc_peer_gather_adam (all-reduce over peer memory fused with Adam + clamp) with 3 emulated ranks on
    one GPU: the three regions are ordinary device buffers, all ranks publish first and gather afterwards (so nothing
    waits), two consecutive steps (both slots).  Against: sum of the ranks' flat buffers in rank order, then
    kc_adam_clamp_multi on that sum — bitwise."""
    import _kc
    import _ops
    torch.manual_seed(0)
    W, sizes = 3, [(7, 5), (7,), (4, 7), (4,)]
    n = sum(int(np.prod(s)) for s in sizes) + 1
    nbytes = int(_kc.lib().kc_peer_region_bytes(_ops._DT[dt], n))
    regions = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(W)]
    ranks = []
    for r in range(W):
        flat = torch.zeros(n, dtype=dt, device="cuda")
        grads, off = [], 0
        for s in sizes:
            k = int(np.prod(s))
            grads.append(flat[off:off + k].view(s))
            off += k
        torch.manual_seed(1)
        params = [torch.rand(s, dtype=dt, device="cuda") * 0.1 for s in sizes]      # same weights on every rank
        adam = _ops.AdamClampMulti(params, grads, [True, False, True, False], lr=1e-2)
        assert adam.enable_peer_allreduce(flat, regions=[t.data_ptr() for t in regions], rank=r, world=W)
        ranks.append((flat, params, adam))
    torch.manual_seed(2)
    ref_params = [p.clone() for p in ranks[0][1]]
    ref_flat = torch.zeros(n, dtype=dt, device="cuda")
    ref_grads, off = [], 0
    for s in sizes:
        k = int(np.prod(s))
        ref_grads.append(ref_flat[off:off + k].view(s))
        off += k
    ref_adam = _ops.AdamClampMulti(ref_params, ref_grads, [True, False, True, False], lr=1e-2)
    L = _kc.lib()
    for step in range(3):
        locals_ = [torch.randn(n, dtype=dt, device="cuda") for _ in range(W)]
        tot = locals_[0].clone()
        for r in range(1, W):
            tot = tot + locals_[r]                                   # rank order, as the kernel adds
        for r, (flat, _, adam) in enumerate(ranks):
            flat.copy_(locals_[r])
            st = _ops._stream(flat.device)
            _kc.check(L.kc_peer_publish(_ops._DT[dt], W, r, C.cast(adam._peer_regions, C.c_void_p), n, _ops._ptr(flat),
                                        _ops._ptr(adam.step_dev), _ops._ptr(adam._peer_tickets), st), "publish")
        for r, (flat, _, adam) in enumerate(ranks):
            st = _ops._stream(flat.device)
            _kc.check(L.kc_peer_gather_adam(_ops._DT[dt], W, r, C.cast(adam._peer_regions, C.c_void_p), n, _ops._ptr(flat),
                                            4, C.cast(adam.arr, C.c_void_p), _ops._ptr(adam.step_dev), _ops._ptr(adam.lr_dev),
                                            0.9, 0.999, 1e-8, 0.0, adam._peer_tickets[1:].data_ptr(), st), "gather")
        ref_flat.copy_(tot)
        ref_adam.run()
        torch.cuda.synchronize()
        for flat, params, adam in ranks:
            assert torch.equal(flat, tot)
            assert int(adam.step_dev.item()) == step + 1
            for p, q in zip(params, ref_params):
                assert torch.equal(p, q)


def test_device_plateau_scheduler_matches_torch():
    """kc_plateau_step (ReduceLROnPlateau on the device, stepped inside the captured training step) against
    torch.optim.lr_scheduler.ReduceLROnPlateau on a loss sequence with plateaus (physics_train.py:206,297)."""
    import _ops
    rng = np.random.default_rng(0)
    losses = np.concatenate([np.linspace(1, 0.5, 20), 0.5 + 0.01 * rng.random(40), np.linspace(0.5, 0.4, 5),
                             0.4 + 0.01 * rng.random(30)]).astype(np.float32)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-2)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, 'min', patience=7, factor=0.5)
    lr_dev = torch.full((1,), 1e-2, dtype=torch.float64, device="cuda")
    dev = _ops.PlateauDevice(lr_dev, patience=7, factor=0.5)
    ld = torch.tensor(losses, device="cuda")
    for i, l in enumerate(losses):
        ref.step(float(l))
        dev.run(ld[i:i + 1])
        assert abs(dev.get_last_lr()[0] - opt.param_groups[0]["lr"]) < 1e-15, i
    assert dev.get_last_lr()[0] < 1e-2


def test_fused_trainer_runs_the_scheduler_on_the_device(golden):
    """TeacherForcedTrainer.fused_step(sync=False) steps ReduceLROnPlateau inside the step: with patience 0 and a loss
    that cannot improve every step, the learning rate on the device halves without any host read."""
    from _train import TeacherForcedTrainer
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    d = golden["train"]
    torch.manual_seed(0)
    r = CosseratRodTorch("cuda", 32)
    setup_robot(r)
    traj = torch.tensor(d["traj"], dtype=torch.float32, device="cuda")
    ctl = torch.tensor(d["controls"], dtype=torch.float32, device="cuda")
    tr = TeacherForcedTrainer(r, traj, ctl, [3, 5, 7, 9], lr=0.5, patience=0, factor=0.5)   # a step size that overshoots
    for _ in range(12):
        tr.fused_step(train=True, sync=False)
    assert tr.sched.get_last_lr()[0] < 0.5
    assert tr.loss_arr == []
