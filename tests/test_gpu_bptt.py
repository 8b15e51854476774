"""Reverse mode through the rollout (kc_rollout_bwd, north-star subsystem 3).  The reference has no implementation of
this (it never differentiates a rollout), so the oracle is a central finite difference of the fp64 oracle rollout
(oracle/rod_oracle.py, itself pinned to the reference's simulate) with respect to individual weights and tensions."""
import numpy as np
import pytest
import torch

from oracle import rod_oracle as O

pytestmark = pytest.mark.gpu
PK = ("W1", "b1", "W2", "b2")


@pytest.fixture(params=["coop", "thread", "tc"], autouse=True)
def coop(request, monkeypatch):
    """All KNODE kernels: one rod per warp with the MLP split over the lanes ("coop", small batches), one rod per thread
    ("thread"), and — fp32, 28 inputs — the tensor-core march ("tc": 16 rods x 8 rows per CTA, MLP on tcgen05; other
    shapes fall through to the launcher's default)."""
    if request.param == "tc":
        monkeypatch.delenv("KC_ROLLOUT_COOP", raising=False)
        monkeypatch.setenv("KC_ROLLOUT_TC", "1")
    else:
        monkeypatch.setenv("KC_ROLLOUT_TC", "0")
        monkeypatch.setenv("KC_ROLLOUT_COOP", "1" if request.param == "coop" else "0")
    return request.param


def setup_case(H=16, B=2, T=7, in_dim=28, seed=0):
    rng = np.random.default_rng(seed)
    P = O.setup_params(O.RodParams(), "youngs")
    ctl = np.stack([np.array(O.calc_controls("sine", 0.7 + 0.2 * b, P.del_t, T)) if b % 2 == 0
                    else 5 + 5 * rng.random((T, 4)) for b in range(B)])
    mlp = {"W1": np.abs(rng.normal(0.01, 0.01, (H, in_dim))) * (0.05 if in_dim == 53 else 1.0),
           "b1": rng.normal(0, 0.01, H), "W2": np.abs(rng.normal(0.01, 0.01, (25, H))) * 0.3,
           "b2": rng.normal(0, 0.01, 25) * 0.3}
    Cw = rng.standard_normal((B, T, 25, P.N))
    return P, ctl, mlp, Cw


def gpu_grads(P, ctl, mlp, Cw, dt):
    import _kc
    import _ops
    dev = lambda a: torch.tensor(np.asarray(a), dtype=dt, device="cuda")
    m = None if mlp is None else _ops.Mlp(*[dev(mlp[k]) for k in PK])
    traj, _, iters = _ops.rollout(_kc.rod_params(P), m, dev(ctl), rows=25, tol=1e-13 if dt == torch.float64 else 0.0)
    assert int(iters.min()) >= 0
    out = _ops.rollout_bwd(_kc.rod_params(P), m, dev(ctl), traj, dev(Cw))
    return traj.cpu().numpy(), [None if g is None else g.cpu().numpy().astype(np.float64) for g in out]


def test_bptt_matches_finite_differences_of_the_oracle():
    P, ctl, mlp, Cw = setup_case()
    traj, (gten, gW1, gb1, gW2, gb2) = gpu_grads(P, ctl, mlp, Cw, torch.float64)
    an = dict(zip(PK, (gW1, gb1, gW2, gb2)))

    def loss(m, c=ctl):
        return float(np.sum(Cw * O.rollout_newton(P, c, m, rows=25, tol=1e-13)))
    np.testing.assert_allclose(traj, O.rollout_newton(P, ctl, mlp, rows=25, tol=1e-13), rtol=0, atol=1e-9)
    e = 1e-6
    for name, idx in [("W1", (3, 5)), ("b1", (7,)), ("W2", (20, 11)), ("b2", (8,))]:
        mp = {k: v.copy() for k, v in mlp.items()}
        mm = {k: v.copy() for k, v in mlp.items()}
        mp[name][idx] += e
        mm[name][idx] -= e
        fd = (loss(mp) - loss(mm)) / (2 * e)
        assert abs(fd - an[name][idx]) < 1e-6 * max(1.0, abs(fd)), (name, idx, fd, an[name][idx])
    for (b, t, i) in [(0, 2, 1), (1, 4, 3)]:
        cp, cm = ctl.copy(), ctl.copy()
        cp[b, t, i] += e
        cm[b, t, i] -= e
        fd = (loss(mlp, cp) - loss(mlp, cm)) / (2 * e)
        assert abs(fd - gten[b, t, i]) < 1e-6 * max(1.0, abs(fd)), (b, t, i, fd, gten[b, t, i])
    assert np.all(gten[:, -1] == 0)   # the last control is never applied (knode.py:102)


@pytest.mark.parametrize("in_dim,H", [(28, 64), (53, 24)])
def test_bptt_fp32_vs_fp64_and_history_inputs(in_dim, H):
    """fp32 kernel against the fp64 kernel at the north-star bar: 1e-4 of each gradient's scale (the SIMT kernels take the
    shooting Jacobian by central differences and measure 1e-5..3e-5; the tensor-core kernel's Jacobian is exact: 1e-6)."""
    P, ctl, mlp, Cw = setup_case(H=H, B=5, T=12, in_dim=in_dim, seed=1)
    if in_dim == 53:
        P = O.setup_params(O.RodParams(), "youngs")
    _, g64 = gpu_grads(P, ctl, mlp, Cw, torch.float64)
    _, g32 = gpu_grads(P, ctl, mlp, Cw, torch.float32)
    for name, a, b in zip(("tensions",) + PK, g32, g64):
        err = np.max(np.abs(a - b)) / np.abs(b).max()
        assert err < 1e-4, (name, err)


def test_bptt_physics_only_tension_gradient():
    P, ctl, _, Cw = setup_case(B=3, T=6, seed=2)
    _, (gten, *rest) = gpu_grads(P, ctl, None, Cw, torch.float64)
    assert all(r is None for r in rest)

    def loss(c):
        return float(np.sum(Cw * O.rollout_newton(P, c, None, rows=25, tol=1e-13)))
    e = 1e-6
    for (b, t, i) in [(0, 0, 0), (2, 3, 2)]:
        cp, cm = ctl.copy(), ctl.copy()
        cp[b, t, i] += e
        cm[b, t, i] -= e
        fd = (loss(cp) - loss(cm)) / (2 * e)
        assert abs(fd - gten[b, t, i]) < 1e-6 * max(1.0, abs(fd))


def test_differentiable_rollout_api_trains():
    """robot.rollout() + loss.backward(): a few Adam steps on a rollout loss reduce it (north-star C3(ii) semantics)."""
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    from physics_controls import synthetic_tensions
    torch.manual_seed(0)
    truth = CosseratRodTorch("cuda", 32)
    setup_robot(truth)
    truth.use_nn = False
    model = CosseratRodTorch("cuda", 32)
    setup_robot(model, "youngs")          # wrong stiffness: the MLP has something to learn
    with torch.no_grad():
        model.nn_models[2].weight.mul_(0.1)
        model.nn_models[2].bias.mul_(0.1)
    tens = torch.tensor(synthetic_tensions(16, 10, truth.del_t, seed=3), device="cuda")
    with torch.no_grad():
        target = truth.rollout(tens)
    opt = torch.optim.Adam(model.nn_models.parameters(), lr=1e-3)
    losses = []
    for _ in range(6):
        traj, iters = model.rollout(tens, return_iters=True)
        assert int(iters.min()) >= 0
        loss = ((traj[:, :, :3] - target[:, :, :3]) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.nn_models.parameters())
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("tag", ["h16", "h48"])
@pytest.mark.parametrize("dt,tol", [(torch.float64, 1e-7), (torch.float32, 1e-4)])
def test_bptt_vs_reference_autograd_golden(golden, coop, tag, dt, tol):
    """FULL gradients (all of W1, b1, W2, b2 and the tensions) against tests/golden/bptt.npz = torch.autograd through the
    reference's own ODE_parallel composed into knode.simulate's loop (tests/golden/make_bptt_golden.py).  Bars: 1e-4 of each
    gradient's scale in fp32 (north star); fp64 1e-7 — the fp64 kernels take the shooting Jacobian by central differences
    (step 1e-6), whose truncation error bounds them; measured errors are ~1e-9."""
    d = golden["bptt"]
    P = O.setup_params(O.RodParams(), "youngs")
    mlp = {k: d[f"{tag}_{k}"] for k in PK}
    traj, grads = gpu_grads(P, d[tag + "_ctl"], mlp, d[tag + "_Cw"], dt)
    np.testing.assert_allclose(traj, d[tag + "_traj"], rtol=0, atol=1e-9 if dt == torch.float64 else 1e-4 * np.abs(d[tag + "_traj"]).max())
    for name, g, ref in zip(("ctl",) + PK, grads, [d[tag + "_gctl"]] + [d[f"{tag}_g{k}"] for k in PK]):
        err = np.max(np.abs(g - ref)) / np.abs(ref).max()
        assert err < tol, (name, err)
