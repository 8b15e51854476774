"""KNODE march with the MLP on tcgen05 / TMEM (csrc/kc_knode_tc.cu): forward rollout and the joint-adjoint backward.

Oracles: the reference's own KNODE rollouts (tests/golden/reference_golden.npz, produced by knode.simulate with transplanted
weights), the fp64 numpy oracle (oracle/rod_oracle.py) on seeded batches, and for gradients the fp64 kernels, which are
themselves checked against finite differences of the oracle (tests/test_gpu_bptt.py) and against the reference-composed
autograd golden (tests/golden/bptt.npz)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import rod_oracle as O

pytestmark = pytest.mark.gpu
PK = ("W1", "b1", "W2", "b2")
TOL32 = 1e-4


def field_err(a, b):
    w = 0.0
    for lo, hi in [(0, 3), (3, 7), (7, 10), (10, 13), (13, 16), (16, 19), (19, 22), (22, 25)]:
        s = max(float(np.abs(b[..., lo:hi, :]).max()), 1e-30)
        w = max(w, float(np.abs(a[..., lo:hi, :] - b[..., lo:hi, :]).max()) / s)
    return w


@pytest.fixture
def tc(monkeypatch):
    monkeypatch.delenv("KC_ROLLOUT_COOP", raising=False)
    monkeypatch.delenv("KC_ROLLOUT_MODE", raising=False)
    monkeypatch.setenv("KC_ROLLOUT_TC", "1")


@pytest.fixture(params=["8", "1"])
def rows(request, monkeypatch):
    """Both forward kernels: 8 rows per rod (Newton with a finite-difference Jacobian per joint march; small batches) and
    one row per rod (Broyden in lock step over the CTA; the launcher's choice from 8192 rods on)."""
    monkeypatch.setenv("KC_ROLLOUT_TC_ROWS", request.param)
    return request.param


@pytest.fixture
def simt(monkeypatch):
    monkeypatch.setenv("KC_ROLLOUT_TC", "0")


@pytest.mark.parametrize("N,K", [(32, 32), (128, 32), (256, 64), (32, 128), (32, 512)])
def test_umma_selftest_a_operand_in_tmem(N, K):
    """mode 3: A written to TMEM with tcgen05.st as packed bf16 pairs, '.ts' MMA, B K-major bf16 in shared memory —
    the layout the march kernels rely on for the activation tile."""
    import _kc
    if (128 + N) * K * 4 > 200 * 1024:
        pytest.skip("tile too large for the self-test's staging")
    A = torch.randint(-8, 9, (128, K), device="cuda").float()
    B = torch.randint(-8, 9, (N, K), device="cuda").float()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = _kc.lib().kc_umma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, 3,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _kc.check(rc, "kc_umma_selftest")
    torch.cuda.synchronize()
    assert torch.equal(D, A @ B.T)     # small integers: exact in bf16 operands and fp32 accumulation


@pytest.mark.parametrize("N,K", [(32, 64), (32, 128), (64, 32)])
def test_umma_selftest_b_operand_from_transposed_image(N, K):
    """mode 4: as mode 3, but B is staged as the K-major image of B^T ([K rows][N]) and read as an MN-major operand with
    LBO = (N/8)*128 (next 8 k) and SBO = 128 (next 8 n): one weight image then serves the GEMM that contracts over its rows
    and the GEMM that contracts over its columns."""
    import _kc
    A = torch.randint(-8, 9, (128, K), device="cuda").float()
    B = torch.randint(-8, 9, (N, K), device="cuda").float()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = _kc.lib().kc_umma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, 4,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _kc.check(rc, "kc_umma_selftest")
    torch.cuda.synchronize()
    assert torch.equal(D, A @ B.T)


def _params(P):
    import _kc
    return _kc.rod_params(P)


def _dev(a, dt=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dt, device="cuda")


@pytest.mark.parametrize("tag", ["h64", "h512"])
def test_knode_rollout_tc_vs_reference(golden, tc, rows, tag):
    import _ops
    d = golden["knode_rollouts"]
    P = O.setup_params(O.RodParams(), "youngs")
    mlp = _ops.Mlp(*[_dev(d[f"{tag}_{k}"]) for k in PK])
    traj, _, iters = _ops.rollout(_params(P), mlp, _dev(d[tag + "_ctl"][None]))
    assert int(iters.min()) >= 0
    assert field_err(traj.cpu().numpy()[0].astype(np.float64), d[tag + "_traj"][:, :25]) < TOL32


def _case(H, B, T, seed):
    rng = np.random.default_rng(seed)
    P = O.setup_params(O.RodParams(), "youngs")
    ctl = np.stack([np.array(O.calc_controls("sine", 0.7 + 0.05 * b, P.del_t, T)) if b % 2 == 0
                    else 5 + 5 * rng.random((T, 4)) for b in range(B)])
    mlp = {"W1": np.abs(rng.normal(0.01, 0.01, (H, 28))), "b1": rng.normal(0, 0.01, H),
           "W2": np.abs(rng.normal(0.01, 0.01, (25, H))) * (16.0 / H) ** 0.5, "b2": rng.normal(0, 0.01, 25) * 0.3}
    return P, ctl, mlp


@pytest.mark.parametrize("H,B,T", [(16, 37, 9), (200, 16, 6), (512, 19, 8), (384, 1, 5), (64, 150, 5)])
def test_knode_rollout_tc_ragged_batches_vs_oracle(tc, rows, H, B, T):
    """Batches that do not fill a CTA (16 rods), hidden sizes that do not fill a chunk (128), one chunk / several chunks."""
    import _ops
    P, ctl, mlp = _case(H, B, T, seed=H + B)
    m = _ops.Mlp(*[_dev(mlp[k]) for k in PK])
    traj, G, iters = _ops.rollout(_params(P), m, _dev(ctl), want_G=True)
    assert int(iters.min()) >= 0
    want = O.rollout_newton(P, ctl.astype(np.float32).astype(np.float64), {k: v.astype(np.float32).astype(np.float64) for k, v in mlp.items()},
                            rows=25, tol=1e-12)
    got = traj.cpu().numpy().astype(np.float64)
    assert field_err(got, want) < TOL32
    assert np.abs(G.cpu().numpy()[:, 1:] - want[:, 1:, 7:13, 0]).max() < TOL32 * max(1.0, np.abs(want[:, :, 7:13, 0]).max())


def test_knode_rollout_tc_matches_simt_and_is_deterministic(monkeypatch):
    import _ops
    P, ctl, mlp = _case(512, 40, 12, seed=5)
    m = _ops.Mlp(*[_dev(mlp[k]) for k in PK])
    monkeypatch.setenv("KC_ROLLOUT_TC", "1")
    a, _, ia = _ops.rollout(_params(P), m, _dev(ctl))
    b, _, _ = _ops.rollout(_params(P), m, _dev(ctl))
    monkeypatch.setenv("KC_ROLLOUT_TC", "0")
    c, _, ic = _ops.rollout(_params(P), m, _dev(ctl))
    assert int(ia.min()) >= 0 and int(ic.min()) >= 0
    assert torch.equal(a, b)                       # same launch twice: bitwise identical
    assert field_err(a.cpu().numpy().astype(np.float64), c.cpu().numpy().astype(np.float64)) < TOL32


def _grads(P, ctl, mlp, Cw, dt):
    import _ops
    m = _ops.Mlp(*[_dev(mlp[k], dt) for k in PK])
    traj, _, iters = _ops.rollout(_params(P), m, _dev(ctl, dt), rows=25, tol=1e-13 if dt == torch.float64 else 0.0)
    assert int(iters.min()) >= 0
    out = _ops.rollout_bwd(_params(P), m, _dev(ctl, dt), traj, _dev(Cw, dt))
    return [g.cpu().numpy().astype(np.float64) for g in out]


@pytest.mark.parametrize("H,B,T", [(64, 5, 12), (512, 21, 7), (100, 16, 4)])
def test_bptt_tc_vs_fp64_kernel(tc, H, B, T):
    """Gradients of sum(Cw * traj) through the rollout: tensor-core fp32 joint-adjoint kernel against the fp64 kernel
    (north-star bar: 1e-4 of each gradient's scale).  The shooting Jacobian is exact here (unit-seeded adjoint marches), not
    a finite difference."""
    P, ctl, mlp = _case(H, B, T, seed=3 * H + B)
    Cw = np.random.default_rng(H).standard_normal((B, T, 25, P.N))
    g64 = _grads(P, ctl, mlp, Cw, torch.float64)
    g32 = _grads(P, ctl, mlp, Cw, torch.float32)
    for name, a, b in zip(("tensions",) + PK, g32, g64):
        err = np.max(np.abs(a - b)) / np.abs(b).max()
        assert err < TOL32, (name, err)


def test_differentiable_rollout_trains_on_tensor_cores(tc):
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    from physics_controls import synthetic_tensions
    torch.manual_seed(0)
    truth = CosseratRodTorch("cuda", 32)
    setup_robot(truth)
    truth.use_nn = False
    model = CosseratRodTorch("cuda", 128)
    setup_robot(model, "youngs")
    with torch.no_grad():
        model.nn_models[2].weight.mul_(0.05)
        model.nn_models[2].bias.mul_(0.1)
    tens = torch.tensor(synthetic_tensions(24, 10, truth.del_t, seed=3), device="cuda")
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


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_rollout_loss_matches_torch_autograd(dt):
    """kc_rollout_loss (fused 4-term loss of a rollout + its cotangent) against the same loss written with torch ops and
    differentiated by autograd (the loss lines of physics_train.py:345-352 with Utils/transformations.quaternion_to_euler)."""
    import _ops
    from Utils.transformations import quaternion_to_euler
    torch.manual_seed(4)
    B, T, N = 5, 7, 10
    key = [3, 5, 7, 9]
    traj = torch.randn(B, T, 25, N, device="cuda", dtype=dt)
    target = traj + 0.1 * torch.randn_like(traj)
    loss, g = _ops.rollout_loss(traj, target, key, scale=0.25)
    x = traj.clone().double().requires_grad_(True)
    tg = target.double()
    k = torch.tensor(key, device="cuda")
    pk, tk = x[:, 1:, :, k], tg[:, 1:, :, k]
    # (quaternion_to_euler force-casts to fp32, Utils/transformations.py:14: restate it in the tensor's own precision)
    def euler(q):
        q = q / q.norm(dim=2, keepdim=True)
        w_, x_, y_, z_ = q[:, :, 0], q[:, :, 1], q[:, :, 2], q[:, :, 3]
        roll = torch.atan2(2 * (w_ * y_ + x_ * z_), 1 - 2 * (y_ * y_ + z_ * z_))
        pitch = torch.asin(torch.clamp(2 * (w_ * z_ - x_ * y_), -1, 1))
        yaw = torch.atan2(2 * (w_ * x_ + y_ * z_), 1 - 2 * (x_ * x_ + z_ * z_))
        return torch.stack([roll, pitch, yaw], dim=2)
    per_traj = ((pk[:, :, :3] - tk[:, :, :3]) ** 2).mean(dim=(2, 3)).sum(1) + \
               ((pk[:, :, 7:19] - tk[:, :, 7:19]) ** 2).mean(dim=(2, 3)).sum(1) + \
               ((euler(pk[:, :, 3:7]) - euler(tk[:, :, 3:7])) ** 2).mean(dim=(2, 3)).sum(1) + \
               ((x[:, 1:, 19:, k - 1] - tg[:, 1:, 19:, k - 1]) ** 2).mean(dim=(2, 3)).sum(1)
    ref = 0.25 * per_traj.sum() / (T - 1)
    ref.backward()
    tol = 1e-12 if dt == torch.float64 else 2e-5
    assert abs(loss.item() - ref.item()) < tol * abs(ref.item()) * 10
    assert (g.double() - x.grad).abs().max().item() < tol * x.grad.abs().max().item() * 10
    # the euler restatement above is the reference's function
    q = torch.randn(4, 9, device="cuda")
    assert torch.allclose(euler(q.T[None, :, :, None].double())[0, :, :, 0].T.float(), quaternion_to_euler(q), atol=1e-5)


def test_bptt_trainer_reduces_the_rollout_loss(tc):
    """_train.BpttTrainer: rollout -> fused loss -> BPTT -> Adam + clamp, all library kernels."""
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    from physics_controls import synthetic_tensions
    from _train import BpttTrainer
    torch.manual_seed(0)
    truth = CosseratRodTorch("cuda", 32)
    setup_robot(truth)
    truth.use_nn = False
    model = CosseratRodTorch("cuda", 256)
    setup_robot(model, "youngs")
    with torch.no_grad():
        model.nn_models[2].weight.mul_(0.05)
        model.nn_models[2].bias.mul_(0.1)
    tens = torch.tensor(synthetic_tensions(48, 12, truth.del_t, seed=3), device="cuda")
    with torch.no_grad():
        target = truth.rollout(tens)
    tr = BpttTrainer(model, target, tens, [3, 5, 7, 9], lr=1e-3)
    losses = [tr.step() for _ in range(8)]
    assert tr.converged()
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert all((p.data >= 0).all() for n, p in model.nn_models.named_parameters() if "weight" in n)
