"""GPU parity of kc_estimate_state (csrc/kc_estimate.cu) through the C ABI: against the golden vectors of the reference's
estimate_state (knode_cosserat_realworld/estimate_state.py:158-242; tests/golden/make_estimate_state.py) and against the
numpy oracle on seeded batches.  fp64: 1e-9 of each field's scale.  fp32: 1e-4 against the oracle evaluated on the SAME
fp32-rounded measurements (the second time differences amplify the input rounding itself by 1/dt^2, which is a property
of the data, not of the kernel)."""
import os

import numpy as np
import pytest
import torch

from oracle import estimate_oracle as E
from oracle import rod_oracle as O
from test_gpu_parity import dev, field_err, params
from test_oracle_golden import EST_CASES, est_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import _kc
    import _ops
    assert torch.cuda.is_available()
    _kc.lib()
    return _ops


def run(ops, P, data, ctl, dt):
    est = ops.estimate_state(params(P), P.L, P.del_t, dev(data, dt), dev(ctl, dt))
    return est.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("case", EST_CASES)
def test_estimate_state_vs_reference(ops, golden, case):
    d = golden["estimate_state"]
    P = est_params(case)
    got = run(ops, P, d[case + "_data"][None], d[case + "_ctl"][None], torch.float64)[0]
    assert field_err(got, d[case + "_est"]) < 1e-9


def measurements(P, B, T, seed):
    """Smooth synthetic measurements (non-unit quaternions) and random tensions."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0, P.L, P.N)[None, None, :]
    t = (np.arange(T) * P.del_t)[None, :, None]
    w = 2 * np.pi / (rng.uniform(15, 40, (B, 1, 1)) * P.del_t)
    ph = rng.uniform(0, 2 * np.pi, (B, 1, 1))
    p = np.stack([0.3 * np.sin(w * t + ph) * s ** 2, 0.2 * np.cos(1.3 * w * t) * s ** 2, s + 0 * t + 0 * w], 2)
    h = np.stack([1 + 0 * s + 0 * t + 0 * w, 0.8 * s * np.sin(w * t), 0.6 * s * np.cos(0.7 * w * t + ph),
                  0.3 * s * np.sin(0.4 * w * t + 1)], 2)
    data = np.concatenate([p, h], 2) * (1 + 1e-3 * rng.standard_normal((B, T, 1, 1)))[..., :1, :]
    return data, 5 + 5 * rng.random((B, T, 4))


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("B,T,N", [(3, 77, 10), (2, 1000, 10), (1, 3, 10), (5, 26, 10), (2, 40, 20), (4, 19, 7), (2, 9, 130)])
def test_estimate_state_batches(ops, dt, B, T, N):
    """Several recordings, lengths that are not a multiple of the time tile (12 steps at N = 10), other node counts; the
    1000-step case is split into time spans that warm their recurrence carry up from zero (kc_estimate.cu)."""
    P = O.RodParams()
    P.N = N
    P.Bse = np.array([[2e-2, 1e-3, 0], [1e-3, 3e-2, 0], [0, 0, 5e-2]])     # dense: the general (non-diagonal) path
    P.compute_intermediate_terms()
    data, ctl = measurements(P, B, T, seed=B * 100 + T)
    if dt == torch.float32:
        data, ctl = data.astype(np.float32).astype(np.float64), ctl.astype(np.float32).astype(np.float64)
    got = run(ops, P, data, ctl, dt)
    want = np.stack([E.estimate_state(P, data[b], ctl[b]) for b in range(B)])
    assert got.shape == (B, T, 25, N)
    assert field_err(got, want) < (1e-9 if dt == torch.float64 else 1e-4)


def test_estimate_state_without_damping_skips_the_recurrence(ops):
    P = O.RodParams()
    P.Bbt = np.zeros((3, 3))
    P.compute_intermediate_terms()
    data, ctl = measurements(P, 2, 30, seed=1)
    want = np.stack([E.estimate_state(P, data[b], ctl[b]) for b in range(2)])
    assert field_err(run(ops, P, data, ctl, torch.float64), want) < 1e-9


def test_drop_in_function(ops, golden):
    """estimate_state(data, tensions, robot) with the reference's signature, types and side effect on robot.vstar."""
    from cosserat_ode import CosseratRod
    from estimate_state import estimate_state
    d = golden["estimate_state"]
    robot = CosseratRod()
    got = estimate_state(d["default_data"], d["default_ctl"], robot)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64 and got.shape == (60, 25, 10)
    assert field_err(got, d["default_est"]) < 1e-9
    np.testing.assert_array_equal(robot.vstar, got[0, 19:22, 0])
    # batched + device tensors in -> device tensor out
    both = torch.tensor(np.stack([d["default_data"], d["default_data"][::-1].copy()]), device="cuda")
    ctl2 = torch.tensor(np.stack([d["default_ctl"], d["default_ctl"]]), device="cuda")
    out = estimate_state(both, ctl2, CosseratRod())
    assert out.is_cuda and tuple(out.shape) == (2, 60, 25, 10)
    np.testing.assert_array_equal(out[0].cpu().numpy(), got)
    with pytest.raises(ValueError):
        estimate_state(d["default_data"][:2], d["default_ctl"][:2], CosseratRod())      # numpy.gradient needs 3 samples
    with pytest.raises(ValueError):
        estimate_state(d["n7_t12_data"], d["n7_t12_ctl"], CosseratRod())                # 7 nodes, robot.N = 10


def test_argument_checks(ops):
    import ctypes as C
    import _kc
    P = params(O.RodParams())
    x = torch.zeros((1, 4, 7, 10), device="cuda", dtype=torch.float64)
    c = torch.zeros((1, 4, 4), device="cuda", dtype=torch.float64)
    o = torch.zeros((1, 4, 25, 10), device="cuda", dtype=torch.float64)
    L = _kc.lib()
    call = lambda dtype, Lr, dtv, B, T, a, b, e: L.kc_estimate_state(dtype, C.byref(P), Lr, dtv, B, T, a, b, e, None)
    assert call(1, 0.4, 0.005, 1, 2, ops._ptr(x), ops._ptr(c), ops._ptr(o)) == -1      # T < 3
    assert call(7, 0.4, 0.005, 1, 4, ops._ptr(x), ops._ptr(c), ops._ptr(o)) == -1      # dtype
    assert call(1, 0.0, 0.005, 1, 4, ops._ptr(x), ops._ptr(c), ops._ptr(o)) == -1      # L
    assert call(1, 0.4, 0.005, 1, 4, None, ops._ptr(c), ops._ptr(o)) == -1             # NULL
    assert call(1, 0.4, 0.005, 0, 4, None, None, None) == 0                            # empty batch


def test_pipeline_simulate_estimate_save_train(tmp_path, monkeypatch):
    """The data path around training, end to end on the GPU and through the reference's on-disk format: rollout ->
    'measured' p, h -> estimate_state -> datas/<name>_estimated.npy = {"traj", "controls"} (estimate_state.py:279-280) ->
    train_segment.py loads it (train_segment.py:104-116) and trains."""
    import train_segment
    from cosserat_ode import CosseratRod
    from estimate_state import estimate_state
    from knode import simulate
    from physics_controls import calc_controls
    rod = CosseratRod(use_fsolve=True)
    T = train_segment.trim_len + 24
    monkeypatch.chdir(tmp_path)
    os.makedirs("datas")
    for name, arg in (("sin_1_0_amp_300", 1.0), ("sin_3_0_amp_300", 3.0)):
        ctl = np.array(calc_controls("sine", arg, rod.del_t, T))
        traj = simulate(rod, ctl)
        est = estimate_state(traj[:, :7, :].copy(), ctl, CosseratRod())
        assert est.shape == (T, 25, rod.N) and np.isfinite(est).all()
        # positions and quaternions pass through (the base x, y and the root's vector part are pinned)
        np.testing.assert_allclose(est[:, :3, 1:], traj[:, :3, 1:], rtol=0, atol=1e-12)
        np.testing.assert_allclose(est[:, 3:7, 1:], traj[:, 3:7, 1:], rtol=0, atol=1e-12)
        np.save(f"datas/{name}_estimated.npy", {"traj": est, "controls": ctl})
    robot, loss_arr = train_segment.main(["--data", "sinesine", "--epochs", "3", "--train_len", "20", "--layers", "64",
                                          "--save_path", str(tmp_path / "m.pth")])
    assert len(loss_arr) == 3 and np.isfinite(loss_arr).all()
    assert os.path.exists(tmp_path / "m.pth")


def test_realworld_simulate_script(tmp_path, monkeypatch, golden):
    """knode_cosserat_realworld/simulate.py as a drop-in script: recorded controls from an `_estimated.npy` dict ->
    data/<name>.npy = {"traj" [steps,50,N], "controls" [steps-1,4]} with trajectory[i] = state after controls[i]
    (simulate.py:63-100).  Checked against the oracle's rollout (class-default rod) and, with --model, that a saved
    whole-object checkpoint (physics_train.py:284-288 format) is picked up."""
    import simulate as sim_script
    from cosserat_ode import CosseratRod
    from cosserat_ode_torch import CosseratRodTorch
    from physics_controls import calc_controls
    monkeypatch.chdir(tmp_path)
    rod = CosseratRod()
    ctl = np.array(calc_controls("sine", 1.0, rod.del_t, 40))
    os.makedirs("datas")
    np.save("datas/rec_estimated.npy", {"traj": np.zeros((40, 25, 10)), "controls": ctl})
    traj, controls = sim_script.main(["--steps", "30", "--save_name", "out", "--real_data_path", "datas/rec_estimated.npy"])
    saved = np.load("data/out.npy", allow_pickle=True).item()
    assert saved["traj"].shape == (30, 50, 10) and saved["traj"].dtype == np.float64
    np.testing.assert_array_equal(saved["controls"], ctl[1:30])
    want = O.rollout_newton(O.RodParams(), np.concatenate([ctl[1:30], ctl[29:30]])[None], rows=50)[0]
    assert field_err(saved["traj"][:, :25], want[:, :25]) < 1e-9
    # with a KNODE checkpoint
    torch.manual_seed(0)
    robot = CosseratRodTorch("cuda", 32)
    with torch.no_grad():
        robot.nn_models[2].weight.mul_(0.02)
        robot.nn_models[2].bias.mul_(0.02)
    torch.save({"robot": robot}, "m.pth")
    traj_nn, _ = sim_script.main(["--steps", "12", "--model", "m.pth", "--save_name", "out_nn",
                                  "--real_data_path", "datas/rec_estimated.npy"])
    mlp = {k: v.detach().cpu().double().numpy() for k, v in zip(("W1", "b1", "W2", "b2"), robot.nn_models.parameters())}
    want_nn = O.rollout_newton(O.RodParams(), np.concatenate([ctl[1:12], ctl[11:12]])[None], mlp=mlp, rows=25)[0]
    assert field_err(traj_nn[:, :25], want_nn) < 1e-9
    assert np.abs(traj_nn[:, :25] - want[:12, :25]).max() > 1e-6          # the residual changed the rollout
