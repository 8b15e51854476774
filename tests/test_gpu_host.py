"""GPU: the time-range form of the rollout (kc_rollout_fwd_range) and the host-buffer entry point (kc_rollout_host:
host tensions in -> host trajectory out, transfers pipelined with the solve) against the one-shot device path."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import rod_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import _kc
    import _ops
    assert torch.cuda.is_available()
    _kc.lib()
    return _ops


def _ctl(B, T, seed=0):
    from physics_controls import synthetic_tensions
    return synthetic_tensions(B, T, 0.05, seed=seed, dtype=np.float64)


def _ranges(ops, P, B, T, dt, rows, cuts, monkeypatch=None):
    import _kc
    plan = ops.RolloutPlan(P, None, B, T, dt, "cuda", rows=rows, want_G=True)
    plan.traj.fill_(float("nan"))
    tens = torch.tensor(_ctl(B, T), dtype=dt, device="cuda")
    edges = [0] + list(cuts) + [T - 1]
    for t0, t1 in zip(edges[:-1], edges[1:]):
        rc = _kc.lib().kc_rollout_fwd_range(plan.code, C.byref(P), None, B, T, ops._ptr(tens), None, None, 0.0, 0, rows,
                                            ops._ptr(plan.traj), ops._ptr(plan.G), ops._ptr(plan.iters),
                                            ops._ptr(plan.ws), plan.nbytes, t0, t1, ops._stream(plan.device))
        _kc.check(rc, "kc_rollout_fwd_range")
    torch.cuda.synchronize()
    return plan.traj.cpu().numpy(), plan.G.cpu().numpy(), plan.iters.cpu().numpy()


@pytest.mark.parametrize("mode", ["wide", "narrow"])
@pytest.mark.parametrize("dt,rows", [(torch.float32, 25), (torch.float32, 50), (torch.float64, 50)])
def test_time_ranges_reproduce_the_one_shot_rollout(ops, monkeypatch, mode, dt, rows):
    import _kc
    monkeypatch.setenv("KC_ROLLOUT_MODE", mode)
    P = _kc.rod_params(O.setup_params(O.RodParams()))
    B, T = 37, 24
    if not _kc.lib().kc_rollout_resumable(ops._DT[dt], C.byref(P), None, B, T, rows):
        pytest.skip("selected kernel runs the full range only")   # fp64 wide (no linearised variant)
    one = _ranges(ops, P, B, T, dt, rows, [])
    many = _ranges(ops, P, B, T, dt, rows, [1, 7, 8, 19])
    np.testing.assert_array_equal(one[0], many[0])      # resuming is exact, not merely within tolerance
    np.testing.assert_array_equal(one[1], many[1])
    np.testing.assert_array_equal(one[2], many[2])
    assert np.isfinite(one[0]).all() and (one[2] >= 0).all()


def test_time_range_argument_checks(ops):
    import _kc
    P = _kc.rod_params(O.setup_params(O.RodParams()))
    plan = ops.RolloutPlan(P, None, 4, 8, torch.float32, "cuda", rows=25)
    tens = torch.zeros((4, 8, 4), device="cuda")
    for t0, t1 in ((-1, 3), (5, 4), (0, 8)):
        rc = _kc.lib().kc_rollout_fwd_range(plan.code, C.byref(P), None, 4, 8, ops._ptr(tens), None, None, 0.0, 0, 25,
                                            ops._ptr(plan.traj), None, None, ops._ptr(plan.ws), plan.nbytes, t0, t1, None)
        assert rc == -1   # KC_EINVAL


@pytest.mark.parametrize("dt,rows,segments,pinned", [(torch.float32, 25, 0, True), (torch.float32, 25, 4, True),
                                                     (torch.float32, 50, 3, False), (torch.float64, 50, 5, True),
                                                     (torch.float32, 25, 64, False)])
def test_host_rollout_matches_the_device_path(ops, dt, rows, segments, pinned):
    import _kc
    P = _kc.rod_params(O.setup_params(O.RodParams()))
    B, T = 50, 21
    ctl = _ctl(B, T, seed=3)
    ref, G, it = ops.rollout(P, None, torch.tensor(ctl, dtype=dt, device="cuda"), rows=rows, want_G=True)
    hp = ops.HostRolloutPlan(P, None, B, T, dt, torch.device("cuda", 0), rows=rows, want_G=True, segments=segments)
    out = torch.full((B, T, rows, int(P.N)), float("nan"), dtype=dt)
    if pinned:
        out = out.pin_memory()
    got = hp.run(ctl, out=out)
    assert got.data_ptr() == out.data_ptr()
    np.testing.assert_array_equal(got.numpy(), ref.cpu().numpy())
    np.testing.assert_array_equal(hp.G_h.numpy(), G.cpu().numpy())
    np.testing.assert_array_equal(hp.iters_h.numpy(), it.cpu().numpy())
    # a second run into the plan's own pinned buffer (reused) gives the same answer
    again = hp.run(torch.tensor(ctl, dtype=dt))
    np.testing.assert_array_equal(again.numpy(), ref.cpu().numpy())


def test_host_rollout_against_the_oracle(ops):
    """fp64 end to end: host tensions -> host trajectory, against the numpy oracle of knode.simulate."""
    import _kc
    Pn = O.setup_params(O.RodParams())
    P = _kc.rod_params(Pn)
    B, T = 6, 12
    ctl = _ctl(B, T, seed=5)
    want = O.rollout_newton(Pn, ctl, tol=1e-13, rows=25)
    hp = ops.HostRolloutPlan(P, None, B, T, torch.float64, torch.device("cuda", 0), rows=25, segments=3)
    got = hp.run(ctl).numpy()
    scale = np.abs(want).max(axis=(0, 1, 3), keepdims=True) + 1e-12
    assert float(np.max(np.abs(got - want) / scale)) < 1e-9


def test_simulate_host_and_device_inputs_agree(ops):
    from cosserat_ode import CosseratRod
    from knode import setup_robot, simulate
    robot = CosseratRod(use_fsolve=True)
    setup_robot(robot)
    ctl = _ctl(9, 15, seed=7)
    a = simulate(robot, ctl, dtype=np.float32, rows=25)
    b = simulate(robot, torch.tensor(ctl, device="cuda"), dtype=np.float32, rows=25)
    np.testing.assert_array_equal(a, b)
    one = simulate(robot, ctl[0])                      # reference signature: [T,4] -> float64 [T,50,N]
    assert one.shape == (15, 50, robot.N) and one.dtype == np.float64
    full, G, it = simulate(robot, ctl, return_info=True)
    np.testing.assert_allclose(full[0], one, rtol=0, atol=0)
    assert G.shape == (9, 15, 6) and it.shape == (9, 15) and (it >= 0).all()


def test_host_rollout_edge_sizes_and_knode(ops, golden):
    """B = 0, T = 1 and a KNODE robot (the warp-cooperative kernel runs the full range in one piece) through the host call."""
    import _kc
    P = _kc.rod_params(O.setup_params(O.RodParams()))
    dev0 = torch.device("cuda", 0)
    hp = ops.HostRolloutPlan(P, None, 0, 5, torch.float32, dev0, rows=25)
    assert tuple(hp.run(np.zeros((0, 5, 4), np.float32)).shape) == (0, 5, 25, int(P.N))
    hp = ops.HostRolloutPlan(P, None, 3, 1, torch.float64, dev0, rows=50)
    one = hp.run(_ctl(3, 1)).numpy()
    assert one.shape == (3, 1, 50, int(P.N)) and np.isfinite(one).all()
    np.testing.assert_array_equal(one[:, 0, :25], one[:, 0, 25:])          # index 0 repeats [y; z] (knode.py:68)
    # KNODE: transplanted weights of the golden fixture, host call against the device call
    d = golden["ode"]
    mlp = ops.Mlp(*[torch.tensor(np.asarray(d[f"h512_{k}"]), dtype=torch.float32, device="cuda") * s
                    for k, s in (("W1", 1.0), ("b1", 1.0), ("W2", 0.02), ("b2", 0.02))])
    B, T = 5, 9
    ctl = _ctl(B, T, seed=11)
    ref, _, it = ops.rollout(P, mlp, torch.tensor(ctl, dtype=torch.float32, device="cuda"), rows=25)
    hp = ops.HostRolloutPlan(P, mlp, B, T, torch.float32, dev0, rows=25, segments=4)
    got = hp.run(ctl)
    assert int(it.min()) >= 0
    np.testing.assert_array_equal(got.numpy(), ref.cpu().numpy())


def test_simulate_500_steps_against_the_reference(ops, golden):
    """Drop-in call with the reference's own signature (knode.simulate(robot, ctl[T,4]) -> float64 [T,50,N]) over the 500
    time indices of BASELINE config 5, against the unmodified reference (tests/golden/make_long_rollout.py); the host path
    solves the rollout in several time ranges, so the range hand-over is crossed repeatedly."""
    from cosserat_ode import CosseratRod
    from knode import setup_robot, simulate
    d = golden["long_rollout"]
    robot = CosseratRod(use_fsolve=True)
    setup_robot(robot)
    for r in range(2):
        got = simulate(robot, d["controls"][r])
        assert got.shape == (500, 50, robot.N) and got.dtype == np.float64
        want = d["traj"][r]
        scale = np.abs(want).max(axis=(0, 2), keepdims=True) + 1e-12
        assert float(np.max(np.abs(got[d["keep"], :25] - want) / scale)) < 1e-9
