"""GPU parity: the CUDA path, called through the C ABI (knode-cosserat_b200/_ops.py -> libknode_cosserat_b200.so),
against the golden vectors of the reference and against the numpy oracle on seeded inputs.
Tolerances: fp64 1e-9 relative to each field's scale, fp32 1e-4 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import rod_oracle as O

pytestmark = pytest.mark.gpu

FIELDS = [(0, 3), (3, 7), (7, 10), (10, 13), (13, 16), (16, 19), (19, 22), (22, 25)]
TOL = {torch.float64: 1e-9, torch.float32: 1e-4}


def field_err(a, b):
    """max over fields of |a-b| / scale_field for [...,25,N] arrays (several fields pass through zero)."""
    worst = 0.0
    for lo, hi in FIELDS:
        if lo >= a.shape[-2]:
            break
        scale = max(float(np.abs(b[..., lo:hi, :]).max()), 1e-30)
        worst = max(worst, float(np.abs(a[..., lo:hi, :] - b[..., lo:hi, :]).max()) / scale)
    return worst


def col_err(a, b, floor=1e-3):
    scale = np.abs(b).reshape(-1, b.shape[-1]).max(0) + floor
    return float(np.max(np.abs(a - b) / scale))


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


def mlp_of(ops, d, tag, dt):
    return ops.Mlp(*[dev(d[f"{tag}_{k}"], dt) for k in ("W1", "b1", "W2", "b2")])


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("mod", [None, "noair", "nsw", "short", "damping", "dampstiff", "lengthstiff", "youngs"])
def test_ode_fwd_physics(ops, golden, dt, mod):
    d = golden["ode"]
    tag = "none" if mod is None else mod
    ys, z = ops.ode_fwd(params(P_setup(mod)), None, dev(d["y"], dt), dev(d["yh"], dt), dev(d["zh"], dt), dev(d["tf"], dt))
    assert col_err(ys.cpu().numpy().astype(np.float64), d[f"np_{tag}_ys"]) < TOL[dt]
    assert col_err(z.cpu().numpy().astype(np.float64), d[f"np_{tag}_z"]) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag", ["h512", "h64hist"])
def test_ode_fwd_knode(ops, golden, dt, tag):
    d = golden["ode"]
    ys, z = ops.ode_fwd(params(P_setup()), mlp_of(ops, d, tag, dt), dev(d["y"], dt), dev(d["yh"], dt), dev(d["zh"], dt),
                        dev(d["tf"], dt))
    assert col_err(ys.cpu().numpy().astype(np.float64), d[f"{tag}_par64_ys"]) < TOL[dt]
    assert col_err(z.cpu().numpy().astype(np.float64), d[f"{tag}_par64_z"]) < TOL[dt]


def test_ode_fwd_dense_matrices_and_empty(ops):
    """Users may set dense Kse/Kbt/Bbt/Bse/J (the diagonal fast path must not be assumed) — and Q = 0 is legal."""
    rng = np.random.default_rng(4)
    P = P_setup()
    P.Bse = 1e-3 * rng.standard_normal((3, 3))
    P.Bbt = P.Bbt + 1e-3 * rng.standard_normal((3, 3))
    P.compute_intermediate_terms()
    P.rhoJ = P.rhoJ + 1e-7 * rng.standard_normal((3, 3))
    y = rng.standard_normal((77, 19))
    y[:, 3] += 3.0
    yh, zh, tf = rng.standard_normal((77, 19)), rng.standard_normal((77, 6)), rng.standard_normal((77, 3))
    ys_o, z_o = O.ode(P, y, yh, zh, tf)
    ys, z = ops.ode_fwd(params(P), None, *[dev(a, torch.float64) for a in (y, yh, zh, tf)])
    assert col_err(ys.cpu().numpy(), ys_o) < 1e-9 and col_err(z.cpu().numpy(), z_o) < 1e-9
    e = torch.empty((0, 19), dtype=torch.float64, device="cuda")
    ys, z = ops.ode_fwd(params(P), None, e, e, e[:, :6], e[:, :3])
    assert ys.shape == (0, 19) and z.shape == (0, 6)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_march_euler_and_rk4(ops, golden, dt):
    import _kc
    d = golden["ode"]
    P = params(P_setup())
    for method, pre in ((_kc.KC_MARCH_EULER, "march"), (_kc.KC_MARCH_RK4, "rk4")):
        y, z = dev(d["march_y0"][None], dt), dev(d["march_z0"][None], dt)
        res = ops.march(P, None, dev(d["march_G"][None], dt), y, z, dev(d["march_yh"][None], dt),
                        dev(d["march_zh"][None], dt), dev(d["march_tensions"][None], dt), method)
        full = np.concatenate([y.cpu().numpy()[0], z.cpu().numpy()[0]]).astype(np.float64)
        ref = np.concatenate([d[pre + "_y"], d[pre + "_z"]])
        assert field_err(full, ref) < TOL[dt]
        assert np.max(np.abs(res.cpu().numpy()[0] - d[pre + "_res"])) < TOL[dt] * 10


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_segment_fwd(ops, golden, dt):
    d = golden["train"]
    P = P_setup()
    traj = d["traj"].astype(np.float32).astype(np.float64)
    ctl = d["controls"].astype(np.float32).astype(np.float64)
    mlp_np = {k: d[f"none_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    mlp = ops.Mlp(*[dev(mlp_np[k], dt) for k in ("W1", "b1", "W2", "b2")])
    key = np.array([3, 5, 7, 9])
    ys, zs = traj[0, :29, :19], traj[0, :29, 19:]
    yh = P.c1 * ys + P.c2 * np.concatenate([ys[:1], ys[:-1]])
    zh = P.c1 * zs + P.c2 * np.concatenate([zs[:1], zs[:-1]])
    want = O.parallel_next_segment_euler(P, traj[0, 1:30], key, yh, zh, ctl[0, :29], mlp_np)
    got = ops.segment_fwd(params(P), mlp, dev(traj[0, 1:30], dt), key, dev(yh, dt), dev(zh, dt), dev(ctl[0, :29], dt))
    assert field_err(got.cpu().numpy().astype(np.float64), want) < TOL[dt]
    if dt == torch.float32:  # the reference's own fp32 output
        assert field_err(got.cpu().numpy().astype(np.float64), d["none_fast_grow_trajs0"].astype(np.float64)) < 1e-4
    want = O.next_segment_euler(P, traj[0, 1:30], yh, zh, ctl[0, :29], mlp_np)
    got = ops.segment_fwd(params(P), mlp, dev(traj[0, 1:30], dt), None, dev(yh, dt), dev(zh, dt), dev(ctl[0, :29], dt))
    assert field_err(got.cpu().numpy().astype(np.float64), want) < TOL[dt]


@pytest.fixture(params=["narrow", "wide"])
def mode(request, monkeypatch):
    """Both rollout kernels: one rod per lane + Broyden ("narrow", large batches) and 8 lanes per rod + Newton with a
    finite-difference Jacobian per joint march ("wide", small batches).  The launcher picks by batch size; the
    environment knob forces one."""
    monkeypatch.setenv("KC_ROLLOUT_MODE", request.param)
    return request.param


ROLLS = [("default_sine", "default", None), ("default_sine200", "default", None), ("setup_sine", "setup", None),
         ("setup_step", "setup", None), ("setup_random", "setup", None)] + \
        [(f"mod_{m}", "setup", m) for m in ["noair", "nsw", "short", "damping", "dampstiff", "lengthstiff", "youngs"]]


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,kind,mod", ROLLS)
def test_rollout_vs_reference_simulate(ops, golden, mode, dt, name, kind, mod):
    d = golden["rollouts"]
    P = O.RodParams() if kind == "default" else P_setup(mod)
    ctl = d[name + "_ctl"]
    traj, G, iters = ops.rollout(params(P), None, dev(ctl[None], dt), rows=50, want_G=True)
    got = traj.cpu().numpy()[0].astype(np.float64)
    ref = d[name + "_traj"]
    assert int(iters.min()) >= 0, "a rod failed to converge"
    assert field_err(got[:, :25], ref[:, :25]) < TOL[dt]
    # rows 25:50 are the BDF2 histories (state differences scaled by 1/dt): same bar relative to their own scale
    assert field_err(got[:, 25:], ref[:, 25:]) < TOL[dt] * 10
    np.testing.assert_allclose(G.cpu().numpy()[0, 1:].astype(np.float64), ref[1:, 7:13, 0], rtol=0,
                               atol=TOL[dt] * max(1.0, np.abs(ref[:, 7:13, 0]).max()))


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("tag", ["h64", "h512", "h32hist"])
@pytest.mark.parametrize("coop", ["0", "1"])
def test_knode_rollout_vs_reference(ops, golden, mode, dt, tag, coop, monkeypatch):
    monkeypatch.setenv("KC_ROLLOUT_COOP", coop)   # 1: one rod per warp, MLP split over the lanes (narrow mode only)
    d = golden["knode_rollouts"]
    P = P_setup("youngs")
    traj, _, iters = ops.rollout(params(P), mlp_of(ops, d, tag, dt), dev(d[tag + "_ctl"][None], dt))
    assert int(iters.min()) >= 0
    assert field_err(traj.cpu().numpy()[0].astype(np.float64), d[tag + "_traj"][:, :25]) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_rollout_batch_ragged_and_edge_sizes(ops, mode, dt):
    """B not a multiple of the warp, mixed tension families, T = 1 and B = 0."""
    rng = np.random.default_rng(7)
    P = P_setup()
    B, T = 37, 14
    ctl = np.stack([np.array(O.calc_controls("sine", 0.5 + 0.1 * b, P.del_t, T)) if b % 2 == 0
                    else 5 + 5 * rng.random((T, 4)) for b in range(B)])
    want = O.rollout_newton(P, ctl, rows=25)
    traj, _, iters = ops.rollout(params(P), None, dev(ctl, dt))
    assert int(iters.min()) >= 0
    assert field_err(traj.cpu().numpy().astype(np.float64), want) < TOL[dt]
    t1, _, _ = ops.rollout(params(P), None, dev(ctl[:, :1], dt))
    assert field_err(t1.cpu().numpy().astype(np.float64), want[:, :1]) < 1e-6
    t0, _, _ = ops.rollout(params(P), None, dev(ctl[:0], dt))
    assert t0.shape == (0, T, 25, P.N)


def test_rollout_full_size_properties(ops, mode):
    """BASELINE config 2 (4096 rods x 100 steps, fp32): size-independent properties — every solve converged, the tip
    residual (free-end boundary condition) is ~0 at every step, duplicated inputs give bitwise-identical rods, and a
    sample of rods matches the fp64 oracle."""
    rng = np.random.default_rng(0)
    P = P_setup()
    B, T = 4096, 100
    ctl = np.empty((B, T, 4))
    i = np.arange(1, T + 1)[None, :, None]
    per = rng.uniform(0.5, 3.0, (B // 2, 1, 1)) / P.del_t
    ph = rng.uniform(0, 2 * np.pi, (B // 2, 1, 1))
    ctl[:B // 2] = 6 + np.sin(2 * np.pi * i / per + ph + np.arange(4)[None, None, :] * np.pi / 2)
    ctl[B // 2:] = 5 + 5 * rng.random((B // 2, T, 4))
    ctl[1] = ctl[0]
    ctl[-1] = ctl[-2]
    traj, _, iters = ops.rollout(params(P), None, dev(ctl, torch.float32))
    assert int(iters.min()) >= 0
    tr = traj.cpu().numpy()
    assert np.isfinite(tr).all()
    assert np.abs(tr[:, 1:, 7:13, -1]).max() < 5e-5          # n(L) = F_tip = 0, m(L) = M_tip = 0
    assert np.array_equal(tr[0], tr[1]) and np.array_equal(tr[-1], tr[-2])
    sel = [0, 5, B // 2 - 1, B // 2, B - 3]
    want = O.rollout_newton(P, ctl[sel], rows=25)
    assert field_err(tr[sel].astype(np.float64), want) < 1e-4


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,kind,mod", [ROLLS[0], ROLLS[2], ROLLS[4], ROLLS[-1]])
def test_rollout_plain_wide_kernel_vs_reference(ops, golden, monkeypatch, dt, name, kind, mod):
    """kc_rollout_wide_kernel itself (8 lanes per rod, Newton, NO linearised final correction): the launcher upgrades every
    small batch to wide-lin, so it is forced here with KC_ROLLOUT_LIN=0 — it is the kernel an fp64 rollout of 1 777 - 4 736
    rods gets (next test)."""
    monkeypatch.setenv("KC_ROLLOUT_MODE", "wide")
    monkeypatch.setenv("KC_ROLLOUT_LIN", "0")
    d = golden["rollouts"]
    P = O.RodParams() if kind == "default" else P_setup(mod)
    traj, G, iters = ops.rollout(params(P), None, dev(d[name + "_ctl"][None], dt), rows=50, want_G=True)
    assert int(iters.min()) >= 0
    got = traj.cpu().numpy()[0].astype(np.float64)
    assert field_err(got[:, :25], d[name + "_traj"][:, :25]) < TOL[dt]
    assert field_err(got[:, 25:], d[name + "_traj"][:, 25:]) < TOL[dt] * 10


def test_rollout_fp64_4096_rods_selects_plain_wide_kernel(ops, monkeypatch):
    """fp64 at BASELINE config 2's batch (4096 rods): the launcher's own choice here is the plain wide kernel (wide-lin's
    fp64 state only fits one wave up to 1 776 rods).  Properties at full batch + sampled rods against the fp64 oracle at
    the fp64 bar, and bitwise agreement with the same kernel forced explicitly (proves which kernel ran)."""
    for k in ("KC_ROLLOUT_MODE", "KC_ROLLOUT_LIN", "KC_ROLLOUT_COOP"):
        monkeypatch.delenv(k, raising=False)
    rng = np.random.default_rng(2)
    P = P_setup()
    B, T = 4096, 24
    ctl = np.stack([np.array(O.calc_controls("sine", 0.5 + 2.5 * rng.random(), P.del_t, T)) if b % 2 == 0
                    else 5 + 5 * rng.random((T, 4)) for b in range(B)])
    ctl[1] = ctl[0]
    traj, _, iters = ops.rollout(params(P), None, dev(ctl, torch.float64))
    assert int(iters.min()) >= 0
    tr = traj.cpu().numpy()
    assert np.isfinite(tr).all() and np.array_equal(tr[0], tr[1])
    assert np.abs(tr[:, 1:, 7:13, -1]).max() < 1e-9
    sel = [0, 7, 2048, 4095]
    assert field_err(tr[sel], O.rollout_newton(P, ctl[sel], rows=25)) < 1e-9
    monkeypatch.setenv("KC_ROLLOUT_MODE", "wide")
    monkeypatch.setenv("KC_ROLLOUT_LIN", "0")
    forced, _, _ = ops.rollout(params(P), None, dev(ctl, torch.float64))
    assert torch.equal(traj, forced)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("N,B", [(20, 40), (7, 19), (13, 33)])
def test_rollout_other_node_counts(ops, mode, dt, N, B):
    """BASELINE config 5 at reduced size: 2x the default node count (and two odd counts) — the kernels take N at run time;
    N != 10 leaves the N = 10 specialisation of the wide kernel and, beyond 10 nodes, its shared-memory budget."""
    P = O.RodParams()
    P.N = N
    setup = O.setup_params(P)          # robot of knode.setup_robot with N nodes (ds = L / (N - 1))
    T = 16
    rng = np.random.default_rng(N)
    ctl = np.stack([np.array(O.calc_controls("sine", 0.6 + 0.05 * b, setup.del_t, T)) if b % 3 else
                    5 + 5 * rng.random((T, 4)) for b in range(B)])
    want = O.rollout_newton(setup, ctl, rows=50)
    traj, G, iters = ops.rollout(params(setup), None, dev(ctl, dt), rows=50, want_G=True)
    assert int(iters.min()) >= 0
    got = traj.cpu().numpy().astype(np.float64)
    assert got.shape == (B, T, 50, N)
    assert field_err(got[:, :, :25], want[:, :, :25]) < TOL[dt]
    assert field_err(got[:, :, 25:], want[:, :, 25:]) < TOL[dt] * 10     # yh, zh = c1*x[t-1] + c2*x[t-2]: cancellation


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_rollout_500_steps_vs_reference(ops, golden, mode, dt):
    """The step count of BASELINE config 5 against the unmodified reference (tests/golden/make_long_rollout.py: two
    500-index rollouts, sine and random tensions, every 20th index kept): no drift over a long horizon."""
    d = golden["long_rollout"]
    traj, _, iters = ops.rollout(params(P_setup()), None, dev(d["controls"], dt))
    assert int(iters.min()) >= 0, "a rod failed to converge"
    got = traj.cpu().numpy().astype(np.float64)
    assert got.shape == (2, 500, 25, 10)
    assert field_err(got[:, d["keep"]], d["traj"]) < TOL[dt]


def test_rollout_config5_shard_properties(ops):
    """BASELINE config 5, one GPU's shard at full size: 8192 rods x 20 nodes x 500 time indices, class-default parameters
    (train_segment.py never calls setup_robot; dt = 0.005), fp32, 8.2 GB of trajectory.  Size-independent properties checked
    on the device — every solve converged, finite, free-end boundary condition met at every step, duplicated inputs give
    bitwise-identical rods — plus sampled rods against the fp64 oracle (first 100 steps) and against the fp64 kernel
    (all 500 steps)."""
    rng = np.random.default_rng(5)
    P = O.RodParams()
    P.N = 20
    P.compute_intermediate_terms()
    B, T = 8192, 500
    ctl = np.empty((B, T, 4), dtype=np.float32)
    i = np.arange(1, T + 1)[None, :, None]
    per = rng.uniform(0.5, 3.0, (B // 2, 1, 1)) / P.del_t
    ph = rng.uniform(0, 2 * np.pi, (B // 2, 1, 1))
    ctl[:B // 2] = 6 + np.sin(2 * np.pi * i / per + ph + np.arange(4)[None, None, :] * np.pi / 2)
    ctl[B // 2:] = 5 + 5 * rng.random((B // 2, T, 4))
    ctl[1] = ctl[0]
    ctl[-1] = ctl[-2]
    traj, _, iters = ops.rollout(params(P), None, torch.tensor(ctl, device="cuda"))
    assert tuple(traj.shape) == (B, T, 25, 20)
    assert int(iters.min()) >= 0, "a rod failed to converge"
    assert bool(torch.isfinite(traj).all())
    scale = float(traj[:, :, 7:13].abs().max())
    assert float(traj[:, 1:, 7:13, -1].abs().max()) < 1e-4 * max(scale, 1.0)      # n(L) = F_tip = 0, m(L) = M_tip = 0
    assert torch.equal(traj[0], traj[1]) and torch.equal(traj[-1], traj[-2])
    sel = [0, 7, B // 2 - 1, B // 2, B - 3]
    got = traj[sel].cpu().numpy().astype(np.float64)
    want64, _, it64 = ops.rollout(params(P), None, torch.tensor(ctl[sel], device="cuda", dtype=torch.float64))
    assert int(it64.min()) >= 0
    w64 = want64.cpu().numpy()
    assert field_err(got, w64) < 1e-4
    oracle = O.rollout_newton(P, ctl[sel][:, :100].astype(np.float64), rows=25)
    assert field_err(w64[:, :100], oracle) < 1e-9
    assert field_err(got[:, :100], oracle) < 1e-4


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["default_sine", "default_sine_slow", "default_random"])
def test_rollout_rk4_vs_reference(ops, golden, dt, name):
    """kc_rollout_fwd_rk4 (the reference's getResidualRK4 as the residual of knode.simulate) against the unmodified
    reference (tests/golden/make_rk4_rollout.py)."""
    d = golden["rk4_rollouts"]
    ref = d[name + "_traj"]
    traj, G, iters = ops.rollout(params(O.RodParams()), None, dev(d[name + "_ctl"][None], dt), rows=50, want_G=True,
                                 method="rk4")
    assert int(iters.min()) >= 0, "a rod failed to converge"
    got = traj.cpu().numpy()[0].astype(np.float64)
    assert field_err(got[:, :25], ref[:, :25]) < TOL[dt]
    assert field_err(got[:, 25:], ref[:, 25:]) < TOL[dt] * 10
    np.testing.assert_allclose(G.cpu().numpy()[0, 1:].astype(np.float64), ref[1:, 7:13, 0], rtol=0,
                               atol=TOL[dt] * max(1.0, np.abs(ref[:, 7:13, 0]).max()))


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_rollout_rk4_batch_knode_and_node_counts(ops, golden, dt):
    """RK4 rollouts of a ragged batch, with the KNODE residual in the march, and at another node count, against the oracle
    (whose RK4 rollout is pinned to the reference); the drop-in keyword knode.simulate(..., method="rk4")."""
    rng = np.random.default_rng(11)
    P = O.RodParams()
    B, T = 37, 12
    ctl = np.stack([np.array(O.calc_controls("sine", 0.5 + 0.1 * b, P.del_t, T)) if b % 2 == 0
                    else 5 + 5 * rng.random((T, 4)) for b in range(B)])
    want = O.rollout_newton(P, ctl, rows=25, method="rk4")
    traj, _, iters = ops.rollout(params(P), None, dev(ctl, dt), method="rk4")
    assert int(iters.min()) >= 0
    assert field_err(traj.cpu().numpy().astype(np.float64), want) < TOL[dt]
    # KNODE residual inside the RK4 march (weights of the golden h64 network, scaled to a trained-size correction)
    d = golden["knode_rollouts"]
    mlp_np = {k: d[f"h64_{k}"].astype(np.float64) * (0.05 if k in ("W2", "b2") else 1.0) for k in ("W1", "b1", "W2", "b2")}
    mlp = ops.Mlp(*[dev(mlp_np[k], dt) for k in ("W1", "b1", "W2", "b2")])
    want_nn = O.rollout_newton(P, ctl[:5], mlp=mlp_np, rows=25, method="rk4")
    traj_nn, _, it_nn = ops.rollout(params(P), mlp, dev(ctl[:5], dt), method="rk4")
    assert int(it_nn.min()) >= 0
    assert field_err(traj_nn.cpu().numpy().astype(np.float64), want_nn) < TOL[dt]
    assert field_err(want_nn, want[:5]) > 1e-6                       # the residual network changes the rollout
    # 13 nodes
    P13 = O.RodParams()
    P13.N = 13
    P13.compute_intermediate_terms()
    want13 = O.rollout_newton(P13, ctl[:6], rows=50, method="rk4")
    t13, _, it13 = ops.rollout(params(P13), None, dev(ctl[:6], dt), rows=50, method="rk4")
    assert int(it13.min()) >= 0
    g13 = t13.cpu().numpy().astype(np.float64)
    assert field_err(g13[:, :, :25], want13[:, :, :25]) < TOL[dt]
    assert field_err(g13[:, :, 25:], want13[:, :, 25:]) < TOL[dt] * 10
    if dt == torch.float64:
        from cosserat_ode import CosseratRod
        from knode import simulate
        one = simulate(CosseratRod(use_fsolve=True), ctl[0], method="rk4")
        assert one.shape == (T, 50, 10) and one.dtype == np.float64
        assert field_err(one[:, :25], want[0]) < 1e-9
