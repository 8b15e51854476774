"""Pins oracle/rod_oracle.py (numpy restatement) to the golden vectors the unmodified reference produced
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import rod_oracle as O

MODS = [None, "noair", "nsw", "short", "damping", "dampstiff", "lengthstiff", "youngs"]


def P_default():
    return O.RodParams()


def P_setup(mod=None):
    return O.setup_params(O.RodParams(), mod)


def mlp_of(d, tag):
    return {k: d[f"{tag}_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}


def test_params_match_reference(golden):
    d = golden["rollouts"]
    for pre, P in (("default_params_", P_default()), ("setup_params_", P_setup())):
        for k in ["ds", "c0", "c1", "c2", "rhoA", "A", "G", "Kse", "Kbt", "Kse_plus_c0_Bse_inv",
                  "Kbt_plus_c0_Bbt_inv", "Kse_vstar", "rhoAg", "rhoJ", "tendon_dirs", "J"]:
            np.testing.assert_allclose(np.asarray(getattr(P, k), dtype=np.float64), d[pre + k], rtol=1e-15, atol=0)


@pytest.mark.parametrize("mod", MODS)
def test_ode_numpy(golden, mod):
    d = golden["ode"]
    tag = "none" if mod is None else mod
    ys, z = O.ode(P_setup(mod), d["y"], d["yh"], d["zh"], d["tf"])
    np.testing.assert_allclose(ys, d[f"np_{tag}_ys"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(z, d[f"np_{tag}_z"], rtol=1e-12, atol=1e-13)


def test_ode_default_params(golden):
    d = golden["ode"]
    ys, z = O.ode(P_default(), d["y"], d["yh"], d["zh"], d["tf"])
    np.testing.assert_allclose(ys, d["np_default_ys"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(z, d["np_default_z"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag,hist", [("h512", False), ("h64hist", True)])
def test_ode_parallel_with_mlp(golden, tag, hist):
    d = golden["ode"]
    P = P_setup()
    mlp = mlp_of(d, tag)
    ys, z = O.ode(P, d["y"], d["yh"], d["zh"], d["tf"], mlp, hist)
    np.testing.assert_allclose(ys, d[f"{tag}_par64_ys"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(z, d[f"{tag}_par64_z"], rtol=1e-11, atol=1e-12)
    # fp32 reference paths agree with the fp64 oracle to fp32 rounding (scaled by each column's magnitude)
    for key in (f"{tag}_par32_ys", f"{tag}_ode32_ys"):
        ref = d[key]
        n = ref.shape[0]
        scale = np.abs(ys[:n]).max(0) + 1e-3
        assert np.max(np.abs(ref - ys[:n]) / scale) < 5e-5
    ys0, z0 = O.ode(P, d["y"], d["yh"], d["zh"], d["tf"])
    scale = np.abs(ys0).max(0) + 1e-3
    assert np.max(np.abs(d[f"{tag}_par32_nonn_ys"] - ys0) / scale) < 5e-5


def test_march_euler_and_rk4(golden):
    d = golden["ode"]
    P = P_setup()
    res, y, z = O.march_euler(P, d["march_G"], d["march_y0"], d["march_z0"], d["march_yh"], d["march_zh"],
                              d["march_tensions"])
    np.testing.assert_allclose(res, d["march_res"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(y, d["march_y"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(z, d["march_z"], rtol=1e-12, atol=1e-14)
    res, y, z = O.march_rk4(P, d["march_G"], d["march_y0"], d["march_z0"], d["march_yh"], d["march_zh"],
                            d["march_tensions"])
    np.testing.assert_allclose(res, d["rk4_res"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(y, d["rk4_y"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(z, d["rk4_z"], rtol=1e-12, atol=1e-14)


def test_torch_residual_fp32(golden):
    """torch getResidualEuler (fp32, MLP on) vs the fp64 oracle march."""
    d = golden["ode"]
    P = P_setup()
    mlp = mlp_of(d, "h512")
    res, y, z = O.march_euler(P, d["march_G"].astype(np.float32).astype(np.float64), d["march_y0"], d["march_z0"],
                              d["march_yh"], d["march_zh"], d["march_tensions"], mlp)
    full = np.concatenate([y, z], 0)
    # column 0 of the reference's full_rod carries the OLD z[:,0]; columns j+1 carry z_new[j]
    ref = d["tres32_full"]
    scale = np.abs(full).max(1, keepdims=True) + 1e-3
    assert np.max(np.abs(ref[:19] - full[:19]) / scale[:19]) < 1e-4
    assert np.max(np.abs(ref[19:, 1:] - full[19:, :-1]) / scale[19:]) < 1e-4
    assert abs(float(d["tres32_total"]) - float(np.sum(res ** 2))) < 1e-4 * max(1.0, float(np.sum(res ** 2)))


def rel_field_err(a, b):
    """max over fields of |a-b| / max(scale_field, |b|), fields = p,h,n,m,q,w,v,u (SURVEY §7 hard parts)."""
    worst = 0.0
    for lo, hi in [(0, 3), (3, 7), (7, 10), (10, 13), (13, 16), (16, 19), (19, 22), (22, 25)]:
        scale = np.abs(b[..., lo:hi, :]).max()
        worst = max(worst, float(np.max(np.abs(a[..., lo:hi, :] - b[..., lo:hi, :]) / max(scale, 1e-30))))
    return worst


@pytest.mark.parametrize("name,P", [("default_sine", "default"), ("setup_sine", "setup"),
                                    ("setup_step", "setup"), ("setup_random", "setup")])
def test_rollout_newton_vs_simulate(golden, name, P):
    d = golden["rollouts"]
    P = P_default() if P == "default" else P_setup()
    ctl = d[name + "_ctl"]
    T = min(len(ctl), 40)
    out = O.rollout_newton(P, ctl[None, :T])[0]
    ref = d[name + "_traj"][:T]
    assert rel_field_err(out[:, :25], ref[:, :25]) < 1e-9
    np.testing.assert_allclose(out[:, 25:], ref[:, 25:], rtol=1e-8, atol=1e-7)  # yh, zh rows


@pytest.mark.parametrize("mod", MODS[1:])
def test_rollout_mods(golden, mod):
    d = golden["rollouts"]
    out = O.rollout_newton(P_setup(mod), d[f"mod_{mod}_ctl"][None])[0]
    assert rel_field_err(out[:, :25], d[f"mod_{mod}_traj"][:, :25]) < 1e-9


def test_rollout_fsolve_literal(golden):
    """The literal restatement (scipy hybrd, default xtol) reproduces the reference's own default-tolerance output to
    that solver tolerance (hybrd stops at xtol=1.49e-8; rounding-order differences move the last iterate)."""
    d = golden["rollouts"]
    out = O.rollout_fsolve(P_default(), d["default_sine_ctl"])
    assert rel_field_err(out[:, :25], d["default_sine_traj_defaulttol"][:, :25]) < 1e-7
    tight = O.rollout_fsolve(P_default(), d["default_sine_ctl"], xtol=1e-13)
    assert rel_field_err(tight[:, :25], d["default_sine_traj"][:, :25]) < 1e-10
    # SURVEY §8d golden tip position
    np.testing.assert_allclose(out[29, :3, -1], [-0.00639067, 0.02322316, 0.39909796], atol=5e-9)


@pytest.mark.parametrize("tag,hist", [("h64", False), ("h512", False), ("h32hist", True)])
def test_knode_rollout(golden, tag, hist):
    d = golden["knode_rollouts"]
    P = P_setup("youngs")
    T = 25
    out = O.rollout_newton(P, d[tag + "_ctl"][None, :T], mlp_of(d, tag), hist)[0]
    assert rel_field_err(out[:, :25], d[tag + "_traj"][:T, :25]) < 1e-8


def test_quaternion_to_euler_and_controls(golden):
    d = golden["misc"]
    eul = O.quaternion_to_euler(d["quat"])
    ok = np.ones(eul.shape[1], bool)
    ok[3] = False  # pitch = asin(1 - 1e-4): the reference's forced fp32 (.float()) loses 4 digits there
    np.testing.assert_allclose(eul[:, ok], d["euler"][:, ok], rtol=2e-6, atol=2e-6)  # ref is fp32
    np.testing.assert_allclose(eul[:, 3], d["euler"][:, 3], rtol=0, atol=5e-4)
    for k in d.files:
        if k.startswith("ctl_"):
            _, ctype, carg, dt, T = k.split("_")
            np.testing.assert_allclose(np.array(O.calc_controls(ctype, float(carg), float(dt), int(T))), d[k],
                                       rtol=0, atol=0)


def test_quaternion_vjp_matches_finite_difference():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((4, 7))
    g = rng.standard_normal((3, 7))
    an = O.quaternion_to_euler_vjp(q, g)
    fd = np.zeros_like(q)
    for i in range(4):
        e = np.zeros_like(q)
        e[i] = 1e-6
        fd[i] = np.sum(g * (O.quaternion_to_euler(q + e) - O.quaternion_to_euler(q - e)), 0) / 2e-6
    np.testing.assert_allclose(an, fd, rtol=1e-6, atol=1e-8)


def test_segment_steps(golden):
    d = golden["train"]
    traj = d["traj"].astype(np.float32).astype(np.float64)
    ctl = d["controls"].astype(np.float32).astype(np.float64)
    P = P_setup()
    mlp = {k: d[f"none_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    key = np.array([3, 5, 7, 9])
    ys, zs = traj[0, :29, :19], traj[0, :29, 19:]
    yp = np.concatenate([ys[:1], ys[:-1]])
    zp = np.concatenate([zs[:1], zs[:-1]])
    out = O.parallel_next_segment_euler(P, traj[0, 1:30], key, P.c1 * ys + P.c2 * yp, P.c1 * zs + P.c2 * zp,
                                        ctl[0, :29], mlp)
    ref = d["none_fast_grow_trajs0"]
    scale = np.abs(ref).max((0, 2), keepdims=True) + 1e-3
    assert np.max(np.abs(out - ref) / scale) < 2e-5
    # slow path, step 3 of trajectory 0 (all nodes)
    t = 3
    yh = P.c1 * traj[0, t, :19] + P.c2 * traj[0, t - 1, :19]
    zh = P.c1 * traj[0, t, 19:] + P.c2 * traj[0, t - 1, 19:]
    mlp_s = {k: d[f"slow_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    full = O.next_segment_euler(P, traj[0, t + 1], yh, zh, ctl[0, t], mlp_s)
    ref = d["slow_grow_traj_t3"]
    scale = np.abs(ref).max(1, keepdims=True) + 1e-3
    assert np.max(np.abs(full - ref) / scale) < 2e-5


@pytest.mark.parametrize("tag,mod,key,ntraj,prefix", [
    ("none", None, [3, 5, 7, 9], 3, "none_fast"), ("youngs", "youngs", [3, 5, 7, 9], 3, "youngs_fast"),
    ("slow", None, [2, 6, 9], 2, "slow"), ("segment", "classdefault", [1, 3, 6, 9], 2, "segment")])
def test_train_step_loss_and_grads(golden, tag, mod, key, ntraj, prefix):
    d = golden["train"]
    traj = d["traj"].astype(np.float32).astype(np.float64)[:ntraj]
    ctl = d["controls"].astype(np.float32).astype(np.float64)[:ntraj]
    P = P_default() if mod == "classdefault" else P_setup(mod)
    mlp = {k: d[f"{tag}_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    loss, grads, _ = O.teacher_forced_loss_and_grads(P, traj, ctl, key, mlp)
    ref_loss = float(d[f"{prefix}_loss"])
    assert abs(loss - ref_loss) < 2e-5 * abs(ref_loss)
    for k in ("W1", "b1", "W2", "b2"):
        ref = d[f"{prefix}_grad_{k}"]
        assert np.max(np.abs(grads[k] - ref)) < 1e-4 * np.abs(ref).max(), k


def test_adam_clamp_two_steps(golden):
    d = golden["train"]
    traj = d["traj"].astype(np.float32).astype(np.float64)
    ctl = d["controls"].astype(np.float32).astype(np.float64)
    P = P_setup()
    params = {k: d[f"none_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    state = {}
    for step in (1, 2):
        loss, grads, _ = O.teacher_forced_loss_and_grads(P, traj, ctl, [3, 5, 7, 9], params)
        params = O.adam_clamp_step(params, grads, state)
        for k in params:
            ref = d[f"none_fast_step{step}_{k}"]
            assert np.max(np.abs(params[k] - ref)) < 2e-4 * max(np.abs(ref).max(), 1e-2), (step, k)
    assert abs(loss - float(d["none_fast_loss2"])) < 5e-3 * abs(float(d["none_fast_loss2"]))
    # weight decay variant (train_segment.py:118-120)
    params = {k: d[f"segment_init_{k}"].astype(np.float64) for k in ("W1", "b1", "W2", "b2")}
    loss, grads, _ = O.teacher_forced_loss_and_grads(P_default(), traj[:2], ctl[:2], [1, 3, 6, 9], params)
    new = O.adam_clamp_step(params, grads, {}, weight_decay=0.1)
    for k in new:
        ref = d[f"segment_step1_{k}"]
        assert np.max(np.abs(new[k] - ref)) < 2e-4 * max(np.abs(ref).max(), 1e-2), k


def test_rollout_500_steps(golden):
    """The step count of BASELINE config 5: two 500-index rollouts of the unmodified reference (tests/golden/
    make_long_rollout.py; sine and random tensions, fsolve xtol 1e-13), every 20th index kept.  Shows that the oracle's
    Newton variant does not drift from the reference's hybrd over a long horizon."""
    d = golden["long_rollout"]
    out = O.rollout_newton(P_setup(), d["controls"], rows=25)
    assert out.shape == (2, 500, 25, 10)
    for r in range(2):
        assert rel_field_err(out[r][d["keep"]], d["traj"][r]) < 1e-9


EST_CASES = ["default", "damped", "n7_t12", "n20_t12", "n10_t3"]


def est_params(case):
    """The robots of tests/golden/make_estimate_state.py."""
    if case == "default":
        return P_default()
    if case == "damped":
        P = P_setup()
        P.Bse = np.diag([2e-2, 3e-2, 5e-2])
        P.compute_intermediate_terms()
        return P
    P = P_default()
    P.N = int(case.split("_")[0][1:])
    P.compute_intermediate_terms()
    return P


@pytest.mark.parametrize("case", EST_CASES)
def test_estimate_state_oracle(golden, case):
    """oracle/estimate_oracle.py against the unmodified reference's estimate_state (estimate_state.py:158-242), including
    node counts other than 10 (the reference's hard-coded loop index 9) and the minimum length T = 3."""
    from oracle import estimate_oracle as E
    d = golden["estimate_state"]
    got = E.estimate_state(est_params(case), d[case + "_data"], d[case + "_ctl"])
    assert rel_field_err(got, d[case + "_est"]) < 1e-11


@pytest.mark.parametrize("name", ["default_sine", "default_sine_slow", "default_random"])
def test_rollout_rk4_residual(golden, name):
    """Rollouts whose shooting residual is getResidualRK4 (cosserat_ode.py:215-255): the oracle's method="rk4" against the
    unmodified reference's simulate run with that residual (tests/golden/make_rk4_rollout.py); the Euler rollout differs
    from these vectors by ~36 % in the worst field, so the comparison discriminates."""
    d = golden["rk4_rollouts"]
    out = O.rollout_newton(P_default(), d[name + "_ctl"][None], method="rk4")[0]
    assert rel_field_err(out[:, :25], d[name + "_traj"][:, :25]) < 1e-10
    np.testing.assert_allclose(out[:, 25:], d[name + "_traj"][:, 25:], rtol=1e-8, atol=1e-7)
    if name == "default_sine":
        euler = O.rollout_newton(P_default(), d[name + "_ctl"][None])[0]
        assert rel_field_err(euler[:, :25], d[name + "_traj"][:, :25]) > 0.1


# ---- gradients through a rollout: autograd through the reference's ODE_parallel composition (make_bptt_golden.py) ----------
@pytest.mark.parametrize("tag", ["h16", "h48"])
def test_bptt_golden_rollout_and_a_gradient_entry(golden, tag):
    """The composed, differentiable rollout the golden gradients come from IS the oracle's rollout (which is pinned to
    knode.simulate), and a finite difference of the oracle reproduces golden gradient entries."""
    d = golden["bptt"]
    P = O.setup_params(O.RodParams(), "youngs")
    mlp = {k: d[f"{tag}_{k}"] for k in ("W1", "b1", "W2", "b2")}
    ctl, Cw = d[tag + "_ctl"], d[tag + "_Cw"]
    got = O.rollout_newton(P, ctl, mlp, rows=25, tol=1e-13)
    np.testing.assert_allclose(got, d[tag + "_traj"], rtol=0, atol=1e-10)

    def loss(m, c=ctl):
        return float(np.sum(Cw * O.rollout_newton(P, c, m, rows=25, tol=1e-13)))
    e = 1e-6
    for name, idx in [("W2", (20, 3)), ("b1", (5,))]:
        mp = {k: v.copy() for k, v in mlp.items()}
        mm = {k: v.copy() for k, v in mlp.items()}
        mp[name][idx] += e
        mm[name][idx] -= e
        fd = (loss(mp) - loss(mm)) / (2 * e)
        an = d[f"{tag}_g{name}"][idx]
        assert abs(fd - an) < 2e-6 * max(1.0, abs(an)), (name, fd, an)
    cp, cm = ctl.copy(), ctl.copy()
    cp[0, 1, 2] += e
    cm[0, 1, 2] -= e
    fd = (loss(mlp, cp) - loss(mlp, cm)) / (2 * e)
    assert abs(fd - d[tag + "_gctl"][0, 1, 2]) < 2e-6 * max(1.0, abs(fd))


# ---- evaluation metrics --------------------------------------------------------------------------------------------------
def test_eval_metric_oracle_pinned_to_scipy_and_to_the_vectorised_dtw(golden):
    """The 'PQ MSE' column of physics_multitrain.py:213-222 is defined through scipy's Rotation.as_euler('zyx'): the closed
    form of the oracle is pinned to scipy itself on random quaternions and on the reference's golden rollouts; the
    cell-by-cell DTW agrees bit for bit with the anti-diagonal host version the drivers used so far."""
    from scipy.spatial.transform import Rotation
    import sys as _s, os as _o
    _s.path.insert(0, _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "knode-cosserat_b200"))
    rng = np.random.default_rng(3)
    q = rng.standard_normal((500, 4))
    want = Rotation.from_quat(q[:, [1, 2, 3, 0]]).as_euler('zyx')
    np.testing.assert_allclose(O.euler_zyx(q), want, rtol=0, atol=1e-12)
    d = golden["rollouts"]
    a, b = d["setup_sine_traj"][:60, :25], d["setup_random_traj"][:60, :25]
    se_pos = (a[:, :3] - b[:, :3]).reshape((-1, 3)) ** 2
    e1 = Rotation.from_quat(a[:, 3:7].transpose((0, 2, 1)).reshape((-1, 4))[:, [1, 2, 3, 0]]).as_euler('zyx')
    e2 = Rotation.from_quat(b[:, 3:7].transpose((0, 2, 1)).reshape((-1, 4))[:, [1, 2, 3, 0]]).as_euler('zyx')
    ref_mse = np.mean(np.concatenate([(e1 - e2) ** 2, se_pos])) * 1000
    assert abs(O.pos_euler_mse(a, b) - ref_mse) < 1e-12 * ref_mse
    from _train import dtw_l1 as dtw_vec
    assert O.dtw_l1(a[:, :3, 9], b[:40, :3, 9]) == dtw_vec(a[:, :3, 9], b[:40, :3, 9])
    assert O.dtw_l1(a[:, :3, 9], a[:, :3, 9]) == 0.0
