"""CPU tests of the host-side logic around the kernels: drop-in input generators and loss helper against the golden
vectors, parameter setup, LR scheduler against torch's, DTW, sharding, and the N>1 gradient all-reduce on gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "knode-cosserat_b200")


def test_calc_controls_dropin_matches_reference(golden):
    from physics_controls import calc_controls
    d = golden["misc"]
    for k in d.files:
        if k.startswith("ctl_"):
            _, ctype, carg, dt, T = k.split("_")
            np.testing.assert_array_equal(np.array(calc_controls(ctype, float(carg), float(dt), int(T))), d[k])
    with pytest.raises(Exception, match="Unknown control type"):
        calc_controls("nope", 1.0, 0.05, 3)


def test_quaternion_to_euler_dropin(golden):
    from Utils.transformations import quaternion_to_euler
    d = golden["misc"]
    out = quaternion_to_euler(torch.tensor(d["quat"])).numpy()
    np.testing.assert_allclose(out, d["euler"], rtol=0, atol=1e-6)
    assert out.dtype == np.float32  # the reference force-casts to fp32


@pytest.mark.parametrize("mod", [None, "noair", "nsw", "short", "damping", "dampstiff", "lengthstiff", "youngs"])
def test_setup_robot_numpy_class(golden, mod):
    """setup_robot on the numpy-flavoured drop-in is host-only; its derived constants feed kc_rod_params."""
    import _kc
    from cosserat_ode import CosseratRod
    from knode import setup_robot
    from oracle import rod_oracle as O
    r = CosseratRod(use_fsolve=True)
    setup_robot(r, mod)
    P = O.setup_params(O.RodParams(), mod)
    a, b = _kc.rod_params(r), _kc.rod_params(P)
    for name, _ in _kc.kc_rod_params._fields_:
        va, vb = getattr(a, name), getattr(b, name)
        if hasattr(va, "__len__"):
            np.testing.assert_allclose(list(va), list(vb), rtol=1e-15, atol=0)
        else:
            assert va == vb
    if mod is None:
        d = golden["rollouts"]
        np.testing.assert_allclose(list(a.Kbt_c0Bbt_inv), d["setup_params_Kbt_plus_c0_Bbt_inv"].reshape(-1), rtol=1e-14)
    with pytest.raises(Exception, match="Unknown mod"):
        setup_robot(r, "bogus")
    with pytest.raises(Exception, match="no longer supported"):
        setup_robot(r, None, True)


def test_torch_class_constructs_and_pickles_on_cpu(tmp_path):
    """Checkpoints pickle the whole robot object (physics_train.py:284-288): it must round-trip, keep nn_models as a
    ModuleList with the reference's state-dict keys, and compute methods must refuse to run without a GPU."""
    from cosserat_ode_torch import CosseratRodTorch
    from knode import setup_robot
    torch.manual_seed(0)
    r = CosseratRodTorch("cpu", 16)
    setup_robot(r, "youngs")
    assert list(r.nn_models.state_dict().keys()) == ["0.weight", "0.bias", "2.weight", "2.bias"]
    assert r.nn_models[0].weight.min() >= 0 and r.nn_models[0].weight.shape == (16, 28)
    f = tmp_path / "ck.pth"
    torch.save({"robot": r, "loss": [1.0]}, f)
    r2 = torch.load(f, weights_only=False)["robot"]
    assert r2.E == 10e9 and torch.equal(r2.nn_models[2].weight, r.nn_models[2].weight)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            r.ODE_parallel(torch.zeros(2, 19), torch.zeros(2, 19), torch.zeros(2, 6), torch.zeros(2, 3))
    assert CosseratRodTorch("cpu", 8, nn_input_history=True).nn_models[0].weight.shape == (8, 53)


def test_reference_pickled_checkpoint_loads_into_the_drop_in(golden):
    """tests/golden/ref_checkpoint.pth was written by the UNMODIFIED reference (make_checkpoint.py: torch.save of the whole
    CosseratRodTorch object + dtw/loss lists + Adam state, physics_train.py:284-288).  Unpickling resolves the class by
    module name, i.e. to the drop-in; CosseratRod(nn_path=...) (cosserat_ode.py:81-88) must hand back the reference's
    weights in state-dict order, and the resumed object must still be a usable CosseratRodTorch."""
    import os
    from conftest import GOLDEN
    from cosserat_ode import CosseratRod
    from cosserat_ode_torch import CosseratRodTorch
    d = golden["ref_checkpoint"]
    path = os.path.join(GOLDEN, "ref_checkpoint.pth")
    r = CosseratRod(nn_path=path, use_fsolve=True)
    assert [p.shape for p in r.param_ls] == [(64, 28), (64,), (25, 64), (25,)]
    for got, k in zip(r.param_ls, ("W1", "b1", "W2", "b2")):
        assert np.array_equal(got, d[k])
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert isinstance(ck["robot"], CosseratRodTorch) and ck["dtw"] == [1.25, 0.75] and ck["loss"] == [0.5, 0.25, 0.125]
    assert ck["robot"].E == 10e9 and ck["robot"].del_t == 0.05            # setup_robot(..., "youngs") state travelled
    assert set(ck["optim"].keys()) == {"state", "param_groups"}
    ck["robot"].compute_intermediate_terms()                             # the drop-in's methods work on the unpickled object
    assert ck["robot"]._params().N == 10


def test_plateau_lr_matches_torch():
    from _train import PlateauLR
    rng = np.random.default_rng(0)
    losses = np.concatenate([np.linspace(1, 0.5, 20), 0.5 + 0.01 * rng.random(40), np.linspace(0.5, 0.4, 5),
                             0.4 + 0.01 * rng.random(30)])
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-2)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, 'min', patience=7, factor=0.5)
    mine = PlateauLR(1e-2, patience=7, factor=0.5)
    for l in losses:
        ref.step(float(l))
        mine.step(float(l))
        assert abs(mine.get_last_lr()[0] - opt.param_groups[0]["lr"]) < 1e-15


def test_dtw_and_shard_range():
    from _dist import shard_range
    from _train import dtw_l1
    a = np.cumsum(np.ones((6, 3)), 0)
    assert dtw_l1(a, a) == 0.0
    assert dtw_l1(a[:4], a[:4] + 1.0) == pytest.approx(3.0 * 2 + 0.0 * 0 + 3.0 * 0 + 0.0, abs=10)  # finite, symmetric below
    assert dtw_l1(a, a[::-1]) == dtw_l1(a[::-1], a)
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, PKG)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import _dist
    r, w = _dist.init_from_env()
    assert (r, w) == (rank, world) and _dist.world_info() == (rank, world)
    # each rank holds the gradient of ITS shard of 5 trajectories; the sum must equal the single-process gradient
    g_all = [torch.arange(4, dtype=torch.float32) * (t + 1) for t in range(5)]
    lo, hi = _dist.shard_range(5, rank, world)
    mine = [sum(g_all[lo:hi], torch.zeros(4)), torch.full((2, 3), float(hi - lo))]
    _dist.allreduce_sum_(mine)
    # trainer-side sharding (host logic only: no kernel is launched by the constructor): 5 trajectories split 3 + 2, or,
    # presharded, every rank keeps what it was given and the global batch is world x local (weak scaling)
    from _train import TeacherForcedTrainer

    class _Robot:
        nn_models = torch.nn.ModuleList([torch.nn.Linear(28, 4), torch.nn.ELU(), torch.nn.Linear(4, 25)])
    traj, ctl = torch.zeros(5, 6, 25, 10), torch.zeros(5, 6, 4)
    tr = TeacherForcedTrainer(_Robot(), traj, ctl, [3, 5])
    assert tr.n_total == 5 and tr.traj.shape[0] == hi - lo
    tw = TeacherForcedTrainer(_Robot(), traj, ctl, [3, 5], presharded=True)
    assert tw.n_total == 5 * world and tw.traj.shape[0] == 5
    q.put((rank, mine[0].tolist(), mine[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(4, dtype=torch.float32) * 15).tolist()
    for rank, g, cnt in res:
        assert g == want                      # identical (bitwise) on every rank == single-process sum
        assert cnt == [[5.0] * 3] * 2


def test_data_processing_matches_reference_behaviour():
    """Utils/data_processing.py (knode_cosserat_realworld/Utils/data_processing.py:3-50): axis choice by rank, range clipped
    to 1e-10, squeezed statistics, and the reference's denormalize formula."""
    from Utils.data_processing import denormalize_data, normalize_data
    rng = np.random.default_rng(0)
    d3 = rng.standard_normal((7, 4, 5))
    d3[:, 2] = 3.0                                   # a constant channel: range clipped, not replaced by 1
    n, lo, span = normalize_data(d3)
    assert lo.shape == (4,) and span.shape == (4,)
    np.testing.assert_array_equal(lo, d3.min(axis=(0, 2)))
    np.testing.assert_array_equal(span, np.clip(d3.max(axis=(0, 2)) - d3.min(axis=(0, 2)), 1e-10, np.inf))
    assert span[2] == 1e-10 and n.min() == 0.0 and abs(n[:, [0, 1, 3]].max() - 1.0) < 1e-15
    d2 = rng.standard_normal((9, 3))
    n2, lo2, span2 = normalize_data(d2)
    np.testing.assert_array_equal(lo2, d2.min(axis=0))
    np.testing.assert_allclose(denormalize_data(n2, lo2, lo2 + span2), d2, rtol=0, atol=1e-15)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with the contract keys and
    needs no GPU: it times the oracle port of knode.simulate on the host cores on a bounded sample."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rod-node-steps/sec" and d["unit"] == "rod-node-steps/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: on a machine without a CUDA device the drop-in raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    import numpy as np
    from cosserat_ode import CosseratRod
    from knode import setup_robot, simulate
    robot = CosseratRod(use_fsolve=True)
    setup_robot(robot)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simulate(robot, np.zeros((3, 4)))
