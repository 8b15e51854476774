"""CPU: the per-rod solver headers the kernels are built from (csrc/kc_rod.cuh, kc_rollout_core.cuh, kc_rollout_wide.cuh —
all `__host__ __device__`) compiled with g++ by the TEST-ONLY harness tests/emul and run one rod at a time against the
golden vectors of the reference.  This is the GPU-less check of the quasi-Newton / history / layout logic that the CUDA
kernels execute; the GPU parity tests remain the proof for the kernels themselves.  The harness is not linked into the
product library and not reachable from the drop-ins."""
import os
import sys

import numpy as np
import pytest

from oracle import rod_oracle as O
from test_oracle_golden import rel_field_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))


@pytest.fixture(scope="module")
def emul():
    import emul as E
    E.lib()
    return E


@pytest.mark.parametrize("mode", [0, 1, 2])      # narrow (Broyden), wide (Newton + FD Jacobian), wide + linearised correction
@pytest.mark.parametrize("name,P", [("default_sine", "default"), ("setup_random", "setup")])
def test_euler_rollout_headers_vs_reference(emul, golden, mode, name, P):
    d = golden["rollouts"]
    Pn = O.RodParams() if P == "default" else O.setup_params(O.RodParams())
    ctl = d[name + "_ctl"][:25]
    traj, iters, G = emul.rollout(Pn, ctl[None], dtype=np.float64, wide=mode)
    assert iters.min() >= 0
    assert rel_field_err(traj[0], d[name + "_traj"][:25, :25]) < 1e-9
    t32, it32, _ = emul.rollout(Pn, ctl[None], dtype=np.float32, wide=mode)
    assert it32.min() >= 0
    assert rel_field_err(t32[0].astype(np.float64), d[name + "_traj"][:25, :25]) < 1e-4


def test_rk4_rollout_headers_vs_reference(emul, golden):
    d = golden["rk4_rollouts"]
    for name in ("default_sine", "default_random"):
        traj, iters, _ = emul.rollout(O.RodParams(), d[name + "_ctl"][None], dtype=np.float64, wide=3)
        assert iters.min() >= 0
        assert rel_field_err(traj[0], d[name + "_traj"][:, :25]) < 1e-9
