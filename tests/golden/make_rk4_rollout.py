#!/usr/bin/env python
"""Golden vectors for rollouts whose shooting residual is the reference's RK4 march: the UNMODIFIED reference's
knode.simulate (knode.py:55-102, which already builds the mid-point histories yh_int / zh_int, :80-81) run on a robot whose
getResidualEuler attribute is bound to its own getResidualRK4 (cosserat_ode.py:215-255; same signature) — the one-line
switch a user of the reference would make.  fsolve tightened to xtol = 1e-13 as in make_golden.simulate_tight.

    python tests/golden/make_rk4_rollout.py        (build container only; seconds)
"""
import os
import sys

import numpy as np
import scipy.optimize

sys.path.insert(0, "/root/reference/knode_cosserat")

import knode as ref_knode  # noqa: E402
from cosserat_ode import CosseratRod  # noqa: E402
from knode import simulate  # noqa: E402
from physics_controls import calc_controls  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def simulate_rk4(robot, ctl):
    robot.getResidualEuler = robot.getResidualRK4
    orig = ref_knode.fsolve
    ref_knode.fsolve = lambda func, x0, args=(): scipy.optimize.fsolve(func, x0, args=args, xtol=1e-13)
    try:
        return simulate(robot, ctl)
    finally:
        ref_knode.fsolve = orig


def main():
    out = {}
    # class-default rod only: with the setup_robot parameters the reference's own RK4 rollout diverges at the second step
    # (NaN from time index 2 on — probably why the reference never calls getResidualRK4), so there is nothing to pin there
    for name, kind, arg, T in (("sine", "sine", 1.0, 40), ("sine_slow", "sine", 3.0, 60), ("random", "random", 2, 40)):
        robot = CosseratRod(use_fsolve=True)
        ctl = np.array(calc_controls(kind, arg, robot.del_t, T))
        out[f"default_{name}_ctl"], out[f"default_{name}_traj"] = ctl, simulate_rk4(robot, ctl)
    for k, v in out.items():
        assert np.isfinite(v).all(), k
    np.savez_compressed(os.path.join(OUT, "rk4_rollouts.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
