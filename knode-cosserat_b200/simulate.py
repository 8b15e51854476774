"""Roll a (KNODE) rod out under recorded tendon tensions — drop-in for knode_cosserat_realworld/simulate.py (same CLI,
same output file).

Reads `{"traj", "controls"}` from --real_data_path (the dict estimate_state.py writes, :279-280), builds
`CosseratRod(use_fsolve=True, nn_path=--model)` with the class-default parameters (simulate.py:30-31), rolls it out from the
straight rod for --steps time indices under `controls[1:steps]` (simulate.py:63-89) and saves
`data/<save_name>.npy = {"traj": float64 [steps,50,N], "controls": float64 [steps-1,4]}` (:94-100).  The time loop with
one fsolve per step is ONE GPU rollout here (knode.simulate -> kc_rollout_host); the reference's plots / animation
(matplotlib, PIL, Utils.visualizer: :101-125) are outside the path and not reproduced.
"""
import argparse
import os

import numpy as np

from cosserat_ode import CosseratRod
from knode import simulate


def build_parser():
    parser = argparse.ArgumentParser(description='Process some integers.')
    parser.add_argument('--steps', type=int, help='an integer number', default=100)
    parser.add_argument('--model', type=str, help='a string', default=None)
    parser.add_argument('--save_name', type=str, help='a string', default="quick_test")
    parser.add_argument('--real_data_path', type=str, help='a string',
                        default='data/real_physical/sin_1_0_amp_300_estimated.npy')
    return parser


def rollout_under_recorded_controls(robot, real_controls, steps):
    """simulate.py:56-92: trajectory[0] is the straight rod, trajectory[i] the state after real_controls[i] (i = 1 …
    steps-1).  knode.simulate returns [initial, after ctl[0], …, after ctl[-2]], so the same list is
    simulate(controls[1:steps] + one unused trailing row)."""
    real_controls = np.asarray(real_controls, dtype=np.float64)
    if steps < 1 or steps > len(real_controls):
        raise IndexError(f"index {steps - 1} is out of bounds for axis 0 with size {len(real_controls)}")
    controls = real_controls[1:steps]
    if steps == 1:
        padded = real_controls[:1]
    else:
        padded = np.concatenate([controls, controls[-1:]])
    trajectory = simulate(robot, padded)
    if steps > 1:
        robot.tendon_tensions = controls[-1].copy()     # the last tensions the reference's loop applied (:70)
    return trajectory.astype(np.float64), controls


def main(argv=None):
    args = build_parser().parse_args(argv)
    real_data = np.load(args.real_data_path, allow_pickle=True).item()
    robot = CosseratRod(use_fsolve=True, nn_path=args.model)
    trajectory, controls = rollout_under_recorded_controls(robot, real_data['controls'], args.steps)
    if not os.path.exists("data"):
        os.makedirs("data")
    np.save('data/' + args.save_name + '.npy', {"traj": trajectory, "controls": controls})
    return trajectory, controls


if __name__ == '__main__':
    main()
