"""Train KNODE on real-robot segments — drop-in for knode_cosserat_realworld/train_segment.py (same CLI).

Data: datas/<name>_estimated.npy dicts {"traj": [T,>=25,N], "controls": [T,4]} produced by the reference's
estimate_state.py (:279-280); trim_len=100, train_len steps, key nodes [1,3,6,9], class-default rod parameters (the
reference never calls setup_robot here), Adam lr 1e-2 with weight decay, plateau scheduler, non-negative clamp
(train_segment.py:35-37,101,117-131,172,207-214).  The reference's stale constructor call `CosseratRodTorch(args.layers)`
(:101, raises TypeError) is fixed.  Each epoch is one fused GPU step (_train.py) instead of the Python loop :132-185.
`--synthetic` builds stand-in data with the GPU rollout when the recorded .npy files are not available.
"""
import argparse
import os

import numpy as np
import torch

import _dist
from _train import TeacherForcedTrainer
from cosserat_ode import CosseratRod
from cosserat_ode_torch import CosseratRodTorch
from knode import simulate
from physics_controls import calc_controls

DATA_SETS = {
    'sinesine': ['datas/sin_1_0_amp_300_estimated.npy', 'datas/sin_3_0_amp_300_estimated.npy'],
    'sinesinerand': ['datas/sin_1_0_amp_300_estimated.npy', 'datas/sin_3_0_amp_300_estimated.npy',
                     'datas/rand_0_60s_estimated.npy'],
    'sinesinestep': ['datas/sin_1_0_amp_300_estimated.npy', 'datas/sin_3_0_amp_300_estimated.npy',
                     'datas/dir_a_tension_950_estimated.npy'],
    'sinesinestepstep': ['datas/sin_1_0_amp_300_estimated.npy', 'datas/sin_3_0_amp_300_estimated.npy',
                         'datas/dir_a_tension_950_estimated.npy', 'datas/dir_a_tension_1250_estimated.npy'],
}
TRAIN = True
CLAMP_WEIGHT = True
trim_len = 100  # trim the first few steps to avoid steps without any movements (train_segment.py:36)


def build_parser():
    parser = argparse.ArgumentParser(description='Train KNODE.')
    parser.add_argument('--epochs', type=int, default=300)
    parser.add_argument('--layers', type=int, default=512)
    parser.add_argument('--weight_decay', type=float, default=1e-1)
    parser.add_argument('--train_len', type=int, default=120)
    parser.add_argument('--save_path', type=str, default="saved_models/quick_test.pth")
    parser.add_argument('--noise_traj', type=float, default=0.01)
    parser.add_argument('--noise_controls', type=float, default=0)
    parser.add_argument('--data', type=str, default='sinesine')
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--synthetic', action='store_true', help='generate stand-in data with the GPU rollout')
    return parser


def load_data(args):
    paths = DATA_SETS.get(args.data, ['datas/sin_1_0_amp_300_estimated.npy'])
    out = []
    if args.synthetic:
        rod = CosseratRod(use_fsolve=True)
        T = trim_len + args.train_len
        ctl = np.array([calc_controls('sine', 1.0 + i, rod.del_t, T) for i in range(len(paths))])
        traj = simulate(rod, ctl)
        return [(traj[i, trim_len:, :25], ctl[i, trim_len:]) for i in range(len(paths))]
    for data_path in paths:
        if not os.path.exists(data_path):
            raise FileNotFoundError(f"{data_path} not found: run the reference's prepare/estimate_state pipeline, or "
                                    "pass --synthetic")
        np_data = np.load(data_path, allow_pickle=True).item()
        out.append((np_data['traj'][trim_len:args.train_len + trim_len, :25],
                    np_data['controls'][trim_len:args.train_len + trim_len]))
    return out


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world = _dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("train_segment.py: no CUDA device (knode-cosserat_b200 has no CPU fallback)")
    device = f"cuda:{torch.cuda.current_device()}"
    torch.manual_seed(args.seed)
    torch_traj_ls, torch_controls_ls = [], []
    for traj_np, controls_np in load_data(args):
        traj = torch.tensor(traj_np).float().to(device) + torch.randn(traj_np.shape).float().to(device) * args.noise_traj
        controls = torch.tensor(controls_np).float().to(device) + \
            torch.randn(controls_np.shape).float().to(device) * args.noise_controls
        torch_traj_ls.append(traj)
        torch_controls_ls.append(controls)
    if rank == 0:
        print("Total number of trajectories: ", len(torch_traj_ls))
        print("training trajectory has shape: ", tuple(torch_traj_ls[0].shape))
        print("training control has shape: ", tuple(torch_controls_ls[0].shape))
    robot = CosseratRodTorch(device, args.layers)
    robot.use_nn = True
    trainer = TeacherForcedTrainer(robot, torch_traj_ls, torch_controls_ls, np.array([1, 3, 6, 9]), lr=1e-2,
                                   weight_decay=args.weight_decay, clamp_weight=CLAMP_WEIGHT)
    for epoch in range(args.epochs):
        total_loss = trainer.step(train=TRAIN)
        if epoch % 10 == 0 and rank == 0:
            print(f"\nEpoch {epoch} of {args.epochs}")
            print(f"\nTotal loss: {total_loss}")
        if not TRAIN:
            break
        if epoch % 50 == 0 and epoch != 0 and rank == 0:
            print("\nsaving model")
            os.makedirs(os.path.dirname(args.save_path) or '.', exist_ok=True)
            torch.save({'robot': robot, 'loss': trainer.loss_arr, 'optim': trainer.optim_state_dict()}, args.save_path)
    if args.save_path is not None and TRAIN and rank == 0:
        os.makedirs(os.path.dirname(args.save_path) or '.', exist_ok=True)
        torch.save({'robot': robot, 'loss': trainer.loss_arr, 'optim': trainer.optim_state_dict()}, args.save_path)
    trainer.close()
    return robot, trainer.loss_arr


if __name__ == "__main__":
    main()
