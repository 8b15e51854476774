"""Train and evaluate multiple KNODE models — drop-in for knode_cosserat/physics_multitrain.py (same CLI).

The reference fans {dataset} x {mod} x {seed} out as physics_train.py subprocesses, two at a time, and scrapes their
stdout (physics_multitrain.py:85-157); then rolls every saved model out with numpy + fsolve and prints a DTW / pos+Euler
MSE table against the physics-only baseline (:169-233).  Here the jobs are independent replicas: under torchrun each
rank takes every world-th job on its own GPU (no collective); a plain run executes them in-process one after the other.
The evaluation rollouts of one job (all eval sets) are one batched GPU launch and stay on the device: the DTW and the
pos + Euler MSE columns come from kc_eval_metrics (exact L1 DTW — fastdtw, which the reference uses, approximates it — and
scipy's as_euler('zyx') in closed form); only the two scalars per cell and the evals/*.npy dumps cross to the host.
All ranks synchronise before the evaluation (the reference joins its training subprocesses first, :152-157).
"""
import argparse
import os

import numpy as np
import torch

import _dist
import _ops
import physics_train
from cosserat_ode import CosseratRod
from knode import setup_robot, simulate
from physics_controls import calc_controls

MODS = ['nsw', 'short', 'youngs', 'lengthstiff']  # physics_multitrain.py:68-76
SPACE = 40


def build_parser():
    parser = argparse.ArgumentParser(description='Train and Evaluate Multiple Models.')
    parser.add_argument('--train', action=argparse.BooleanOptionalAction, default=True)
    parser.add_argument('--eval', action=argparse.BooleanOptionalAction, default=True)
    parser.add_argument('--original', action=argparse.BooleanOptionalAction, default=False, help="use original parameters")
    parser.add_argument('--epochs', type=int, default=1000)
    parser.add_argument('--fast', action=argparse.BooleanOptionalAction, default=False,
                        help="use fast but inaccurate training")
    parser.add_argument('--n_seeds', type=int, default=1)
    parser.add_argument('--save_dir', type=str, default='saved_models')
    return parser


def split_list(a_list):
    half = len(a_list) // 2
    return a_list[:half], a_list[half:]


def pct_error(new, old):
    if old == 0:
        return 0 if new == 0 else float('inf')
    return (new - old) / old * 100


def wait_for_all_ranks(rank, world, paths, timeout_s=24 * 3600):
    """The reference evaluates only after every training subprocess has been joined (physics_multitrain.py:152-157).
    Under torchrun the jobs of other ranks may still be running when rank 0 finishes its own: barrier through
    torch.distributed when a process group can be made, and in any case wait until every expected checkpoint exists
    (checkpoints are written atomically, physics_train.save_checkpoint)."""
    import time
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        dist.barrier()
    t0 = time.time()
    missing = [p for p in paths if not os.path.exists(p)]
    while missing and time.time() - t0 < timeout_s:
        time.sleep(1.0)
        missing = [p for p in missing if not os.path.exists(p)]
    if missing:
        raise FileNotFoundError(f"checkpoints never appeared: {missing}")


def model_path(args, data, mod, seed):
    filename = '_'.join('-'.join(s).replace('.', '_') for s in split_list(data.split(' ')))
    return f'{args.save_dir}/physics_{filename}_{mod}_trainlen_30_{args.epochs}_epoch_{seed}.pth'


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, world = _dist.world_info()
    if 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1:
        rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    if args.original:
        raise Exception("--original parameter no longer supported")
    datas = ['sine sine 0.5 1.0', 'sine sine random 0.5 1.0 0.0']   # physics_multitrain.py:43-46
    eval_set = ['sine 1.5', 'step 1.5']                                # :62-65

    jobs = [(d, m, s) for d in datas for m in MODS for s in range(args.n_seeds)]
    if args.train:
        t_start = __import__("time").time()
        for i, (data, mod, seed) in enumerate(jobs):
            if i % world != rank:
                continue
            control_type, control_arg = split_list(data.split(' '))
            argv_job = ['--verbose', '--no-eval', '--epochs', str(args.epochs), '--seed', str(seed), '--mod', mod,
                        '--save_dir', args.save_dir]
            if args.fast:
                argv_job.append('--fast')
            print(f'[rank {rank}] training {data};{mod};{seed}')
            physics_train.main(argv_job + [*control_type, *control_arg], distributed=False)

    if args.eval:
        wait_for_all_ranks(rank, world, [model_path(args, *j) for j in jobs],
                           timeout_s=24 * 3600 if (world > 1 and args.train) else 0)
        if args.train and rank == 0:   # a checkpoint older than this run would be a stale model from an earlier run
            stale = [p for p in (model_path(args, *j) for j in jobs) if os.path.getmtime(p) < t_start - 1.0]
            if stale:
                raise RuntimeError(f"stale checkpoints (not written by this run): {stale}")
    if args.eval and rank == 0:
        dev = torch.device("cuda", torch.cuda.current_device())
        robot_reference = CosseratRod(use_fsolve=True)
        setup_robot(robot_reference)
        ctl_eval = np.array([calc_controls(e.split(' ')[0], float(e.split(' ')[1]), robot_reference.del_t, 100)
                             for e in eval_set])
        ctl_eval_dev = torch.tensor(ctl_eval, device=dev)
        ref_dev = simulate(robot_reference, ctl_eval_dev, device_out=True)[:, :, :25].contiguous()  # one batched launch
        ref = ref_dev.cpu().numpy()
        print(' ' * SPACE, end='')
        for e in eval_set:
            print((';' + e + ' DTW').ljust(20), end='')
            print((';' + e + ' PQ MSE').ljust(20), end='')
        print()
        os.makedirs('evals', exist_ok=True)
        baselines = {}
        for data in [None, *datas]:
            for mod in MODS:
                for seed in range(args.n_seeds):
                    if data is None:
                        data_short = f'baseline {mod}'
                        robot = CosseratRod(use_fsolve=True, nn_path=None)
                    else:
                        data_short = f'{data} {mod} {seed}'
                        robot = CosseratRod(use_fsolve=True, nn_path=model_path(args, data, mod, seed))
                    setup_robot(robot, mod)
                    print(data_short.ljust(SPACE), end='')
                    trajs_dev = simulate(robot, ctl_eval_dev, device_out=True)
                    dtws, mses = _ops.eval_metrics(trajs_dev[:, :, :25].contiguous(), ref_dev, node=9)   # (:211-222) on the GPU
                    dtws, mses, trajs = dtws.cpu().numpy(), mses.cpu().numpy(), trajs_dev.cpu().numpy()
                    for k, evall in enumerate(eval_set):
                        trajectory, interp = trajs[k], ref[k]
                        fn = evall.replace(' ', '_') + '+' + data_short.replace(' ', '_')
                        np.save(f'evals/physics_{fn}_trainlen_30_{args.epochs}_epochs.npy',
                                {"tensions": ctl_eval[k], "reference": interp, "predicted": trajectory})
                        dtw, mse = float(dtws[k]), float(mses[k])
                        if data is None:
                            baselines[(evall, mod)] = {'dtw': dtw, 'mse': mse}
                            print(';{0:.2f}'.format(dtw).ljust(20), end='')
                            print(';{0:.2f}'.format(mse).ljust(20), end='')
                        else:
                            base = baselines[(evall, mod)]
                            print(f';{dtw:.2f} ({pct_error(dtw, base["dtw"]):+.1f}%)'.ljust(20), end='')
                            print(f';{mse:.2f} ({pct_error(mse, base["mse"]):+.1f}%)'.ljust(20), end='')
                    print()
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()                 # the other ranks leave only after rank 0 has read every checkpoint
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
