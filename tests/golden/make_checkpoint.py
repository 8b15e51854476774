#!/usr/bin/env python
"""A checkpoint written exactly as the reference writes it (physics_train.py:284-288: torch.save of a dict holding the WHOLE
pickled CosseratRodTorch object, the dtw / loss lists and the optimiser state), produced with the unmodified reference
classes, plus the KNODE rollout the reference computes from those weights (numpy rod with the transplanted MLP,
physics_train.py:103-116) — SURVEY §8f-4: the drop-in CosseratRod(nn_path=...) must load it (cosserat_ode.py:81-88).

    python tests/golden/make_checkpoint.py  ->  tests/golden/ref_checkpoint.pth, tests/golden/ref_checkpoint.npz
"""
import os
import sys

import numpy as np
import scipy.optimize
import torch

REF = "/root/reference/knode_cosserat"
sys.path.insert(0, REF)
import knode as ref_knode  # noqa: E402
from cosserat_ode import CosseratRod  # noqa: E402
from cosserat_ode_torch import CosseratRodTorch  # noqa: E402
from knode import setup_robot, simulate  # noqa: E402
from physics_controls import calc_controls  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(11)
robot = CosseratRodTorch("cpu", 64)
setup_robot(robot, "youngs")
with torch.no_grad():
    robot.nn_models[2].weight.mul_(0.5)
optimizer = torch.optim.Adam(robot.nn_models.parameters(), lr=1e-2, weight_decay=0.0)     # physics_train.py:198-199
torch.save({"robot": robot, "dtw": [1.25, 0.75], "loss": [0.5, 0.25, 0.125], "optim": optimizer.state_dict()},
           os.path.join(OUT, "ref_checkpoint.pth"))

# what the reference's evaluation computes from these weights (physics_train.py:136-158): numpy rod + transplanted MLP
nr = CosseratRod(use_fsolve=True)
setup_robot(nr, "youngs")
nr.nn_model = robot.nn_models
nr.param_ls = [layer.detach().cpu().numpy() for _, layer in robot.nn_models.state_dict().items()]
nr.nn_path = "whatever"
ctl = np.array(calc_controls("sine", 0.8, nr.del_t, 25))
saved = ref_knode.fsolve
ref_knode.fsolve = lambda f, x0, args=(): scipy.optimize.fsolve(f, x0, args=args, xtol=1e-13)
try:
    traj = simulate(nr, ctl)
finally:
    ref_knode.fsolve = saved
sd = robot.nn_models.state_dict()
np.savez_compressed(os.path.join(OUT, "ref_checkpoint.npz"), ctl=ctl, traj=traj,
                    W1=sd["0.weight"].numpy(), b1=sd["0.bias"].numpy(), W2=sd["2.weight"].numpy(), b2=sd["2.bias"].numpy())
print("ref_checkpoint.pth", os.path.getsize(os.path.join(OUT, "ref_checkpoint.pth")), "bytes; traj", traj.shape)
