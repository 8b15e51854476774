"""CosseratRod — drop-in for knode_cosserat/cosserat_ode.py (the numpy/fp64 rod the reference uses for data generation
and evaluation), backed by the fp64 instantiations of the sm_100a kernels.

Same constructor, attributes and method signatures as the reference (cosserat_ode.py:4-255).  Arrays go in and out as
numpy float64 on the host, exactly like the original, but every arithmetic method stages them to the GPU and runs the
CUDA kernels through the C ABI — there is no CPU fallback, so a CUDA device is required to *call* them (constructing the
object is host-only).  `knode.simulate` does not loop over these methods: it launches the rollout kernel directly.
"""
import numpy as np
import torch

import _kc
import _ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("knode-cosserat_b200 has no CPU fallback: a CUDA device is required")
    return torch.device("cuda", torch.cuda.current_device())


def _t(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(_dev(), dtype)


class CosseratRod:
    def __init__(self, nn_path=None, use_fsolve=False, nn_input_history=False):
        self.verbose = False
        self.use_fsolve = use_fsolve
        self.nn_path = nn_path
        self.nn_input_history = nn_input_history
        # Parameters - Section 2.1 (cosserat_ode.py:15-29)
        self.L = 0.4
        self.N = 10
        self.E = 109e9
        self.r = 0.0012
        self.rho = 8000
        self.vstar = np.array([0, 0, 1])
        self.g = np.array([0, 0, -9.81])
        self.Bse = np.zeros((3, 3))
        self.Bbt = np.diag([3e-2, 3e-2, 3e-2])
        self.C = np.array([1e-4, 1e-4, 1e-4])
        self.del_t = 0.005
        self.F_tip = np.zeros(3)
        self.M_tip = np.zeros(3)
        # tendons (:32-41)
        self.T0 = 5
        self.n_tendons = 4
        self.tendon_tensions = None
        theta = np.pi / self.n_tendons
        self.tendon_offset = 0.02
        self.tendon_dirs = np.array([
            [np.cos(theta), np.sin(theta), 0],
            [np.cos(theta + np.pi / 2), np.sin(theta + np.pi / 2), 0],
            [np.cos(theta + np.pi), np.sin(theta + np.pi), 0],
            [np.cos(theta + 3 * np.pi / 2), np.sin(theta + 3 * np.pi / 2), 0]])
        # Boundary Conditions - Section 2.4 (:44-47)
        self.p0 = np.zeros(3)
        self.h0 = np.array([1, 0, 0, 0])
        self.q0 = np.zeros(3)
        self.w0 = np.zeros(3)
        self.compute_intermediate_terms()
        if self.nn_path is not None:
            self.nn_model, self.param_ls = self.get_nn_from_file()

    def __getstate__(self):
        """Checkpoints pickle the whole object (physics_train.py:284-288): keep it picklable (drop the ctypes cache)."""
        d = dict(self.__dict__)
        d.pop("_kc_params_cache", None)
        return d

    def compute_intermediate_terms(self):
        """cosserat_ode.py:58-78 (host-side scalar / 3x3 setup, passed to the kernels through kc_rod_params)."""
        self.A = np.pi * self.r ** 2
        self.G = self.E / (2 * (1 + 0.3))
        self.ds = self.L / (self.N - 1)
        self.J = np.diag([np.pi * self.r ** 4 / 4, np.pi * self.r ** 4 / 4, np.pi * self.r ** 4 / 2])
        self.Kse = np.diag([self.G * self.A, self.G * self.A, self.E * self.A])
        self.Kbt = np.diag([self.E * self.J[0, 0], self.E * self.J[1, 1], self.G * self.J[2, 2]])
        self.c0 = 1.5 / self.del_t
        self.c1 = -2 / self.del_t
        self.c2 = 0.5 / self.del_t
        self.Kse_plus_c0_Bse_inv = np.linalg.inv(self.Kse + self.c0 * self.Bse)
        self.Kbt_plus_c0_Bbt_inv = np.linalg.inv(self.Kbt + self.c0 * self.Bbt)
        self.Kse_vstar = self.Kse @ self.vstar
        self.rhoA = self.rho * self.A
        self.rhoAg = self.rho * self.A * self.g
        self.rhoJ = self.rho * self.J

    def get_nn_from_file(self):
        """cosserat_ode.py:81-88.  The reference hard-codes map_location='mps'; any checkpoint is mapped to the host here
        and the weights are staged to the GPU per call."""
        nn_model = torch.load(self.nn_path, map_location=torch.device('cpu'), weights_only=False)['robot'].nn_models
        param_ls = []
        for _, layer in nn_model.state_dict().items():
            param_ls.append(layer.detach().cpu().numpy())
        return nn_model, param_ls

    # ------------------------------------------------------------------------------------------------------------
    def _params(self):
        return _kc.rod_params(self)

    def _mlp(self, dtype=torch.float64):
        """The transplanted MLP (param_ls = [W1, b1, W2, b2], consumed positionally as cosserat_ode.py:110-111 does)."""
        if self.nn_path is None:
            return None
        if len(self.param_ls) != 4:
            raise ValueError("only the Linear-ELU-Linear KNODE network of the reference is supported "
                             f"(got {len(self.param_ls)} parameter tensors)")
        return _ops.Mlp(*[_t(p, dtype) for p in self.param_ls])

    def get_nn_output(self, input, model, param_ls):
        """Numpy MLP inference (cosserat_ode.py:90-112) on kc_mlp_fwd; `model` must be Linear-ELU-Linear."""
        acts = [str(m) for m in model if not str(m).startswith(('Linear', 'Dropout('))]
        if len(param_ls) != 4 or acts != ['ELU(alpha=1.0)']:
            raise ValueError("only the Linear-ELU-Linear KNODE network of the reference is supported")
        x = _t(np.asarray(input, dtype=np.float64).reshape(1, -1))
        return _ops.mlp_fwd(_ops.Mlp(*[_t(p) for p in param_ls]), x)[0].cpu().numpy()

    def ODE(self, y, yh, zh, tendon_forces):
        """One node, fp64 (cosserat_ode.py:114-186): [19],[19],[6],[3] -> ([19],[6])."""
        ys, z = _ops.ode_fwd(self._params(), self._mlp(), _t(y).reshape(1, 19), _t(yh).reshape(1, 19),
                             _t(zh).reshape(1, 6), _t(tendon_forces).reshape(1, 3))
        return ys[0].cpu().numpy(), z[0].cpu().numpy()

    def _march(self, method, G, y, z, yh, zh):
        yd, zd = _t(y).unsqueeze(0).contiguous(), _t(z).unsqueeze(0).contiguous()
        res = _ops.march(self._params(), self._mlp(), _t(G).reshape(1, 6), yd, zd, _t(yh).unsqueeze(0),
                         _t(zh).unsqueeze(0), _t(self.tendon_tensions).reshape(1, 4), method)
        # the reference mutates y and z in place and its callers rely on it (knode.py:89,96)
        y[...] = yd[0].cpu().numpy()
        z[...] = zd[0].cpu().numpy()
        r = res[0].cpu().numpy()
        return r if self.use_fsolve else float(np.sum(r ** 2))

    def getResidualEuler(self, G, y, z, yh, yh_int, zh, zh_int):
        """cosserat_ode.py:188-213 (yh_int, zh_int are unused by Euler's method)."""
        return self._march(_kc.KC_MARCH_EULER, G, y, z, yh, zh)

    def getResidualRK4(self, G, y, z, yh, yh_int, zh, zh_int):
        """cosserat_ode.py:215-255; the kernel rebuilds the mid-point histories exactly as knode.py:80-81 does."""
        return self._march(_kc.KC_MARCH_RK4, G, y, z, yh, zh)
