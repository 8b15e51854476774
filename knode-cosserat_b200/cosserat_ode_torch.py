"""CosseratRodTorch — drop-in for knode_cosserat/cosserat_ode_torch.py, backed by hand-written sm_100a kernels.

Same class name, constructor, public attributes and method signatures as the reference (cosserat_ode_torch.py:5-437),
so the reference's scripts and pickled checkpoints keep working; every method that does arithmetic on the hot path
launches kernels of libknode_cosserat_b200.so through the C ABI (include/knode_cosserat.h) instead of the ~150 eager
aten launches per call of the original.  There is no CPU fallback: tensors must live on a CUDA device when a compute
method is called (constructing or unpickling the object on a CPU-only machine is fine).

Differences, all deliberate and documented in DESIGN.md:
  * `ODE`/`getNextSegmentEuler` are differentiable w.r.t. *all* tensor inputs (the reference builds R and hs_mat with
    torch.tensor(...), which silently detaches them, cosserat_ode_torch.py:158-161,185-188); parameter gradients — the
    only ones the training loops use — are identical.
  * `ODE`, `getResidualEuler` keep the dtype of their inputs instead of force-casting to fp32 (:161,333), so the class
    also works under torch.set_default_dtype(torch.float64).
  * `getResidualEuler` is not differentiable (nothing in the reference differentiates it).
"""
import numpy as np
import torch
import torch.nn as nn

import _kc
import _ops


class _ODEFunction(torch.autograd.Function):
    """kc_ode_fwd / kc_ode_bwd as one differentiable op: (y, yh, zh, tf, W1, b1, W2, b2) -> (ys, z)."""

    @staticmethod
    def forward(ctx, P, use_nn, y, yh, zh, tf, W1, b1, W2, b2):
        mlp = _ops.Mlp(W1, b1, W2, b2) if use_nn else None
        ys, z = _ops.ode_fwd(P, mlp, y, yh, zh, tf)
        ctx.P, ctx.use_nn = P, use_nn
        ctx.save_for_backward(y, yh, zh, tf, W1, b1, W2, b2)
        return ys, z

    @staticmethod
    def backward(ctx, g_ys, g_z):
        y, yh, zh, tf, W1, b1, W2, b2 = ctx.saved_tensors
        mlp = _ops.Mlp(W1, b1, W2, b2) if ctx.use_nn else None
        need_in = any(ctx.needs_input_grad[2:6])
        need_par = ctx.use_nn and any(ctx.needs_input_grad[6:10])
        out = _ops.ode_bwd(ctx.P, mlp, y, yh.to(y.dtype), zh.to(y.dtype), tf.to(y.dtype), g_ys.contiguous(),
                           g_z.contiguous(), need_inputs=need_in, need_params=need_par)
        gi = [g if (g is not None and ctx.needs_input_grad[2 + i]) else None for i, g in enumerate(out[:4])]
        gp = [g if (g is not None and ctx.needs_input_grad[6 + i]) else None for i, g in enumerate(out[4:])]
        return (None, None, *gi, *gp)


class _MLPFunction(torch.autograd.Function):
    """kc_mlp_fwd / kc_mlp_bwd: x[Q,in] -> W2 ELU(W1 x + b1) + b2."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2):
        mlp = _ops.Mlp(W1, b1, W2, b2)
        ctx.save_for_backward(x, W1, b1, W2, b2)
        return _ops.mlp_fwd(mlp, x)

    @staticmethod
    def backward(ctx, g):
        x, W1, b1, W2, b2 = ctx.saved_tensors
        out = _ops.mlp_bwd(_ops.Mlp(W1, b1, W2, b2), x, g.contiguous(), need_input=ctx.needs_input_grad[0],
                           need_params=any(ctx.needs_input_grad[1:]))
        return tuple(g_ if ctx.needs_input_grad[i] else None for i, g_ in enumerate(out))


class _RolloutFunction(torch.autograd.Function):
    """kc_rollout_fwd / kc_rollout_bwd as one differentiable op: (tensions, W1, b1, W2, b2) -> traj[B,T,25,N].
    Reverse mode = back-propagation through time with the shooting solve differentiated implicitly."""

    @staticmethod
    def forward(ctx, P, use_nn, tol, tensions, W1, b1, W2, b2):
        mlp = _ops.Mlp(W1, b1, W2, b2) if use_nn else None
        traj, _, iters = _ops.rollout(P, mlp, tensions, tol=tol, rows=25)
        ctx.P, ctx.use_nn = P, use_nn
        ctx.save_for_backward(tensions, traj, W1, b1, W2, b2)
        ctx.mark_non_differentiable(iters)
        return traj, iters

    @staticmethod
    def backward(ctx, g_traj, _g_iters):
        tensions, traj, W1, b1, W2, b2 = ctx.saved_tensors
        mlp = _ops.Mlp(W1, b1, W2, b2) if ctx.use_nn else None
        out = _ops.rollout_bwd(ctx.P, mlp, tensions, traj, g_traj.contiguous(),
                               want_g_tensions=ctx.needs_input_grad[3],
                               want_params=ctx.use_nn and any(ctx.needs_input_grad[4:8]))
        gp = [g if (g is not None and ctx.needs_input_grad[4 + i]) else None for i, g in enumerate(out[1:])]
        return (None, None, None, out[0] if ctx.needs_input_grad[3] else None, *gp)


class CosseratRodTorch:
    def __init__(self, device, n_layers, nn_input_history=False):
        self.device = device
        self.use_nn = True
        self.nn_input_history = nn_input_history
        self.verbose = False
        self.y = None  # robot state [p;h;n;m;q;w]
        self.z = None  # state related to time [v;u]
        # Parameters - Section 2.1 (cosserat_ode_torch.py:14-26)
        self.L = 0.4
        self.N = 10
        self.E = 109e9
        self.r = 0.0012
        self.rho = 8000.
        self.vstar = torch.tensor([0., 0., 1.], device=self.device)
        self.g = torch.tensor([0, 0, -9.81], device=self.device)
        self.Bse = torch.zeros((3, 3), device=self.device)
        self.Bbt = torch.diag(torch.tensor([3e-2, 3e-2, 3e-2], device=self.device))
        self.C = torch.tensor([1e-4, 1e-4, 1e-4], device=self.device)
        self.del_t = 0.005
        self.F_tip = torch.zeros(3, device=self.device)
        self.M_tip = torch.zeros(3, device=self.device)

        # tendons (:29-39)
        self.T0 = 5
        self.n_tendons = 4
        self.tendon_tensions = None
        theta = torch.tensor(torch.pi) / self.n_tendons
        self.tendon_offset = 0.02
        self.tendon_dirs = torch.tensor([
            [torch.cos(theta), torch.sin(theta), 0],
            [torch.cos(theta + torch.pi / 2), torch.sin(theta + torch.pi / 2), 0],
            [torch.cos(theta + torch.pi), torch.sin(theta + torch.pi), 0],
            [torch.cos(theta + 3 * torch.pi / 2), torch.sin(theta + 3 * torch.pi / 2), 0]
        ]).to(self.device)

        # Boundary Conditions - Section 2.4 (:42-45)
        self.p0 = torch.zeros(3).to(self.device)
        self.h0 = torch.tensor([1., 0., 0., 0.]).to(self.device)
        self.q0 = torch.zeros(3).to(self.device)
        self.w0 = torch.zeros(3).to(self.device)

        self.compute_intermediate_terms()

        self.residualArgs = {"yh": None, "zh": None, "tendon_forces": None}

        # KNODE residual MLP (:60-62) and its init (:76-84)
        self.layers = [nn.Linear(53 if self.nn_input_history else 28, n_layers),
                       nn.ELU(),
                       nn.Linear(n_layers, 25)]
        for i in range(len(self.layers)):
            if str(self.layers[i])[:6] == 'Linear':
                self.non_negative_normal_init(self.layers[i], mean=0.01, std=0.01)
                nn.init.normal_(self.layers[i].bias, mean=0.0, std=0.01)
        self.nn_models = nn.ModuleList(self.layers).to(self.device)

    def non_negative_normal_init(self, m, mean, std):
        """|N(mean, std)| weights (cosserat_ode_torch.py:90-105)."""
        if isinstance(m, nn.Linear) or isinstance(m, nn.Conv2d):
            assert mean >= 0, "Mean must be non-negative"
            with torch.no_grad():
                m.weight.data.normal_(mean, std).abs_()

    def __getstate__(self):
        """Checkpoints pickle the whole object (physics_train.py:284-288): keep it picklable (drop the ctypes cache)."""
        d = dict(self.__dict__)
        d.pop("_kc_params_cache", None)
        return d

    def compute_intermediate_terms(self):
        """Dependent parameters (cosserat_ode_torch.py:108-129).  One-off host-side scalar/3x3 setup: the kernels
        receive these through kc_rod_params, refreshed on every call."""
        self.A = torch.pi * self.r ** 2
        self.G = self.E / (2 * (1 + 0.3))
        self.ds = self.L / (self.N - 1)
        self.J = torch.diag(torch.tensor([torch.pi * self.r ** 4 / 4, torch.pi * self.r ** 4 / 4,
                                          torch.pi * self.r ** 4 / 2])).to(self.device)
        self.Kse = torch.diag(torch.tensor([self.G * self.A, self.G * self.A, self.E * self.A])).to(self.device)
        self.Kbt = torch.diag(torch.tensor([self.E * self.J[0, 0], self.E * self.J[1, 1],
                                            self.G * self.J[2, 2]])).to(self.device)
        self.c0 = 1.5 / self.del_t
        self.c1 = -2 / self.del_t
        self.c2 = 0.5 / self.del_t
        self.Kse_plus_c0_Bse_inv = torch.inverse(self.Kse + self.c0 * self.Bse)
        self.Kbt_plus_c0_Bbt_inv = torch.inverse(self.Kbt + self.c0 * self.Bbt)
        self.Kse_vstar = torch.matmul(self.Kse, self.vstar)
        self.rhoA = self.rho * self.A
        self.rhoAg = self.rho * self.A * self.g
        self.rhoJ = self.rho * self.J

    # ------------------------------------------------------------------------------------------------------------
    # kernel plumbing
    # ------------------------------------------------------------------------------------------------------------
    def _params(self):
        return _kc.rod_params(self)

    def _weights(self):
        return (self.nn_models[0].weight, self.nn_models[0].bias, self.nn_models[2].weight, self.nn_models[2].bias)

    def _mlp(self):
        return _ops.Mlp(*self._weights()) if self.use_nn else None

    def forward(self, x):
        """MLP forward (cosserat_ode_torch.py:131-134) on kc_mlp_fwd; accepts [..., in_dim]."""
        shape = x.shape[:-1]
        out = _MLPFunction.apply(x.reshape(-1, x.shape[-1]).contiguous(), *self._weights())
        return out.reshape(*shape, 25)

    def ODE(self, y, yh, zh, tendon_forces):
        """One node (cosserat_ode_torch.py:137-214): [19],[19],[6],[3] -> ([19],[6])."""
        ys, z = self.ODE_parallel(y.unsqueeze(0), yh.unsqueeze(0), zh.unsqueeze(0), tendon_forces.unsqueeze(0))
        return ys[0], z[0]

    def ODE_parallel(self, ys, yhs, zhs, tendon_forcess):
        """Q nodes at once (cosserat_ode_torch.py:217-322): [Q,19],[Q,19],[Q,6],[Q,3] -> ([Q,19],[Q,6])."""
        dt = ys.dtype
        if self.use_nn:
            W = [w if w.dtype == dt else w.to(dt) for w in self._weights()]
        else:
            W = [None, None, None, None]
        return _ODEFunction.apply(self._params(), bool(self.use_nn), ys.contiguous(), yhs.to(dt).contiguous(),
                                  zhs.to(dt).contiguous(), tendon_forcess.to(dt).contiguous(), *W)

    def getResidualEuler(self, G):
        """Shooting residual + Euler march (cosserat_ode_torch.py:325-367): reads self.y, self.z, residualArgs, stores
        self.y; returns (sum of squared tip residuals, full_rod[25,N])."""
        y, z, yh, zh = self.y, self.z, self.residualArgs["yh"], self.residualArgs["zh"]
        with torch.no_grad():
            dt = y.dtype
            y_new = y.detach().clone().contiguous().unsqueeze(0)
            z_new = z.detach().to(dt).clone().contiguous().unsqueeze(0)
            res = _ops.march(self._params(), self._mlp(), G.detach().to(dt).reshape(1, 6), y_new, z_new,
                             yh.detach().to(dt).unsqueeze(0), zh.detach().to(dt).unsqueeze(0),
                             self.tendon_tensions.detach().to(dt).reshape(1, 4))
            y_new, z_new = y_new[0], z_new[0]
            # column 0 keeps the OLD z[:,0]; column j+1 carries the z produced at node j (:336,353)
            full_rod = torch.cat([y_new, torch.cat([z[:, :1].to(dt), z_new[:, :-1]], dim=1)], dim=0)
            self.y = y_new
            total_residual = torch.sum(res[0] ** 2)
        return total_residual, full_rod

    def getNextSegmentEuler(self, G):
        """Teacher-forced one-step prediction at every node (cosserat_ode_torch.py:370-399): G[25,N] -> [25,N]."""
        yh, zh = self.residualArgs["yh"], self.residualArgs["zh"]
        needs_grad = torch.is_grad_enabled() and (
            G.requires_grad or yh.requires_grad or zh.requires_grad
            or (self.use_nn and any(w.requires_grad for w in self._weights())))
        if not needs_grad:
            return _ops.segment_fwd(self._params(), self._mlp(), G.unsqueeze(0), None, yh.unsqueeze(0),
                                    zh.unsqueeze(0), self.tendon_tensions.reshape(1, 4))[0]
        y = G[:19, :]
        z = G[19:, :]
        tf = torch.matmul(self.tendon_tensions.to(G.dtype), self.tendon_dirs.to(G.dtype))
        Nm1 = self.N - 1
        dys, zs_new = self.ODE_parallel(y[:, :Nm1].T, yh[:, :Nm1].T, zh[:, :Nm1].T, tf.unsqueeze(0).expand(Nm1, 3))
        y_next = y[:, :Nm1].T + self.ds * dys
        first = torch.cat([y[:, 0], z[:, 0]], dim=0).unsqueeze(1)
        return torch.cat([first, torch.cat([y_next, zs_new], dim=1).T], dim=1)

    def parallelGetNextSegmentEuler(self, Gs, segment_idxs, args):
        """Batched teacher-forced prediction at key nodes (cosserat_ode_torch.py:401-437):
        Gs[S,25,N], segment_idxs[K], args{yh[S,19,N], zh[S,6,N], tendon_tensions[S,4]} -> [S,25,K]."""
        yhs, zhs, tens = args["yh"], args["zh"], args["tendon_tensions"]
        idx = np.asarray(segment_idxs.cpu() if torch.is_tensor(segment_idxs) else segment_idxs).reshape(-1)
        needs_grad = torch.is_grad_enabled() and (
            Gs.requires_grad or yhs.requires_grad or zhs.requires_grad
            or (self.use_nn and any(w.requires_grad for w in self._weights())))
        if not needs_grad:
            return _ops.segment_fwd(self._params(), self._mlp(), Gs, idx, yhs, zhs, tens)
        S, K = Gs.shape[0], idx.size
        sel = torch.as_tensor(idx - 1, device=Gs.device, dtype=torch.long)
        all_ys = Gs[:, :19, :].transpose(1, 2)[:, sel].reshape(S * K, 19)
        all_yh = yhs.transpose(1, 2)[:, sel].reshape(S * K, 19)
        all_zh = zhs.transpose(1, 2)[:, sel].reshape(S * K, 6)
        tf = torch.matmul(tens.to(Gs.dtype), self.tendon_dirs.to(Gs.dtype))
        all_tf = tf.unsqueeze(1).expand(S, K, 3).reshape(S * K, 3)
        dys, zs_new = self.ODE_parallel(all_ys, all_yh, all_zh, all_tf)
        ys_next = all_ys + self.ds * dys
        return torch.cat([ys_next, zs_new], dim=1).reshape(S, K, 25).transpose(1, 2)

    # ------------------------------------------------------------------------------------------------------------
    # differentiable time rollout (extension: the reference has no torch rollout and never back-propagates through one)
    # ------------------------------------------------------------------------------------------------------------
    def rollout(self, tensions, tol=0.0, return_iters=False):
        """Roll B rods out under tensions[B,T,4] (or [T,4]) from the straight rod: -> traj[B,T,25,N], same semantics as
        knode.simulate (index 0 = initial state).  Differentiable w.r.t. the MLP parameters and the tensions
        (kc_rollout_bwd: BPTT, implicit differentiation of the shooting solve)."""
        single = tensions.ndim == 2
        t3 = tensions.unsqueeze(0) if single else tensions
        dt = t3.dtype
        if self.use_nn:
            W = [w if w.dtype == dt else w.to(dt) for w in self._weights()]
        else:
            W = [None, None, None, None]
        traj, iters = _RolloutFunction.apply(self._params(), bool(self.use_nn), float(tol), t3.contiguous(), *W)
        if single:
            traj, iters = traj[0], iters[0]
        return (traj, iters) if return_iters else traj

    # ------------------------------------------------------------------------------------------------------------
    # fused training step (what the drop-in training scripts call instead of the Python loops)
    # ------------------------------------------------------------------------------------------------------------
    def teacher_forced_step(self, traj, controls, key_pt_idx, want_pred=False):
        """Loss and MLP gradients of one full-batch teacher-forced step over traj[B,T,25,N], controls[B,T,4]
        (physics_train.py:313-368 == :215-267 at its key nodes == train_segment.py:140-185), fused on the GPU
        (kc_train_step).  Returns (loss: float64 device scalar tensor, (gW1, gb1, gW2, gb2), pred|None)."""
        return _ops.train_step(self._params(), _ops.Mlp(*self._weights()), traj, controls, key_pt_idx, want_pred)
