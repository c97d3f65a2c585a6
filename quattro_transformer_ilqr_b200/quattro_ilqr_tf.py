"""``iLQR_TF`` — the reference's solver class on top of the B200 engine.

Mirror of quattro_ilqr_tf/quattro_ilqr_tf.py:50-612 of the reference: same constructor signature,
attributes (``x0, u, tol, max_iter, logs, total_iter, tf_window, state_offset, *_time``), methods
(``simulate, compute_total_cost, backward_pass, backward_pass_segment, forward_pass, optimize,
get_time, set_state_offset, get_state_offset``) and return types, so the cart-pole and quadrotor
MPC classes drop in.  Two differences, both forced by the hardware:

* the three callables are resolved to a built-in device model (``device_models.resolve``); the
  Python functions themselves are never called;
* ``x0`` may be (B, n) and ``u_init`` (B, N, m): the whole batch is then solved at once and the
  results are arrays with a leading batch axis.  With the reference's own argument types
  (``x0`` (n,), ``u_init`` list of (m,) arrays) behaviour and return types are the reference's.

Derivatives are exact instead of central finite differences, so gains differ from the reference
at its own finite-difference-noise level (SURVEY.md hard part 2; measured in tests/).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import _lib
from .device_models import resolve
from .engine import BatchedILQR

_ALPHAS = _lib.ALPHAS


class iLQR_TF:
    def __init__(self, dynamics, cost, cost_final, x0, u_init, horizon, dt=0.01, max_iter=100, tol=1e-3,
                 tf=None, tf_window=10, blend_mode=False, enable_log=True, device=None):
        self.f, self.L, self.Lf = dynamics, cost, cost_final
        self.device_model = resolve(dynamics, cost, cost_final)
        self.x0 = x0
        self.u = u_init
        self.horizon = horizon
        self.dt = dt
        self.total_iter = 0
        self.max_iter = max_iter
        self.tol = tol
        self.state_offset = np.zeros(np.asarray(x0, dtype=float).shape[-1])
        if tf_window >= horizon:
            raise ValueError("tf_window must be less than the horizon.")
        self.tf_window = tf_window
        self.tf = tf
        self.blend_mode = blend_mode          # both branches of the reference are identical (:491-543)
        self.enable_log = enable_log
        self.logs = []
        self.log_the_optimal_solution = False
        self.backward_pass_time = []
        self.forward_pass_time = []
        self.total_time = []
        self.inference_time = []
        self.device = device
        self._engine = None
        if self.tf is not None:
            if hasattr(self.tf, "prompt_len"):
                self.tf_window = self.tf.prompt_len
            if not hasattr(self.tf, "forward_device"):
                raise TypeError("tf must be a quattro_ilqr_tf.TransformerILQR (device-resident predictor); "
                                "a host-side predict() cannot be called from the GPU engine")

    # ------------------------------------------------------------------ helpers
    def _batched(self):
        return np.ndim(self.x0) == 2

    def _problem(self, u_seq=None):
        """(x0 (B,n), u (B,N,m)) as float64 arrays from the current attributes."""
        u_seq = self.u if u_seq is None else u_seq
        if self._batched():
            return np.asarray(self.x0, dtype=np.float64), np.asarray(u_seq, dtype=np.float64)
        x0 = np.asarray(self.x0, dtype=np.float64).reshape(1, -1)
        u = np.asarray(u_seq, dtype=np.float64).reshape(1, self.horizon, -1)
        return x0, u

    def _get_engine(self, B):
        e = self._engine
        if e is None or e.B != B or e.N != self.horizon or e.model is not self.device_model:
            e = BatchedILQR(self.device_model, B, self.horizon, device=self.device, tf=self.tf)
            self._engine = e
        e.tf = self.tf
        return e

    def _sync(self):
        torch.cuda.synchronize(self._engine.device)

    def _out_u(self, u):        # (B,N,m) tensor -> reference type
        u = u.cpu().numpy()
        return u if self._batched() else [u[0, t].copy() for t in range(self.horizon)]

    def _out_x(self, x):
        x = x.cpu().numpy()
        return x if self._batched() else x[0]

    def _check_tf_shapes(self):
        tf, N = self.tf, self.horizon
        # a-Q(12): the reference silently mis-aligns gains when horizon != target_len + prompt_len
        if tf.target_len + tf.prompt_len != N:
            raise ValueError(f"horizon {N} != transformer target_len {tf.target_len} + prompt_len {tf.prompt_len}")
        if (N + 1) + tf.prompt_len + tf.target_len > tf.max_seq_len:
            raise ValueError("token count exceeds the transformer's max_seq_len")

    # ------------------------------------------------------------------ reference methods
    def simulate(self, u_seq):
        x0, u = self._problem(u_seq)
        e = self._get_engine(x0.shape[0])
        e.set_problem(x0, u)
        return self._out_x(e.x_batch())

    def compute_total_cost(self, x_seq, u_seq):
        """Cost of rolling ``u_seq`` out from ``self.x0`` (x_seq is implied by the dynamics)."""
        x0, u = self._problem(u_seq)
        e = self._get_engine(x0.shape[0])
        e.set_problem(x0, u)
        c = e.cost.cpu().numpy()
        return c if self._batched() else float(c[0])

    def _gains_out(self, K, k, start):
        K, k = K.cpu().numpy().astype(np.float64), k.cpu().numpy().astype(np.float64)
        if self._batched():
            return k[:, start:], K[:, start:]
        return [k[0, t] for t in range(start, self.horizon)], [K[0, t] for t in range(start, self.horizon)]

    def backward_pass(self, x_seq, u_seq):
        """Riccati gains along the rollout of ``u_seq`` (x_seq is recomputed on the device)."""
        return self.backward_pass_segment(x_seq, u_seq, 0)

    def backward_pass_segment(self, x_seq, u_seq, start_idx):
        t0 = time.time()
        x0, u = self._problem(u_seq)
        e = self._get_engine(x0.shape[0])
        e.set_problem(x0, u)
        e.linearize(start_idx)
        e.riccati(start_idx)
        K, k = e.gains_batch()
        self._sync()
        self.backward_pass_time.append(time.time() - t0)
        return self._gains_out(K, k, start_idx)

    # ------------------------------------------------------------------ the solver
    def optimize(self, x_ref, verbose=False):
        t_total = time.time()
        x0, u0 = self._problem()
        B = x0.shape[0]
        e = self._get_engine(B)
        tf = self.tf
        if tf is not None:
            self._check_tf_shapes()
            shift = np.asarray(self.state_offset, dtype=np.float64).reshape(-1) - np.asarray(x_ref, dtype=np.float64).reshape(-1)
            e._tf_shift = torch.as_tensor(shift, dtype=torch.float64).to(e.device)     # per engine, not on the shared wrapper
        e.set_problem(x0, u0)
        start = 0 if tf is None else self.horizon - self.tf_window
        log = self.enable_log
        timed = log or not self._batched()
        for iteration in range(self.max_iter):
            if log:
                x_seq = e.x_batch()
                cur = e.cost.clone()
                was_active = (e.status == 0).cpu().numpy()
            t0 = time.time()
            e.linearize(start)
            e.riccati(start)
            if timed:
                self._sync()
                self.backward_pass_time.append(time.time() - t0)
            if tf is not None:
                t1 = time.time()
                tf.forward_device(e)
                if timed:
                    self._sync()
                    self.inference_time.append(time.time() - t1)
            t2 = time.time()
            e.linesearch(self.tol)
            n_active = int(e.n_active.item())
            self.forward_pass_time.append(time.time() - t2)
            if log:
                self._append_log(e, iteration, x_seq, cur, was_active, start)
            self.total_iter = iteration
            if verbose:
                a = e.alpha.cpu().numpy()
                c = e.cost.cpu().numpy()
                print(f"Iteration {iteration}, Cost: {c if B > 1 else c[0]:}, Alpha index: {a if B > 1 else a[0]}")
            if n_active == 0:
                break
        u_t, x_t = e.u_batch(), e.x_batch()
        u_out, x_out = self._out_u(u_t), self._out_x(x_t)
        if self._batched():
            self.total_iter = e.iters.cpu().numpy() - 1
        self.u = u_out
        self.total_time.append(time.time() - t_total)
        return u_out, x_out

    def _append_log(self, e, iteration, x_seq, cur, was_active, start):
        K, k = e.gains_batch()
        alpha = e.alpha.cpu().numpy()
        found = alpha >= 0
        new_x, new_u, new_c = e.x_batch().cpu().numpy(), e.u_batch().cpu().numpy(), e.cost.cpu().numpy()
        cur, x_seq = cur.cpu().numpy(), x_seq.cpu().numpy()
        kk, KK = ("k_seq", "K_seq") if self.tf is None else ("k_seq_seg", "K_seq_seg")
        k_out, K_out = self._gains_out(K, k, start)
        if self._batched():
            entry = {"iteration": iteration, "x_seq": x_seq, "u_seq": new_u, "current_cost": cur, kk: k_out, KK: K_out,
                     "alpha": np.where(found, np.take(_ALPHAS, np.maximum(alpha, 0)), np.nan), "new_x_seq": new_x,
                     "new_u_seq": new_u, "new_cost": np.where(found, new_c, np.nan), "found_update": found,
                     "active": was_active}
        else:
            f = bool(found[0])
            u_list = [new_u[0, t].copy() for t in range(self.horizon)]
            entry = {"iteration": iteration, "x_seq": x_seq[0], "u_seq": u_list, "current_cost": float(cur[0]),
                     kk: k_out, KK: K_out, "alpha": _ALPHAS[int(alpha[0])] if f else None,
                     "new_x_seq": new_x[0] if f else None, "new_u_seq": u_list if f else None,
                     "new_cost": float(new_c[0]) if f else None, "found_update": f}
            if self.tf is not None:
                entry["k_seq_full"], entry["K_seq_full"] = self._gains_out(K, k, 0)
        self.logs.append(entry)

    def forward_pass(self, x_seq, u_seq, k_seq, K_seq, alpha=1.0):
        """One candidate rollout (reference :377-390): returns (new_x_seq, new_u_seq, total_cost).
        ``x_seq`` is implied by ``u_seq`` and ``self.x0`` and is recomputed on the device."""
        t0 = time.time()
        x0, u = self._problem(u_seq)
        B = x0.shape[0]
        e = self._get_engine(B)
        e.set_problem(x0, u)
        K = np.asarray(K_seq, dtype=np.float32).reshape(B, self.horizon, e.m, e.n)
        k = np.asarray(k_seq, dtype=np.float32).reshape(B, self.horizon, e.m)
        e.load_gains(K, k)
        xn, un, cost = e.forward_pass(alpha)
        cost = cost.cpu().numpy()
        self.forward_pass_time.append(time.time() - t0)
        return self._out_x(xn), self._out_u(un), (cost if self._batched() else float(cost[0]))

    def get_time(self):
        if self.tf is not None:
            return (self.total_time, self.backward_pass_time, self.forward_pass_time, self.inference_time)
        return (self.total_time, self.backward_pass_time, self.forward_pass_time)

    def set_state_offset(self, x_offset):
        self.state_offset = x_offset

    def get_state_offset(self):
        return self.state_offset.copy()
