"""Batched iLQR engine: device buffers + the per-iteration kernel sequence.

One iteration of the reference's ``optimize`` loop (quattro_ilqr_tf/quattro_ilqr_tf.py:428-475)
for B independent trajectories is

    [rollout + cost]      only before the first iteration; afterwards the accepted candidate of the
                          line search IS the next nominal (SURVEY.md a-Q(11))
    qx_linearize          A_t, B_t, cost expansions for the whole horizon     (:182-275)
    qx_riccati            K_t, k_t                                             (:282-319 / :327-366)
    [qx_tf_forward]       transformer mode: predicted gains for t < N - P      (:489-543)
    qx_linesearch         first accepted alpha, stopping rule, next iterate    (:440-472)

PyTorch only owns memory and streams here; all arithmetic is in libquattro_sm100a.so.
Device layouts are documented in include/quattro_b200.h.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib
from .device_models import DeviceModel


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _on_device(fn):
    """Run a method with the engine's device current: kernels and cudaMemsetAsync are issued on a stream of
    ``self.device``, which is an error if another device is current (the caller need not have called set_device)."""
    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


class BatchedILQR:
    def __init__(self, model: DeviceModel, B: int, N: int, device=None, tf=None, two_phase_linesearch=True,
                 block_path=True):
        if not torch.cuda.is_available():
            raise RuntimeError("quattro_transformer_ilqr_b200 needs a CUDA device (B200, sm_100a); no CPU fallback")
        self.lib = _lib.load()
        self.model, self.B, self.N = model, int(B), int(N)
        self.n, self.m = model.n, model.m
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("quattro_transformer_ilqr_b200 needs a CUDA device (B200, sm_100a); no CPU fallback")
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.tf = tf
        n, m, mid = self.n, self.m, model.model_id
        dev, f64, f32, i32 = self.device, torch.float64, torch.float32, torch.int32
        self.layout = self.lib.qx_layout(mid)
        self.rec = self.lib.qx_record_floats(mid)
        self.x = [torch.zeros((N + 1, n, B), dtype=f64, device=dev) for _ in range(2)]
        self.u = [torch.zeros((N, m, B), dtype=f64, device=dev) for _ in range(2)]
        self.x0 = torch.zeros((n, B), dtype=f64, device=dev)
        self.cost = torch.zeros(B, dtype=f64, device=dev)
        self.status = torch.zeros(B, dtype=i32, device=dev)
        self.iters = torch.zeros(B, dtype=i32, device=dev)
        self.alpha = torch.full((B,), -1, dtype=i32, device=dev)
        self.n_active = torch.zeros(1, dtype=i32, device=dev)
        # expansion / terminal are allocated by _select_path(): compact block records (QX_LAYOUT_BLK) when the
        # parameters qualify (quadrotor, Euler, diagonal weights), the model's generic layout otherwise
        self.block_path = bool(block_path)
        self.exp_layout = None
        self.expansion = self.terminal = None
        self.gains = torch.zeros(self.lib.qx_gains_bytes(mid, B, N) // 4, dtype=f32, device=dev)
        # scratch of the staged line search: queue, flags and candidate costs (quadrotor), winner hints (cart-pole)
        ws = self.lib.qx_linesearch_workspace_bytes(mid, B, N)
        self.ls_workspace = torch.empty(ws, dtype=torch.uint8, device=dev) if (ws and two_phase_linesearch) else None
        self.cur = 0
        self.params = model.params()
        self._select_path()
        self.launches = 0          # kernels launched through the C ABI (bench.py reports it)

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def refresh_params(self):
        self.params = self.model.params()
        self._select_path()

    def _select_path(self):
        """Pick the expansion layout for the current parameters and (re)allocate its buffers."""
        mid, B, N = self.model.model_id, self.B, self.N
        blk = self.block_path and bool(self.lib.qx_blk_supported(C.byref(self.params)))
        layout = _lib.QX_LAYOUT_BLK if blk else self.layout
        if layout == self.exp_layout:
            return
        self.exp_layout = layout
        self.expansion = self.terminal = None           # release before allocating the other layout
        # captured graphs (iterate_pair) and the compaction sub-engines bake these buffers' addresses in
        self.__dict__.pop("_pair_graphs", None)
        self.__dict__.pop("_sub_engines", None)
        eb = self.lib.qx_blk_expansion_bytes(B, N) if blk else self.lib.qx_expansion_bytes(mid, B, N)
        tb = self.lib.qx_blk_terminal_bytes(B) if blk else self.lib.qx_terminal_bytes(mid, B)
        self.expansion = torch.zeros(eb // 4, dtype=torch.float32, device=self.device)
        self.terminal = torch.zeros(tb // 4, dtype=torch.float32, device=self.device)

    @_on_device
    def set_problem(self, x0, u_init, rollout=True):
        """x0 (B, n), u_init (B, N, m): host arrays or tensors (any device); resets the solver state."""
        B, N, n, m = self.B, self.N, self.n, self.m
        x0 = torch.as_tensor(x0, dtype=torch.float64)
        u = torch.as_tensor(u_init, dtype=torch.float64)
        if tuple(x0.shape) != (B, n) or tuple(u.shape) != (B, N, m):
            raise ValueError(f"expected x0 {(B, n)} and u_init {(B, N, m)}, got {tuple(x0.shape)}, {tuple(u.shape)}")
        self.x0.copy_(x0.to(self.device, non_blocking=True).t())
        self.u[self.cur].copy_(u.to(self.device, non_blocking=True).permute(1, 2, 0))
        self.status.zero_()
        self.iters.zero_()
        self.alpha.fill_(-1)
        self.refresh_params()
        if rollout:
            self.rollout()

    # ------------------------------------------------------------------ kernels
    @_on_device
    def rollout(self):
        c = self.cur
        _lib.check(self.lib.qx_rollout_cost(C.byref(self.params), _ptr(self.x0), _ptr(self.u[c]), _ptr(self.x[c]),
                                            _ptr(self.cost), None, self.B, self.N, self.stream), "qx_rollout_cost")
        self.launches += 1

    @_on_device
    def linearize(self, t_start=0):
        c = self.cur
        fn = self.lib.qx_linearize_blk if self.exp_layout == _lib.QX_LAYOUT_BLK else self.lib.qx_linearize
        _lib.check(fn(C.byref(self.params), _ptr(self.x[c]), _ptr(self.u[c]), _ptr(self.expansion),
                      _ptr(self.terminal), _ptr(self.status), self.B, self.N, t_start, self.stream), "qx_linearize")
        self.launches += 1

    @_on_device
    def riccati(self, t_start=0, reg=1e-6):
        if self.exp_layout == _lib.QX_LAYOUT_BLK:
            _lib.check(self.lib.qx_riccati_blk(C.byref(self.params), _ptr(self.expansion), _ptr(self.terminal),
                                               _ptr(self.gains), _ptr(self.status), self.B, self.N, t_start,
                                               C.c_float(reg), self.stream), "qx_riccati_blk")
        else:
            _lib.check(self.lib.qx_riccati(self.model.model_id, _ptr(self.expansion), _ptr(self.terminal),
                                           _ptr(self.gains), _ptr(self.status), self.B, self.N, t_start, C.c_float(reg),
                                           self.stream), "qx_riccati")
        self.launches += 1

    @_on_device
    def linesearch(self, tol, force=False):
        c = self.cur
        self.n_active.zero_()
        _lib.check(self.lib.qx_linesearch(C.byref(self.params), _ptr(self.x[c]), _ptr(self.u[c]), _ptr(self.gains),
                                          _ptr(self.cost), _ptr(self.status), _ptr(self.iters), _ptr(self.alpha),
                                          _ptr(self.x[1 - c]), _ptr(self.u[1 - c]), _ptr(self.n_active),
                                          _ptr(self.ls_workspace), float(tol),
                                          _lib.QX_LS_FORCE_ACTIVE if force else 0, self.B, self.N, self.stream),
                   "qx_linesearch")
        # the staged search of the quadrotor (diagonal weights, workspace given) is three kernels: A, B, W,
        # plus the mirror copy of the finished trajectories when trajectories may finish (not forced)
        # the cart-pole with a workspace is two: the six alphas side by side, then the winners that are not alpha = 1
        staged = self.ls_workspace is not None and self.model.model_id == _lib.QX_MODEL_QUADROTOR and self.params.diag_cost
        par = self.ls_workspace is not None and self.model.model_id == _lib.QX_MODEL_CARTPOLE
        self.launches += (3 if force else 4) if staged else (2 if par else 1)
        self.cur = 1 - c

    @_on_device
    def forward_pass(self, alpha):
        """Candidate rollout with step size `alpha` from the current nominal and gains; returns
        (x (B,N+1,n), u (B,N,m), cost (B,)) tensors.  The candidate goes to scratch tensors of its own: the solver
        state - including the copy of a finished trajectory's final iterate in the other ping-pong buffer - is
        not touched."""
        c = self.cur
        scratch = self.__dict__.get("_fp_scratch")
        if scratch is None:
            scratch = self._fp_scratch = (torch.empty_like(self.x[0]), torch.empty_like(self.u[0]))
        xs, us = scratch
        cost = torch.empty_like(self.cost)
        _lib.check(self.lib.qx_forward_pass(C.byref(self.params), _ptr(self.x[c]), _ptr(self.u[c]), _ptr(self.gains),
                                            float(alpha), _ptr(xs), _ptr(us), _ptr(cost),
                                            self.B, self.N, self.stream), "qx_forward_pass")
        self.launches += 1
        return (xs.permute(2, 0, 1).clone(memory_format=torch.contiguous_format),
                us.permute(2, 0, 1).clone(memory_format=torch.contiguous_format), cost)

    @_on_device
    def backward(self):
        """Gains for the whole horizon: Riccati, or Riccati tail + transformer head."""
        if self.tf is None:
            self.linearize(0)
            self.riccati(0)
        else:
            start = self.N - self.tf.prompt_len
            self.linearize(start)
            self.riccati(start)
            self.tf.forward_device(self)

    def iterate(self, tol=1e-3, force=False):
        self.backward()
        self.linesearch(tol, force)

    @_on_device
    def iterate_pair(self, tol=1e-3, force=False):
        """Two iterations replayed as ONE CUDA graph: the launches of a pair - linearise, Riccati, [the transformer
        forward,] the line-search stages, twice - go out in one call, which removes the launch gaps (~2 % of the classic
        iteration at 16 384 trajectories, tools/graph_probe.py; 1-3 % in transformer mode, tools/tf_graph_probe.py).  A pair, because the iterate ping-pongs between two buffer sets and a
        graph bakes the pointers in: after two iterations the roles are back where the capture found them.  Graphs
        are cached per (buffer parity, tol, force, model parameters); the first call with a new key runs plainly
        (and warms the kernels up), the second captures.  Results are those of two ``iterate`` calls, bit for bit."""
        graphs = self.__dict__.setdefault("_pair_graphs", {})
        # transformer mode: the forward's launches bake in the weight, shift and workspace buffers and the path switches
        tf_key = None if self.tf is None else self.tf.graph_key(self)
        key = (self.cur, float(tol), bool(force), bytes(self.params), self.exp_layout, tf_key)
        state = graphs.get(key)
        if state is None:
            graphs[key] = False
            self.iterate(tol, force)
            self.iterate(tol, force)
            return
        if state is False:
            g = torch.cuda.CUDAGraph()
            launches0, cur0 = self.launches, self.cur
            with torch.cuda.graph(g):
                self.iterate(tol, force)
                self.iterate(tol, force)
            state = graphs[key] = (g, self.launches - launches0)
            self.launches = launches0
            assert self.cur == cur0
        g, n = state
        g.replay()
        self.launches += n

    def solve(self, max_iter=100, tol=1e-3, fixed_iters=None, check_every=1, compact=False, min_compact=1024):
        """Run to convergence (reference stopping rule, per trajectory) or for exactly
        ``fixed_iters`` forced iterations.  Returns the number of batch iterations launched.

        ``compact``: trajectories finish at very different iteration counts (5 ... 30 for the quadrotor), and a
        kernel over a batch in which a random 20 % is still active costs as much as over the full batch (every warp
        still holds an active trajectory).  So once at most half of the batch is active, the active trajectories are
        gathered into a smaller engine (power-of-two sizes, cached), solved there - recursively - and scattered
        back.  Trajectories are independent and the kernels are invariant to the batch composition
        (tests: slices solved alone are bit-identical), so the results do not change.  Off by default: measured on
        a B200 it pays only around 10^5 trajectories (131 072: 110 -> 100 ms per solve); at 16 384 an iteration is at
        its latency floor anyway (50 dependent steps per kernel) and the gathers cost more than they save."""
        if fixed_iters is not None:
            for _ in range(fixed_iters):
                self.iterate(tol, force=True)
            return fixed_iters
        done = 0
        for it in range(max_iter):
            self.iterate(tol)
            done = it + 1
            if (it + 1) % check_every == 0:
                n_act = int(self.n_active.item())
                if n_act == 0:
                    break
                if compact and self.tf is None and 2 * n_act <= self.B and self.B > min_compact and done < max_iter:
                    done += self._solve_compacted(n_act, max_iter - done, tol, check_every, min_compact)
                    break
        return done

    def _solve_compacted(self, n_act, max_iter, tol, check_every, min_compact):
        idx = torch.nonzero(self.status == 0).squeeze(1)
        size = max(min_compact, 1 << (n_act - 1).bit_length())
        cache = self.__dict__.setdefault("_sub_engines", {})
        sub = cache.get(size)
        if sub is None:
            sub = cache[size] = BatchedILQR(self.model, size, self.N, device=self.device,
                                            two_phase_linesearch=self.ls_workspace is not None, block_path=self.block_path)
        sub.refresh_params()
        c = self.cur
        sub.cur = 0
        sub.x[0][:, :, :n_act] = self.x[c].index_select(2, idx)
        sub.u[0][:, :, :n_act] = self.u[c].index_select(2, idx)
        sub.status.fill_(1)                      # padding columns are never touched
        sub.status[:n_act] = 0
        for name in ("cost", "iters", "alpha"):
            getattr(sub, name)[:n_act] = getattr(self, name).index_select(0, idx)
        before = sub.launches
        done = sub.solve(max_iter=max_iter, tol=tol, check_every=check_every, compact=True, min_compact=min_compact)
        self.launches += sub.launches - before
        xs, us = sub.x[sub.cur][:, :, :n_act], sub.u[sub.cur][:, :, :n_act]
        for buf in (0, 1):                       # finished trajectories hold their iterate in both ping-pong buffers
            self.x[buf].index_copy_(2, idx, xs)
            self.u[buf].index_copy_(2, idx, us)
        for name in ("cost", "iters", "alpha", "status"):
            getattr(self, name).index_copy_(0, idx, getattr(sub, name)[:n_act])
        self.n_active.copy_(sub.n_active)
        return done

    # ------------------------------------------------------------------ receding horizon, device resident
    @_on_device
    def advance_receding_horizon(self, x_next=None):
        """One MPC shift without leaving the device: what ``control_step`` does on the host after ``optimize``
        (cartpole_mpc.py:326-332, quadrotor_mpc.py:115-122) - x0 <- the next plant state, u <- u[1:] + [u[-1]] -
        followed by the nominal rollout of the shifted controls.  ``x_next`` (n, B) device tensor is the measured
        state; by default the model's own prediction x_1 of the current solution is used (plant == model).
        Returns the control that was applied, u_0 of the solution, as an (m, B) tensor."""
        c = self.cur
        u_applied = self.u[c][0].clone()
        self.x0.copy_(self.x[c][1] if x_next is None else x_next)
        self.u[c][:-1] = self.u[c][1:].clone()          # the last row keeps u[-1]
        self.status.zero_()
        self.iters.zero_()
        self.alpha.fill_(-1)
        self.rollout()
        return u_applied

    def mpc_sweep(self, x0, u_init, steps, max_iter=100, tol=1e-3, fixed_iters=None, check_every=4):
        """Closed-loop sweep of B independent MPC instances (BASELINE config 5): ``steps`` control steps, each a
        solve from the warm start of the previous one, the plant being the model itself.  Only the per-iteration
        active counter crosses PCIe.  Returns (u_applied (steps, m, B), x_visited (steps + 1, n, B)) on the device."""
        self.set_problem(x0, u_init)
        us = torch.empty((steps, self.m, self.B), dtype=torch.float64, device=self.device)
        xs = torch.empty((steps + 1, self.n, self.B), dtype=torch.float64, device=self.device)
        xs[0] = self.x0
        for s in range(steps):
            self.solve(max_iter=max_iter, tol=tol, fixed_iters=fixed_iters, check_every=check_every)
            us[s] = self.advance_receding_horizon()
            xs[s + 1] = self.x0
        return us, xs

    # ------------------------------------------------------------------ views in the user layout
    # (always fresh copies: for B == 1 ``.contiguous()`` on the permuted view would alias the live ping-pong buffer,
    # which the next line search overwrites)
    def u_batch(self):
        return self.u[self.cur].permute(2, 0, 1).clone(memory_format=torch.contiguous_format)          # (B, N, m)

    def x_batch(self):
        return self.x[self.cur].permute(2, 0, 1).clone(memory_format=torch.contiguous_format)          # (B, N+1, n)

    def gains_batch(self):
        """K (B, N, m, n), k (B, N, m) as float32 tensors."""
        B, N, n, m = self.B, self.N, self.n, self.m
        G = m + m * n
        if self.layout == _lib.QX_LAYOUT_REC:        # float4 [N][G/4][B]
            g = self.gains.view(N, G // 4, B, 4).permute(2, 0, 1, 3).reshape(B, N, G)
        else:
            g = self.gains.view(N, G, B).permute(2, 0, 1)
        return (g[..., m:].reshape(B, N, m, n).clone(memory_format=torch.contiguous_format),
                g[..., :m].clone(memory_format=torch.contiguous_format))

    def load_gains(self, K, k):
        """Upload gains given as K (B, N, m, n), k (B, N, m) into the device record layout."""
        B, N, n, m = self.B, self.N, self.n, self.m
        g = torch.cat([torch.as_tensor(k, dtype=torch.float32).reshape(B, N, m),
                       torch.as_tensor(K, dtype=torch.float32).reshape(B, N, m * n)], dim=-1).to(self.device)
        if self.layout == _lib.QX_LAYOUT_REC:
            self.gains.copy_(g.view(B, N, (m + m * n) // 4, 4).permute(1, 2, 0, 3).reshape(-1))
        else:
            self.gains.copy_(g.permute(1, 2, 0).reshape(-1))

    def expansion_batch(self):
        """dict of (B, N, ...) float32 tensors decoded from the record layout (tests / debugging)."""
        B, N, n, m, R = self.B, self.N, self.n, self.m, self.rec
        if self.exp_layout == _lib.QX_LAYOUT_BLK:
            return self._expansion_batch_blk()
        if self.layout == _lib.QX_LAYOUT_REC:
            e = self.expansion.view(N, B, R).permute(1, 0, 2)
            t = self.terminal.view(B, n + n * n)
        else:
            e = self.expansion.view(N, R, B).permute(2, 0, 1)
            t = self.terminal.view(n + n * n, B).t()
        F = e[..., : n * (n + m)].reshape(B, N, n, n + m)      # [A | B] interleaved row by row
        out = {"A": F[..., :n], "B": F[..., n:]}
        o = n * (n + m)
        for name, shape in (("lx", (n,)), ("lu", (m,)), ("lxx", (n, n)), ("luu", (m, m)), ("lux", (m, n))):
            sz = int(np.prod(shape))
            out[name] = e[..., o:o + sz].reshape((B, N) + shape)
            o += sz
        out["VxN"] = t[:, :n]
        out["VxxN"] = t[:, n:].reshape(B, n, n)
        return out

    def _expansion_batch_blk(self):
        """Decode the compact block records (include/quattro_b200.h, QX_LAYOUT_BLK) into the dense expansion."""
        B, N, n, m = self.B, self.N, self.n, self.m
        p, dev, f32 = self.params, self.device, torch.float32
        noct = (B + 7) // 8
        r = self.expansion.view(N, noct, 12, 8, 4).permute(1, 3, 0, 2, 4).reshape(noct * 8, N, 12, 4)[:B]   # (B,N,slot,4)
        dt = float(p.dt)
        one, zero = torch.ones((B, N), dtype=f32, device=dev), torch.zeros((B, N), dtype=f32, device=dev)

        def mat(rows):
            return torch.stack([torch.stack(row, dim=-1) for row in rows], dim=-2)
        # stored: the entries of the four blocks that vary; the others are 0, 1 or dt by construction (qx_blk.cuh)
        G = mat([[r[..., 0, 0], r[..., 0, 1], r[..., 0, 2]], [r[..., 1, 0], r[..., 1, 1], r[..., 1, 2]],
                 [r[..., 2, 0], r[..., 2, 1], zero]])
        Hp = mat([[r[..., 2, 3], r[..., 3, 0], zero], [r[..., 3, 1], one, zero], [r[..., 3, 2], r[..., 3, 3], one]])
        E = mat([[dt * one, r[..., 4, 0], r[..., 4, 1]], [zero, r[..., 4, 2], r[..., 4, 3]], [zero, r[..., 5, 0], r[..., 5, 1]]])
        Sp = mat([[one, r[..., 5, 2], r[..., 5, 3]], [r[..., 6, 0], one, r[..., 6, 1]], [r[..., 6, 2], r[..., 6, 3], one]])
        d = torch.stack([r[..., 0, 3], r[..., 1, 3], r[..., 2, 2]], dim=-1)
        A = torch.eye(n, dtype=f32, device=dev).repeat(B, N, 1, 1)
        A[..., 0:3, 3:6] = dt * torch.eye(3, dtype=f32, device=dev)
        A[..., 3:6, 6:9] = G
        A[..., 6:9, 6:9] = Hp
        A[..., 6:9, 9:12] = E
        A[..., 9:12, 9:12] = Sp
        Bm = torch.zeros((B, N, n, m), dtype=f32, device=dev)
        Bm[..., 3:6, :] = d[..., None]
        tr = [dt * p.phys[4] / p.phys[1], dt * p.phys[4] / p.phys[2], dt * p.phys[6] / p.phys[3]]
        sg = torch.tensor([[-1, 1, 1, -1], [1, 1, -1, -1], [1, -1, 1, -1]], dtype=f32, device=dev)
        Bm[..., 9:12, :] = sg * torch.tensor(tr, dtype=f32, device=dev)[:, None]
        lu, luu_d = r[..., 7, :], r[..., 8, :]
        Q = torch.tensor(list(p.Q), dtype=torch.float64).view(12, 12).diagonal().to(dev, f32)
        Qf = torch.tensor(list(p.Qf), dtype=torch.float64).view(12, 12).diagonal().to(dev, f32)
        return {"A": A, "B": Bm, "lx": r[..., 9:12, :].reshape(B, N, 12), "lu": lu,
                "lxx": torch.diag(2 * Q).repeat(B, N, 1, 1), "luu": torch.diag_embed(luu_d),
                "lux": torch.zeros((B, N, m, n), dtype=f32, device=dev),
                "VxN": self.terminal.view(B, 16)[:, :12], "VxxN": torch.diag(2 * Qf).repeat(B, 1, 1)}
