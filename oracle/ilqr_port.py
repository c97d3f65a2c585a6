"""CPU oracle: scalar restatement of the reference iLQR iteration (fp64).  TEST INFRASTRUCTURE ONLY.

This is "oracle-FD" of SURVEY.md section 8c: one trajectory at a time, derivatives by the
same central finite differences (eps = 1e-5) the reference takes, the same operation order,
so that on the same machine it reproduces the reference's numbers including its
finite-difference noise.  It is pinned bit-for-bit (costs, gains, alphas) to outputs of the
unmodified reference captured in ``tests/golden/`` (see ``tests/golden/make_golden.py``).
It is also the timed CPU baseline of ``bench.py`` (kind "port"): like the reference it spends
>90 % of its time calling the Python cost function from the finite-difference loops.

Reference (paths relative to /root/reference): quattro_ilqr_tf/quattro_ilqr_tf.py
  rollout :127-132, total cost :138-143, terminal FD :149-174, Jacobian FD :182-204,
  cost FD :217-275, Riccati :282-319 (segment :327-366), forward pass :377-390,
  optimize classic :426-479, transformer :482-591.
"""
from __future__ import annotations

import numpy as np

ALPHAS = (1.0, 0.5, 0.25, 0.1, 0.05, 0.01)   # quattro_ilqr_tf.py:440
EPS = 1e-5
REG = 1e-6                                   # quattro_ilqr_tf.py:304


def _unit(n, i, h):
    e = np.zeros(n)
    e[i] = h
    return e


def fd_grad_bump(fn, z, eps=EPS):
    """Central difference with in-place +/- eps bumps of a copy (reference :149-157)."""
    g = np.zeros_like(z)
    for i in range(len(z)):
        zp = np.copy(z)
        zm = np.copy(z)
        zp[i] += eps
        zm[i] -= eps
        g[i] = (fn(zp) - fn(zm)) / (2 * eps)
    return g


def fd_hess_bump(fn, z, eps=EPS):
    """4-point stencil on bumped copies (reference :163-174)."""
    n = len(z)
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            pts = []
            for si, sj in ((1, 1), (1, -1), (-1, 1), (-1, -1)):
                w = np.copy(z)
                w[i] += si * eps
                w[j] += sj * eps
                pts.append(fn(w))
            H[i, j] = (pts[0] - pts[1] - pts[2] + pts[3]) / (4 * eps ** 2)
    return H


def fd_jacobians(f, x, u, eps=EPS):
    """A = df/dx, B = df/du column by column (reference :182-204)."""
    n, m = x.shape[0], u.shape[0]
    A = np.zeros((n, n))
    B = np.zeros((n, m))
    for i in range(n):
        d = _unit(n, i, eps)
        A[:, i] = (f(x + d, u) - f(x - d, u)) / (2 * eps)
    for i in range(m):
        d = _unit(m, i, eps)
        B[:, i] = (f(x, u + d) - f(x, u - d)) / (2 * eps)
    return A, B


def fd_cost_expansion(L, x, u, eps=EPS):
    """L_x, L_u, L_xx, L_uu, L_xu by central differences (reference :217-275)."""
    n, m = x.shape[0], u.shape[0]
    Lx = np.zeros(n)
    Lu = np.zeros(m)
    L(x, u)  # the reference evaluates (and discards) the base cost, :222
    for i in range(n):
        d = _unit(n, i, eps)
        Lx[i] = (L(x + d, u) - L(x - d, u)) / (2 * eps)
    for i in range(m):
        d = _unit(m, i, eps)
        Lu[i] = (L(x, u + d) - L(x, u - d)) / (2 * eps)
    Lxx = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            a = _unit(n, i, eps)
            b = _unit(n, j, eps)
            Lxx[i, j] = (L(x + a + b, u) - L(x + a - b, u) - L(x - a + b, u) + L(x - a - b, u)) / (4 * eps ** 2)
    Luu = np.zeros((m, m))
    for i in range(m):
        for j in range(m):
            a = _unit(m, i, eps)
            b = _unit(m, j, eps)
            Luu[i, j] = (L(x, u + a + b) - L(x, u + a - b) - L(x, u - a + b) + L(x, u - a - b)) / (4 * eps ** 2)
    Lxu = np.zeros((n, m))
    for i in range(n):
        for j in range(m):
            a = _unit(n, i, eps)
            b = _unit(m, j, eps)
            Lxu[i, j] = (L(x + a, u + b) - L(x + a, u - b) - L(x - a, u + b) + L(x - a, u - b)) / (4 * eps ** 2)
    return Lx, Lu, Lxx, Luu, Lxu


def riccati_step(A, B, Lx, Lu, Lxx, Luu, Lux, Vx, Vxx):
    """One step of the recursion (reference :297-317): returns k, K, new V_x, V_xx.

    Gains use inv(Q_uu + 1e-6 I); the value update uses the un-regularised Q_uu.
    """
    Qx = Lx + A.T @ Vx
    Qu = Lu + B.T @ Vx
    Qxx = Lxx + A.T @ Vxx @ A
    Qux = Lux + B.T @ Vxx @ A
    Quu = Luu + B.T @ Vxx @ B
    inv = np.linalg.inv(Quu + REG * np.eye(Quu.shape[0]))
    k = -inv @ Qu
    K = -inv @ Qux
    Vx_new = Qx + K.T @ Quu @ k + K.T @ Qu + Qux.T @ k
    Vxx_new = Qxx + K.T @ Quu @ K + K.T @ Qux + Qux.T @ K
    Vxx_new = 0.5 * (Vxx_new + Vxx_new.T)
    return k, K, Vx_new, Vxx_new


class ILQRPort:
    """Single-trajectory solver with the reference's control flow and logging."""

    def __init__(self, model, x0, u_init, horizon, max_iter=100, tol=1e-3, tf=None, tf_window=10,
                 derivatives="fd"):
        self.model = model
        self.f = model.f
        self.L = model.cost_scalar
        self.Lf = model.final_cost_scalar
        self.x0 = np.asarray(x0, float)
        self.u = [np.asarray(v, float) for v in u_init]
        self.horizon = horizon
        self.max_iter = max_iter
        self.tol = tol
        self.tf = tf
        if tf_window >= horizon:
            raise ValueError("tf_window must be less than the horizon.")
        self.tf_window = tf.prompt_len if (tf is not None and hasattr(tf, "prompt_len")) else tf_window
        self.state_offset = np.zeros_like(self.x0)
        self.derivatives = derivatives      # "fd" (reference semantics) or "ad" (exact)
        self.logs = []
        self.total_iter = 0

    # -- pieces ---------------------------------------------------------------
    def simulate(self, u_seq):
        xs = [self.x0]
        for u in u_seq:
            xs.append(self.f(xs[-1], u))
        return np.array(xs)

    def total_cost(self, x_seq, u_seq):
        c = 0.0
        for t in range(self.horizon):
            c += self.L(x_seq[t], u_seq[t])
        c += self.Lf(x_seq[-1])
        return c

    def _expand(self, x, u):
        if self.derivatives == "fd":
            A, B = fd_jacobians(self.f, x, u)
            Lx, Lu, Lxx, Luu, Lxu = fd_cost_expansion(self.L, x, u)
            return A, B, Lx, Lu, Lxx, Luu, Lxu.T
        A, B = self.model.jacobians(x, u)
        return (A, B) + tuple(self.model.cost_expansion(x, u))

    def _terminal(self, xN):
        if self.derivatives == "fd":
            return fd_grad_bump(self.Lf, xN), fd_hess_bump(self.Lf, xN)
        return self.model.final_expansion(xN)

    def backward(self, x_seq, u_seq, start=0):
        """Riccati sweep for t in [start, N); outputs indexed t - start (reference :282-366)."""
        Vx, Vxx = self._terminal(x_seq[-1])
        ks = [None] * (self.horizon - start)
        Ks = [None] * (self.horizon - start)
        for t in reversed(range(start, self.horizon)):
            k, K, Vx, Vxx = riccati_step(*self._expand(x_seq[t], u_seq[t]), Vx, Vxx)
            ks[t - start] = k
            Ks[t - start] = K
        return ks, Ks

    def forward(self, x_seq, u_seq, ks, Ks, alpha):
        """alpha scales the feedback term too (reference :381-383)."""
        xs = [self.x0]
        us = []
        for t in range(self.horizon):
            du = ks[t] + Ks[t] @ (xs[t] - x_seq[t])
            us.append(u_seq[t] + alpha * du)
            xs.append(self.f(xs[t], us[t]))
        xs = np.array(xs)
        return xs, us, self.total_cost(xs, us)

    def gains(self, x_seq, u_seq, x_ref):
        """Full Riccati (classic) or Riccati tail + predicted head (transformer mode, :489-543)."""
        if self.tf is None:
            ks, Ks = self.backward(x_seq, u_seq)
            return ks, Ks, ks, Ks
        start = self.horizon - self.tf_window
        ks_seg, Ks_seg = self.backward(x_seq, u_seq, start)
        k_arr = np.array(ks_seg)
        K_arr = np.array(Ks_seg)
        m, n = K_arr.shape[1], K_arr.shape[2]
        prompt = np.concatenate([k_arr, K_arr.reshape(k_arr.shape[0], m * n)], axis=-1)
        pred = self.tf.predict(x_seq - x_ref + self.state_offset, prompt)
        pred = pred.reshape(pred.shape[0], m, 1 + n)
        full_k = np.concatenate([pred[:, :, 0], k_arr], axis=0)
        full_K = np.concatenate([pred[:, :, 1:], K_arr], axis=0)
        return full_k, full_K, ks_seg, Ks_seg

    # -- the loop ---------------------------------------------------------------
    def optimize(self, x_ref=None):
        u_seq = self.u
        for it in range(self.max_iter):
            x_seq = self.simulate(u_seq)
            cur = self.total_cost(x_seq, u_seq)
            ks, Ks, ks_log, Ks_log = self.gains(x_seq, u_seq, x_ref)
            found, alpha_sel, new_x, new_u, new_c = False, None, None, None, None
            for a in ALPHAS:                      # first accepted alpha, "<=" (reference :440-451)
                cx, cu, cc = self.forward(x_seq, u_seq, ks, Ks, a)
                if cc <= cur:
                    found, alpha_sel, new_x, new_u, new_c = True, a, cx, cu, cc
                    u_seq = cu
                    break
            self.logs.append(dict(iteration=it, x_seq=x_seq, u_seq=u_seq, current_cost=cur, k_seq=ks_log,
                                  K_seq=Ks_log, k_full=ks, K_full=Ks, alpha=alpha_sel, new_x_seq=new_x,
                                  new_u_seq=new_u, new_cost=new_c, found_update=found))
            self.total_iter = it
            if (not found) or (abs(cur - new_c) < self.tol):
                break
        self.u = u_seq
        return u_seq, self.simulate(u_seq)
