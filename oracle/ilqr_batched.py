"""CPU oracle: vectorised fp64 restatement of the batched iLQR iteration.  TEST INFRASTRUCTURE ONLY.

"oracle-AD" of SURVEY.md section 8c: the reference's iteration (quattro_ilqr_tf.py:424-591) for B
independent trajectories at once, with exact derivatives in place of the reference's central
finite differences (so it isolates engine error from the reference's FD noise; the AD-vs-FD gap
is measured in tests/test_oracle.py against the goldens).  Each piece is also exposed on its own
(rollout, linearise, Riccati, line search) so every CUDA kernel has a same-input checker.

Shapes: x (B, N+1, n), u (B, N, m), K (B, N, m, n), k (B, N, m), costs (B,).
"""
from __future__ import annotations

import numpy as np

ALPHAS = np.array([1.0, 0.5, 0.25, 0.1, 0.05, 0.01])   # quattro_ilqr_tf.py:440
REG = 1e-6                                               # quattro_ilqr_tf.py:304


def rollout(model, x0, u):
    """simulate (:127-132) + compute_total_cost (:138-143, left-to-right sum) for a batch."""
    B, N, _ = u.shape
    x = np.empty((B, N + 1, model.n))
    x[:, 0] = x0
    cost = np.zeros(B)
    for t in range(N):
        cost = cost + model.cost(x[:, t], u[:, t])
        x[:, t + 1] = model.f(x[:, t], u[:, t])
    cost = cost + model.final_cost(x[:, N])
    return x, cost


def linearize(model, x, u):
    """A, B and the quadratic cost expansion for every (trajectory, t) (:182-275, exact)."""
    N = u.shape[1]
    A, Bm = model.jacobians(x[:, :N], u)
    lx, lu, lxx, luu, lux = model.cost_expansion(x[:, :N], u)
    VxN, VxxN = model.final_expansion(x[:, N])
    return dict(A=A, B=Bm, lx=lx, lu=lu, lxx=lxx, luu=luu, lux=lux, VxN=VxN, VxxN=VxxN)


def riccati(exp, t_start=0, reg=REG, dtype=np.float64):
    """Backward recursion (:282-319; segment variant :327-366 when t_start > 0).

    Returns K (B, N - t_start, m, n), k (B, N - t_start, m).  ``dtype=np.float32`` runs the same
    recursion in single precision (used to bound the fp32-vs-fp64 gap of the CUDA kernel).
    """
    A, Bm = exp["A"].astype(dtype), exp["B"].astype(dtype)
    lx, lu = exp["lx"].astype(dtype), exp["lu"].astype(dtype)
    lxx, luu, lux = exp["lxx"].astype(dtype), exp["luu"].astype(dtype), exp["lux"].astype(dtype)
    Vx, Vxx = exp["VxN"].astype(dtype), exp["VxxN"].astype(dtype)
    Bsz, N, n, m = A.shape[0], A.shape[1], A.shape[2], Bm.shape[3]
    K = np.zeros((Bsz, N - t_start, m, n), dtype)
    k = np.zeros((Bsz, N - t_start, m), dtype)
    I = (reg * np.eye(m)).astype(dtype)
    T = lambda M: np.swapaxes(M, -1, -2)
    for t in range(N - 1, t_start - 1, -1):
        At, Bt = A[:, t], Bm[:, t]
        Qx = lx[:, t] + np.einsum("bji,bj->bi", At, Vx)
        Qu = lu[:, t] + np.einsum("bji,bj->bi", Bt, Vx)
        Qxx = lxx[:, t] + T(At) @ Vxx @ At
        Qux = lux[:, t] + T(Bt) @ Vxx @ At
        Quu = luu[:, t] + T(Bt) @ Vxx @ Bt
        inv = np.linalg.inv(Quu + I)
        kt = -np.einsum("bij,bj->bi", inv, Qu)
        Kt = -inv @ Qux
        K[:, t - t_start], k[:, t - t_start] = Kt, kt
        Vx = (Qx + np.einsum("bji,bj->bi", Kt, np.einsum("bij,bj->bi", Quu, kt))
              + np.einsum("bji,bj->bi", Kt, Qu) + np.einsum("bji,bj->bi", Qux, kt))
        Vxx = Qxx + T(Kt) @ Quu @ Kt + T(Kt) @ Qux + T(Qux) @ Kt
        Vxx = 0.5 * (Vxx + T(Vxx))
    return K, k


def forward_pass(model, x, u, K, k, alpha):
    """forward_pass (:377-390) for a batch; alpha is a scalar or (B,)."""
    B, N, _ = u.shape
    a = np.broadcast_to(np.asarray(alpha, float), (B,))[:, None]
    xn = np.empty_like(x)
    un = np.empty_like(u)
    xn[:, 0] = x[:, 0]
    cost = np.zeros(B)
    for t in range(N):
        du = k[:, t] + np.einsum("bij,bj->bi", K[:, t], xn[:, t] - x[:, t])
        un[:, t] = u[:, t] + a * du
        cost = cost + model.cost(xn[:, t], un[:, t])
        xn[:, t + 1] = model.f(xn[:, t], un[:, t])
    cost = cost + model.final_cost(xn[:, N])
    return xn, un, cost


def line_search(model, x, u, K, k, cur_cost):
    """First alpha with cand_cost <= cur_cost, per trajectory (:440-451).

    Returns x_new, u_new, cost_new, alpha_idx (-1 = no alpha accepted; x/u/cost then unchanged
    copies of the inputs), margins (B, 6) = cand_cost - cur_cost for every alpha (diagnostics).
    """
    B = u.shape[0]
    x_new, u_new = x.copy(), u.copy()
    cost_new = cur_cost.copy()
    idx = np.full(B, -1, np.int64)
    margins = np.empty((B, len(ALPHAS)))
    for j, a in enumerate(ALPHAS):
        cx, cu, cc = forward_pass(model, x, u, K, k, a)
        margins[:, j] = cc - cur_cost
        take = (idx < 0) & (cc <= cur_cost)
        x_new[take], u_new[take], cost_new[take] = cx[take], cu[take], cc[take]
        idx[take] = j
    return x_new, u_new, cost_new, idx, margins


def pack_prompt(K_seg, k_seg):
    """Inference-time prompt layout [k_0..k_{m-1} | K row-major] (:496-502, quirk a-Q(1))."""
    B, P, m, n = K_seg.shape
    return np.concatenate([k_seg, K_seg.reshape(B, P, m * n)], axis=-1)


def unpack_prediction(pred, m, n):
    """(B, T, m(1+n)) -> k (B,T,m), K (B,T,m,n): interleaved [k_i | K_i,:] rows (:510-514)."""
    B, Tn, _ = pred.shape
    p = pred.reshape(B, Tn, m, 1 + n)
    return p[..., 0], p[..., 1:]


def gains(model, x, u, tf=None, tf_window=0, x_ref=None, state_offset=None, riccati_dtype=np.float64):
    """Full Riccati, or Riccati tail + transformer-predicted head (:489-543)."""
    exp = linearize(model, x, u)
    if tf is None:
        return riccati(exp, 0, dtype=riccati_dtype)
    N = u.shape[1]
    start = N - tf_window
    K_seg, k_seg = riccati(exp, start, dtype=riccati_dtype)
    off = 0.0 if state_offset is None else state_offset
    pred = tf.predict_batch(x - x_ref + off, pack_prompt(K_seg, k_seg))
    k_head, K_head = unpack_prediction(pred.astype(np.float64), model.m, model.n)
    return (np.concatenate([K_head, K_seg.astype(np.float64)], axis=1),
            np.concatenate([k_head, k_seg.astype(np.float64)], axis=1))


def optimize(model, x0, u_init, max_iter=100, tol=1e-3, tf=None, tf_window=0, x_ref=None,
             state_offset=None, fixed_iters=None, keep_logs=False, riccati_dtype=np.float64):
    """Batched optimize (:424-591) with per-trajectory stopping.

    ``fixed_iters=K`` forces every trajectory through exactly K iterations (benchmark mode).
    Returns dict(u, x, cost, iters (number executed, = total_iter + 1), alpha_idx (last), logs).
    """
    x0 = np.asarray(x0, float)
    u = np.array(u_init, float)
    B = u.shape[0]
    x, cost = rollout(model, x0, u)
    active = np.ones(B, bool)
    iters = np.zeros(B, np.int64)
    last_alpha = np.full(B, -1, np.int64)
    logs = []
    n_it = fixed_iters if fixed_iters is not None else max_iter
    for it in range(n_it):
        if not active.any():
            break
        K, k = gains(model, x, u, tf, tf_window, x_ref, state_offset, riccati_dtype)
        xn, un, cn, idx, margins = line_search(model, x, u, K, k, cost)
        found = idx >= 0
        done = (~found) | (np.abs(cost - cn) < tol)
        if keep_logs:
            logs.append(dict(x=x.copy(), u=u.copy(), cost=cost.copy(), K=K, k=k, alpha_idx=idx.copy(),
                             new_cost=cn.copy(), active=active.copy(), margins=margins))
        upd = active
        x[upd], u[upd], cost[upd] = xn[upd], un[upd], cn[upd]
        last_alpha[upd] = idx[upd]
        iters[upd] += 1
        if fixed_iters is None:
            active = active & ~done
    return dict(u=u, x=x, cost=cost, iters=iters, alpha_idx=last_alpha, logs=logs)
