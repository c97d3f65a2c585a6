"""CPU oracle: system models (fp64 NumPy).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path
(``quattro_transformer_ilqr_b200``) never does.

Restates the two example systems of the reference (paths relative to
/root/reference):

* cart-pole  : examples/cartpole/cartpole_dynamics.py:48-105 (dynamics),
               examples/cartpole/cartpole_mpc.py:187-189,255-256,268-269 (cost)
* quadrotor  : examples/quadrotor/quadrotor_dynamics.py:64-195 (dynamics),
               examples/quadrotor/quadrotor_mpc.py:40-46,80-100 (cost)

Every function broadcasts over leading batch axes: ``x`` is (..., n) and ``u``
is (..., m).  Two flavours of the scalar cost are kept on purpose:

* ``cost_scalar`` / ``final_cost_scalar`` evaluate one (x, u) with the same
  NumPy expression the reference uses (``dx.T @ Q @ dx``), so the scalar port
  in ``oracle/ilqr_port.py`` is pinned to the reference's floating-point
  behaviour including its finite-difference noise;
* ``cost`` / ``final_cost`` are the batched forms used by the vectorised oracle.

Analytic derivatives (``jacobians``, ``cost_expansion``, ``final_expansion``)
are exact; the reference obtains the same quantities by central finite
differences (quattro_ilqr_tf/quattro_ilqr_tf.py:149-275).
"""
from __future__ import annotations

import numpy as np


def _is_diag(M):
    return np.count_nonzero(M - np.diag(np.diagonal(M))) == 0


class _ModelBase:
    name = "base"
    n = 0
    m = 0

    def __init__(self, dt=0.01, method="euler"):
        if method not in ("euler", "rk4"):
            # same error the reference raises (cartpole_dynamics.py:108)
            raise ValueError(f"Unknown integration method: {method}")
        self.dt = float(dt)
        self.method = method

    # -- discrete map ------------------------------------------------------
    def f(self, x, u):
        """x_{t+1} = f(x_t, u_t); reference ``discrete_dynamics``."""
        dt = self.dt
        if self.method == "euler":
            return x + dt * self.fc(x, u)
        k1 = self.fc(x, u)
        k2 = self.fc(x + 0.5 * dt * k1, u)
        k3 = self.fc(x + 0.5 * dt * k2, u)
        k4 = self.fc(x + dt * k3, u)
        return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)

    def jacobians(self, x, u):
        """Exact A = df/dx (..., n, n) and B = df/du (..., n, m) of the discrete map."""
        dt = self.dt
        n = self.n
        I = np.eye(n)
        if self.method == "euler":
            J, G = self.fc_jac(x, u)
            return I + dt * J, dt * G
        k1 = self.fc(x, u)
        J1, G1 = self.fc_jac(x, u)
        x2 = x + 0.5 * dt * k1
        k2 = self.fc(x2, u)
        J2, G2 = self.fc_jac(x2, u)
        x3 = x + 0.5 * dt * k2
        k3 = self.fc(x3, u)
        J3, G3 = self.fc_jac(x3, u)
        x4 = x + dt * k3
        J4, G4 = self.fc_jac(x4, u)
        D1x, D1u = J1, G1
        D2x = J2 @ (I + 0.5 * dt * D1x)
        D2u = J2 @ (0.5 * dt * D1u) + G2
        D3x = J3 @ (I + 0.5 * dt * D2x)
        D3u = J3 @ (0.5 * dt * D2u) + G3
        D4x = J4 @ (I + dt * D3x)
        D4u = J4 @ (dt * D3u) + G4
        A = I + (dt / 6.0) * (D1x + 2.0 * D2x + 2.0 * D3x + D4x)
        B = (dt / 6.0) * (D1u + 2.0 * D2u + 2.0 * D3u + D4u)
        return A, B

    # -- quadratic part of the cost shared by both systems -------------------
    def _quad(self, M, v):
        if _is_diag(M):
            return np.sum(np.diagonal(M) * v * v, axis=-1)
        return np.einsum("...i,ij,...j->...", v, M, v)

    def final_cost(self, x):
        return self._quad(self.Qf, x - self.x_ref)

    def final_cost_scalar(self, x):
        dx = x - self.x_ref
        return dx.T @ self.Qf @ dx

    def final_expansion(self, x):
        """V_x (..., n), V_xx (..., n, n) at the terminal state (exact)."""
        S = self.Qf + self.Qf.T
        dx = x - self.x_ref
        Vx = dx @ S.T
        Vxx = np.broadcast_to(S, dx.shape[:-1] + S.shape).copy()
        return Vx, Vxx


class CartPoleModel(_ModelBase):
    """Cart-pole, x = [X, Xdot, theta, theta_dot], u = [F]; theta = 0 is upright."""

    name = "cartpole"
    n = 4
    m = 1

    def __init__(self, dt=0.01, method="euler", m_cart=1.0, m_pole=0.1, length=0.15,
                 gravity=9.81, Q=None, R=None, Qf=None, x_ref=None):
        super().__init__(dt, method)
        self.m_cart, self.m_pole, self.length, self.gravity = m_cart, m_pole, length, gravity
        self.Q = np.diag([5.0, 0.1, 10.0, 0.1]) if Q is None else np.asarray(Q, float)
        self.R = np.diag([0.001]) if R is None else np.asarray(R, float)
        self.Qf = np.diag([50.0, 6.0, 100.0, 0.1]) if Qf is None else np.asarray(Qf, float)
        self.x_ref = np.zeros(4) if x_ref is None else np.asarray(x_ref, float)

    def fc(self, x, u):
        Xd, th, thd = x[..., 1], x[..., 2], x[..., 3]
        F = u[..., 0]
        M, m, l, g = self.m_cart, self.m_pole, self.length, self.gravity
        s, c = np.sin(th), np.cos(th)
        tm = M + m
        temp = (F + m * l * (thd ** 2) * s) / tm
        thdd = (-g * s + c * temp) / (l * (4.0 / 3.0 - (m * c ** 2) / tm))
        xdd = temp - (m * l * thdd * c) / tm
        return np.stack([Xd, xdd, thd, thdd], axis=-1)

    def fc_jac(self, x, u):
        th, thd = x[..., 2], x[..., 3]
        F = u[..., 0]
        M, m, l, g = self.m_cart, self.m_pole, self.length, self.gravity
        s, c = np.sin(th), np.cos(th)
        tm = M + m
        ml = m * l
        temp = (F + ml * thd * thd * s) / tm
        t_th = ml * thd * thd * c / tm
        t_thd = 2.0 * ml * thd * s / tm
        t_F = 1.0 / tm
        D = l * (4.0 / 3.0 - m * c * c / tm)
        D_th = l * (2.0 * m * c * s / tm)
        Nn = -g * s + c * temp
        N_th = -g * c - s * temp + c * t_th
        N_thd = c * t_thd
        N_F = c * t_F
        a = Nn / D
        a_th = (N_th * D - Nn * D_th) / (D * D)
        a_thd = N_thd / D
        a_F = N_F / D
        k = ml / tm
        x_th = t_th - k * (a_th * c - a * s)
        x_thd = t_thd - k * c * a_thd
        x_F = t_F - k * c * a_F
        J = np.zeros(x.shape[:-1] + (4, 4))
        G = np.zeros(x.shape[:-1] + (4, 1))
        J[..., 0, 1] = 1.0
        J[..., 1, 2] = x_th
        J[..., 1, 3] = x_thd
        J[..., 2, 3] = 1.0
        J[..., 3, 2] = a_th
        J[..., 3, 3] = a_thd
        G[..., 1, 0] = x_F
        G[..., 3, 0] = a_F
        return J, G

    def cost(self, x, u):
        return self._quad(self.Q, x - self.x_ref) + self._quad(self.R, u)

    def cost_scalar(self, x, u):
        dx = x - self.x_ref
        return float(dx.T @ self.Q @ dx + u.T @ self.R @ u)

    def final_cost_scalar(self, x):
        return float(super().final_cost_scalar(x))

    def cost_expansion(self, x, u):
        """l_x, l_u, l_xx, l_uu, l_ux of the running cost (exact)."""
        Sx = self.Q + self.Q.T
        Su = self.R + self.R.T
        dx = x - self.x_ref
        lead = dx.shape[:-1]
        lx = dx @ Sx.T
        lu = u @ Su.T
        lxx = np.broadcast_to(Sx, lead + Sx.shape).copy()
        luu = np.broadcast_to(Su, lead + Su.shape).copy()
        lux = np.zeros(lead + (self.m, self.n))
        return lx, lu, lxx, luu, lux


class QuadrotorModel(_ModelBase):
    """12-state / 4-rotor quadrotor; x = [p(3), v(3), phi, theta, psi, p, q, r]."""

    name = "quadrotor"
    n = 12
    m = 4
    k_yaw = 0.01  # quadrotor_dynamics.py:143

    def __init__(self, dt=0.01, method="euler", mass=1.0, Ix=0.02, Iy=0.02, Iz=0.04, arm=0.1,
                 gravity=9.81, Q=None, R=None, Qf=None, x_ref=None, alpha=1000.0, beta=10.0):
        super().__init__(dt, method)
        self.mass, self.Ix, self.Iy, self.Iz, self.arm, self.g = mass, Ix, Iy, Iz, arm, gravity
        self.Q = (np.diag([10.0, 10.0, 50.0, 1.0, 1.0, 1.0, 10.0, 10.0, 50.0, 1.0, 1.0, 1.0])
                  if Q is None else np.asarray(Q, float))
        self.R = np.diag([0.01] * 4) if R is None else np.asarray(R, float)
        self.Qf = (np.diag([100.0, 100.0, 500.0, 10.0, 10.0, 10.0, 100.0, 100.0, 500.0, 10.0, 10.0, 10.0])
                   if Qf is None else np.asarray(Qf, float))
        if x_ref is None:
            x_ref = np.zeros(12)
            x_ref[2] = 0.5
        self.x_ref = np.asarray(x_ref, float)
        self.alpha = float(alpha)
        self.beta = float(beta)

    def fc(self, x, u):
        vx, vy, vz = x[..., 3], x[..., 4], x[..., 5]
        phi, th, psi = x[..., 6], x[..., 7], x[..., 8]
        p, q, r = x[..., 9], x[..., 10], x[..., 11]
        u1, u2, u3, u4 = u[..., 0], u[..., 1], u[..., 2], u[..., 3]
        T = u1 + u2 + u3 + u4
        cphi, sphi = np.cos(phi), np.sin(phi)
        cth, sth = np.cos(th), np.sin(th)
        cpsi, spsi = np.cos(psi), np.sin(psi)
        ax = (T / self.mass) * (spsi * sphi + cpsi * sth * cphi)
        ay = (T / self.mass) * (cpsi * sphi - spsi * sth * cphi)
        az = -self.g + (T / self.mass) * (cth * cphi)
        phid = p + q * sphi * np.tan(th) + r * cphi * np.tan(th)
        thd = q * cphi - r * sphi
        psid = (q * sphi + r * cphi) / np.cos(th)
        tau_phi = self.arm * ((u2 + u3) - (u1 + u4))
        tau_th = self.arm * ((u1 + u2) - (u3 + u4))
        tau_psi = self.k_yaw * (u1 - u2 + u3 - u4)
        pd = ((self.Iy - self.Iz) / self.Ix) * (q * r) + (tau_phi / self.Ix)
        qd = ((self.Iz - self.Ix) / self.Iy) * (p * r) + (tau_th / self.Iy)
        rd = ((self.Ix - self.Iy) / self.Iz) * (p * q) + (tau_psi / self.Iz)
        return np.stack([vx, vy, vz, ax, ay, az, phid, thd, psid, pd, qd, rd], axis=-1)

    def fc_jac(self, x, u):
        phi, th, psi = x[..., 6], x[..., 7], x[..., 8]
        p, q, r = x[..., 9], x[..., 10], x[..., 11]
        T = u[..., 0] + u[..., 1] + u[..., 2] + u[..., 3]
        cphi, sphi = np.cos(phi), np.sin(phi)
        cth, sth = np.cos(th), np.sin(th)
        cpsi, spsi = np.cos(psi), np.sin(psi)
        tth = sth / cth
        sec = 1.0 / cth
        Tm = T / self.mass
        im = 1.0 / self.mass
        J = np.zeros(x.shape[:-1] + (12, 12))
        G = np.zeros(x.shape[:-1] + (12, 4))
        J[..., 0, 3] = 1.0
        J[..., 1, 4] = 1.0
        J[..., 2, 5] = 1.0
        dirx = spsi * sphi + cpsi * sth * cphi
        diry = cpsi * sphi - spsi * sth * cphi
        dirz = cth * cphi
        J[..., 3, 6] = Tm * (spsi * cphi - cpsi * sth * sphi)
        J[..., 3, 7] = Tm * (cpsi * cth * cphi)
        J[..., 3, 8] = Tm * diry
        J[..., 4, 6] = Tm * (cpsi * cphi + spsi * sth * sphi)
        J[..., 4, 7] = Tm * (-spsi * cth * cphi)
        J[..., 4, 8] = -Tm * dirx
        J[..., 5, 6] = -Tm * cth * sphi
        J[..., 5, 7] = -Tm * sth * cphi
        for j in range(4):
            G[..., 3, j] = im * dirx
            G[..., 4, j] = im * diry
            G[..., 5, j] = im * dirz
        w = q * sphi + r * cphi
        J[..., 6, 6] = (q * cphi - r * sphi) * tth
        J[..., 6, 7] = w * sec * sec
        J[..., 6, 9] = 1.0
        J[..., 6, 10] = sphi * tth
        J[..., 6, 11] = cphi * tth
        J[..., 7, 6] = -w
        J[..., 7, 10] = cphi
        J[..., 7, 11] = -sphi
        J[..., 8, 6] = (q * cphi - r * sphi) * sec
        J[..., 8, 7] = w * sth * sec * sec
        J[..., 8, 10] = sphi * sec
        J[..., 8, 11] = cphi * sec
        c1 = (self.Iy - self.Iz) / self.Ix
        c2 = (self.Iz - self.Ix) / self.Iy
        c3 = (self.Ix - self.Iy) / self.Iz
        J[..., 9, 10] = c1 * r
        J[..., 9, 11] = c1 * q
        J[..., 10, 9] = c2 * r
        J[..., 10, 11] = c2 * p
        J[..., 11, 9] = c3 * q
        J[..., 11, 10] = c3 * p
        G[..., 9, :] = (self.arm / self.Ix) * np.array([-1.0, 1.0, 1.0, -1.0])
        G[..., 10, :] = (self.arm / self.Iy) * np.array([1.0, 1.0, -1.0, -1.0])
        G[..., 11, :] = (self.k_yaw / self.Iz) * np.array([1.0, -1.0, 1.0, -1.0])
        return J, G

    def softplus(self, z):
        # quadrotor_mpc.py:80  softplus(z; beta) = log1p(exp(beta z)) / beta
        return np.log1p(np.exp(self.beta * z)) / self.beta

    def cost(self, x, u):
        pen = np.sum(self.softplus(-u) ** 2, axis=-1)
        return self._quad(self.Q, x - self.x_ref) + self._quad(self.R, u) + self.alpha * pen

    def cost_scalar(self, x, u):
        dx = x - self.x_ref
        c = dx.T @ self.Q @ dx + u.T @ self.R @ u
        penalty = np.sum(self.softplus(-u) ** 2)
        c += self.alpha * penalty
        return c

    def cost_expansion(self, x, u):
        Sx = self.Q + self.Q.T
        Su = self.R + self.R.T
        dx = x - self.x_ref
        lead = dx.shape[:-1]
        z = -u
        sp = self.softplus(z)
        sg = 1.0 / (1.0 + np.exp(-self.beta * z))        # softplus'(z)
        sp2 = self.beta * sg * (1.0 - sg)                  # softplus''(z)
        lx = dx @ Sx.T
        lu = u @ Su.T - 2.0 * self.alpha * sp * sg
        lxx = np.broadcast_to(Sx, lead + Sx.shape).copy()
        luu = np.broadcast_to(Su, lead + Su.shape).copy()
        d2 = 2.0 * self.alpha * (sg * sg + sp * sp2)
        idx = np.arange(self.m)
        luu[..., idx, idx] += d2
        lux = np.zeros(lead + (self.m, self.n))
        return lx, lu, lxx, luu, lux


def make_model(name, **kw):
    if name == "cartpole":
        return CartPoleModel(**kw)
    if name == "quadrotor":
        return QuadrotorModel(**kw)
    raise ValueError(f"unknown model {name!r}")


# Synthetic initial-state distributions: the sampling ranges of
# examples/cartpole/training/training_data_collection.py:297-298 and
# examples/quadrotor/training/training_data_collection.py:272-276.
def sample_x0(name, B, seed=0):
    if name == "cartpole":
        rng = np.random.default_rng(seed)
        x0 = np.zeros((B, 4))
        x0[:, 0] = rng.uniform(-0.5, 0.5, B)
        x0[:, 2] = rng.uniform(-0.5, 0.5, B)
        return x0
    if name == "quadrotor":
        from scipy.stats import qmc
        lo = np.array([-0.3, -0.3, 0.49, -0.2, -0.2, -0.5])
        hi = np.array([0.3, 0.3, 0.51, 0.2, 0.2, 0.5])
        s = qmc.scale(qmc.LatinHypercube(d=6, seed=seed).random(n=B), lo, hi)
        x0 = np.zeros((B, 12))
        x0[:, 0:3] = s[:, 0:3]
        x0[:, 6:9] = s[:, 3:6]
        return x0
    raise ValueError(name)
