"""Receding-horizon MPC front-ends (the callers of the hot path) and the plant models.

Host-side mirrors of the reference's example classes, so the examples run against this package
on a machine without the reference checkout:

  CartPoleDynamics   examples/cartpole/cartpole_dynamics.py:14-108
  CartPoleMPC        examples/cartpole/cartpole_mpc.py:139-216, 244-269, 303-332 (iLQR modes only;
                     the LQR / switcher blending of :10-116, :287-301 is host glue outside the hot path)
  QuadrotorDynamics  examples/quadrotor/quadrotor_dynamics.py:17-198
  QuadrotorMPC       examples/quadrotor/quadrotor_mpc.py:12-124

The NumPy dynamics/cost methods exist for users who step a plant or evaluate a cost on the host;
the solver never calls them (it resolves the owner object to a CUDA device model).
``batch=B`` makes the controller drive B independent plants at once (x_current of shape (B, n)).
"""
from __future__ import annotations

import numpy as np

from .quattro_ilqr_tf import iLQR_TF


def _integrate(fc, x, u, dt, method):
    if method == "euler":
        return x + dt * fc(x, u)
    if method == "rk4":
        k1 = fc(x, u)
        k2 = fc(x + 0.5 * dt * k1, u)
        k3 = fc(x + 0.5 * dt * k2, u)
        k4 = fc(x + dt * k3, u)
        return x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    raise ValueError(f"Unknown integration method: {method}")


class CartPoleDynamics:
    def __init__(self, m_cart=1.0, m_pole=0.1, length=0.15, gravity=9.81):
        self.m_cart, self.m_pole, self.length, self.gravity = m_cart, m_pole, length, gravity

    def continuous_dynamics(self, x, u):
        _, Xd, th, thd = x
        M, m, l, g = self.m_cart, self.m_pole, self.length, self.gravity
        s, c = np.sin(th), np.cos(th)
        tm = M + m
        temp = (u[0] + m * l * thd ** 2 * s) / tm
        thdd = (-g * s + c * temp) / (l * (4.0 / 3.0 - m * c ** 2 / tm))
        return np.array([Xd, temp - m * l * thdd * c / tm, thd, thdd])

    def discrete_dynamics(self, x, u, dt, method="euler"):
        return _integrate(self.continuous_dynamics, x, u, dt, method)


class QuadrotorDynamics:
    def __init__(self, mass=1.0, Ix=0.02, Iy=0.02, Iz=0.04, arm=0.1, gravity=9.81):
        self.m, self.Ix, self.Iy, self.Iz, self.arm, self.g = mass, Ix, Iy, Iz, arm, gravity
        self.k_yaw = 0.01

    def continuous_dynamics(self, x, u):
        phi, th, psi, p, q, r = x[6:12]
        T = np.sum(u)
        cphi, sphi, cth, sth, cpsi, spsi = np.cos(phi), np.sin(phi), np.cos(th), np.sin(th), np.cos(psi), np.sin(psi)
        acc = (T / self.m) * np.array([spsi * sphi + cpsi * sth * cphi, cpsi * sphi - spsi * sth * cphi, cth * cphi])
        acc[2] -= self.g
        ang = np.array([p + (q * sphi + r * cphi) * np.tan(th), q * cphi - r * sphi, (q * sphi + r * cphi) / cth])
        tau = np.array([self.arm * ((u[1] + u[2]) - (u[0] + u[3])), self.arm * ((u[0] + u[1]) - (u[2] + u[3])),
                        self.k_yaw * (u[0] - u[1] + u[2] - u[3])])
        rate = np.array([(self.Iy - self.Iz) / self.Ix * q * r + tau[0] / self.Ix,
                         (self.Iz - self.Ix) / self.Iy * p * r + tau[1] / self.Iy,
                         (self.Ix - self.Iy) / self.Iz * p * q + tau[2] / self.Iz])
        return np.concatenate([x[3:6], acc, ang, rate])

    def discrete_dynamics(self, x, u, dt, method="euler"):
        return _integrate(self.continuous_dynamics, x, u, dt, method)


class _MPCBase:
    def discrete_dynamics(self, x, u):
        return self.dynamics.discrete_dynamics(x, u, self.dt, method=self.integration_method)

    def final_cost(self, x):
        dx = x - self.x_ref
        return float(dx @ self.Qf @ dx)

    def _shift_warm_start(self, u_seq):
        """u <- u[1:] + [u[-1]] (cartpole_mpc.py:331, quadrotor_mpc.py:121-122)."""
        if isinstance(u_seq, list):
            self.ilqr.u = u_seq[1:] + [u_seq[-1]]
        else:                                            # batched (B, N, m)
            self.ilqr.u = np.concatenate([u_seq[:, 1:], u_seq[:, -1:]], axis=1)


class CartPoleMPC(_MPCBase):
    def __init__(self, horizon=30, dt=0.01, integration_method="rk4", transformer_model=None,
                 log_filename="ilqr_log.pkl", ilqr_only=False, ilqr_tf_only=False, batch=None, device=None):
        self.horizon, self.dt, self.integration_method, self.log_filename = horizon, dt, integration_method, log_filename
        self.ilqr_only, self.ilqr_tf_only = ilqr_only, ilqr_tf_only
        self.dynamics = CartPoleDynamics()
        self.x_ref = np.zeros(4)
        self.Q = np.diag([5.0, 0.1, 10.0, 0.1])
        self.R = np.diag([0.001])
        self.Qf = np.diag([50.0, 6.0, 100.0, 0.1])
        if batch is None:
            self.u_init = [np.array([0.0]) for _ in range(horizon)]
            x0 = self.x_ref
        else:
            self.u_init = np.zeros((batch, horizon, 1))
            x0 = np.tile(self.x_ref, (batch, 1))
        self.ilqr = iLQR_TF(dynamics=self.discrete_dynamics, cost=self.running_cost, cost_final=self.final_cost,
                            x0=x0, u_init=self.u_init, horizon=horizon,
                            tf=transformer_model if ilqr_tf_only else None, tol=1e-1, device=device)

    def running_cost(self, x, u):
        dx = x - self.x_ref
        return float(dx @ self.Q @ dx + u @ self.R @ u)

    def control_step(self, x_current):
        self.ilqr.x0 = x_current
        u_seq, x_seq = self.ilqr.optimize(x_ref=self.x_ref)
        self._shift_warm_start(u_seq)
        return x_seq, (u_seq[0] if isinstance(u_seq, list) else u_seq[:, 0])


class QuadrotorMPC(_MPCBase):
    def __init__(self, horizon=30, dt=0.01, integration_method="rk4", transformer_model=None,
                 log_filename="quad_ilqr_log.pkl", batch=None, device=None):
        self.horizon, self.dt, self.integration_method, self.log_filename = horizon, dt, integration_method, log_filename
        self.dynamics = QuadrotorDynamics()
        self.x_ref = np.zeros(12)
        self.x_ref[2] = 0.5
        self.Q = np.diag([10.0, 10.0, 50.0, 1.0, 1.0, 1.0, 10.0, 10.0, 50.0, 1.0, 1.0, 1.0])
        self.R = np.diag([0.01, 0.01, 0.01, 0.01])
        self.Qf = np.diag([100.0, 100.0, 500.0, 10.0, 10.0, 10.0, 100.0, 100.0, 500.0, 10.0, 10.0, 10.0])
        self.alpha, self.beta = 1000.0, 10.0
        self.transformer_model = transformer_model
        if batch is None:
            self.u_init = [np.zeros(4) for _ in range(horizon)]
            x0 = self.x_ref
        else:
            self.u_init = np.zeros((batch, horizon, 4))
            x0 = np.tile(self.x_ref, (batch, 1))
        self.ilqr = iLQR_TF(dynamics=self.discrete_dynamics, cost=self.running_cost, cost_final=self.final_cost,
                            x0=x0, u_init=self.u_init, horizon=horizon, tf=transformer_model, device=device)
        offset = self.ilqr.get_state_offset()
        offset[2] = 0.5
        self.ilqr.set_state_offset(offset)

    def softplus(self, x, beta):
        return np.log1p(np.exp(beta * x)) / beta

    def running_cost(self, x, u):
        dx = x - self.x_ref
        return dx @ self.Q @ dx + u @ self.R @ u + self.alpha * np.sum(self.softplus(-u, self.beta) ** 2)

    def control_step(self, x_current):
        self.ilqr.x0 = x_current
        u_seq, x_seq = self.ilqr.optimize(x_ref=self.x_ref, verbose=False)
        self._shift_warm_start(u_seq)
        return x_seq, u_seq
