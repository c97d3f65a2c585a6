"""Device-model registry: maps the reference's Python callables onto the built-in CUDA models.

``iLQR_TF.__init__`` takes three Python callables (quattro_ilqr_tf/quattro_ilqr_tf.py:66-86 in the
reference).  A GPU engine cannot call them, so the callables are resolved to one of the two
systems the reference ships (cart-pole, quadrotor) and their parameters are read off the owning
MPC object at EVERY ``optimize()`` call, because the reference's callers mutate them between
calls (``mpc.x_ref`` in the key callbacks of cartpole_sim.py:107-112, ``ilqr.tol`` :78, ...).

Accepted ``dynamics`` arguments
  * a bound method of an MPC-like object (``CartPoleMPC.discrete_dynamics`` /
    ``QuadrotorMPC.discrete_dynamics``: examples/cartpole/cartpole_mpc.py:207-216,
    examples/quadrotor/quadrotor_mpc.py:55-63) whose owner exposes ``dynamics``, ``dt``,
    ``integration_method``, ``Q``, ``R``, ``Qf``, ``x_ref`` (and ``alpha``, ``beta`` for the quadrotor);
  * a :class:`DeviceModel` instance (explicit descriptor).
Anything else raises ``TypeError`` — there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_INTEG = {"euler": _lib.QX_EULER, "rk4": _lib.QX_RK4}


class DeviceModel:
    """Explicit descriptor of a built-in model; attributes may be mutated between solves."""

    def __init__(self, name, dt=0.01, integration_method="euler", Q=None, R=None, Qf=None, x_ref=None,
                 alpha=1000.0, beta=10.0, **phys):
        if name not in ("cartpole", "quadrotor"):
            raise ValueError(f"unknown device model {name!r} (built-in: cartpole, quadrotor)")
        self.name = name
        self.dt = dt
        self.integration_method = integration_method
        if name == "cartpole":
            self.n, self.m, self.model_id = 4, 1, _lib.QX_MODEL_CARTPOLE
            self.phys = dict(m_cart=1.0, m_pole=0.1, length=0.15, gravity=9.81)
            self.Q = np.diag([5.0, 0.1, 10.0, 0.1]) if Q is None else Q
            self.R = np.diag([0.001]) if R is None else R
            self.Qf = np.diag([50.0, 6.0, 100.0, 0.1]) if Qf is None else Qf
            self.x_ref = np.zeros(4) if x_ref is None else x_ref
            self.alpha, self.beta = 0.0, 1.0
        else:
            self.n, self.m, self.model_id = 12, 4, _lib.QX_MODEL_QUADROTOR
            self.phys = dict(mass=1.0, Ix=0.02, Iy=0.02, Iz=0.04, arm=0.1, gravity=9.81, k_yaw=0.01)
            self.Q = np.diag([10.0, 10.0, 50.0, 1.0, 1.0, 1.0, 10.0, 10.0, 50.0, 1.0, 1.0, 1.0]) if Q is None else Q
            self.R = np.diag([0.01] * 4) if R is None else R
            self.Qf = (np.diag([100.0, 100.0, 500.0, 10.0, 10.0, 10.0, 100.0, 100.0, 500.0, 10.0, 10.0, 10.0])
                       if Qf is None else Qf)
            if x_ref is None:
                x_ref = np.zeros(12)
                x_ref[2] = 0.5
            self.x_ref = x_ref
            self.alpha, self.beta = alpha, beta
        unknown = set(phys) - set(self.phys)
        if unknown:
            raise TypeError(f"unknown physical parameters for {name}: {sorted(unknown)}")
        self.phys.update(phys)

    # ------------------------------------------------------------------
    def _live(self):
        """(phys list, dt, method, Q, R, Qf, x_ref, alpha, beta) read now."""
        order = (("m_cart", "m_pole", "length", "gravity") if self.name == "cartpole"
                 else ("mass", "Ix", "Iy", "Iz", "arm", "gravity", "k_yaw"))
        return ([self.phys[k] for k in order], self.dt, self.integration_method, self.Q, self.R, self.Qf,
                self.x_ref, self.alpha, self.beta)

    def params(self):
        """Build the QxParams block from the CURRENT attribute values."""
        phys, dt, method, Q, R, Qf, x_ref, alpha, beta = self._live()
        if method not in _INTEG:
            # same error the reference's discrete_dynamics raises (cartpole_dynamics.py:108)
            raise ValueError(f"Unknown integration method: {method}")
        n, m = self.n, self.m
        Q, R, Qf = (np.asarray(M, dtype=np.float64) for M in (Q, R, Qf))
        x_ref = np.asarray(x_ref, dtype=np.float64).reshape(-1)
        if Q.shape != (n, n) or Qf.shape != (n, n) or R.shape != (m, m) or x_ref.shape != (n,):
            raise ValueError(f"{self.name}: cost matrices / x_ref have the wrong shape")
        p = _lib.QxParams()
        p.model_id, p.integrator, p.n, p.m = self.model_id, _INTEG[method], n, m
        p.diag_cost = int(all(np.count_nonzero(M - np.diag(np.diagonal(M))) == 0 for M in (Q, R, Qf)))
        p.dt = float(dt)
        for i, v in enumerate(phys):
            p.phys[i] = float(v)
        for i in range(n):
            p.x_ref[i] = x_ref[i]
            for j in range(n):
                p.Q[i * n + j] = Q[i, j]
                p.Qf[i * n + j] = Qf[i, j]
        for i in range(m):
            for j in range(m):
                p.R[i * m + j] = R[i, j]
        p.barrier_alpha, p.barrier_beta = float(alpha), float(beta)
        return p


class _OwnerModel(DeviceModel):
    """Model whose parameters live on an MPC object (read at every params() call)."""

    def __init__(self, name, owner):
        super().__init__(name)
        self.owner = owner

    def _live(self):
        o, d = self.owner, self.owner.dynamics
        if self.name == "cartpole":
            phys = [d.m_cart, d.m_pole, d.length, d.gravity]
            alpha, beta = 0.0, 1.0
        else:
            phys = [d.m, d.Ix, d.Iy, d.Iz, d.arm, d.g, getattr(d, "k_yaw", 0.01)]
            alpha, beta = o.alpha, o.beta
        return phys, o.dt, o.integration_method, o.Q, o.R, o.Qf, o.x_ref, alpha, beta


def resolve(dynamics, cost=None, cost_final=None):
    """Map the reference-style callables onto a device model or raise TypeError."""
    if isinstance(dynamics, DeviceModel):
        return dynamics
    dm = getattr(dynamics, "device_model", None)
    if isinstance(dm, DeviceModel):
        return dm
    owner = getattr(dynamics, "__self__", None)
    dyn = getattr(owner, "dynamics", None)
    need = ("dt", "integration_method", "Q", "R", "Qf", "x_ref")
    if owner is not None and dyn is not None and all(hasattr(owner, a) for a in need):
        for fn in (cost, cost_final):
            if fn is not None and getattr(fn, "__self__", owner) is not owner:
                raise TypeError("dynamics, cost and cost_final must be methods of the same MPC object")
        if all(hasattr(dyn, a) for a in ("m_cart", "m_pole", "length", "gravity")):
            return _OwnerModel("cartpole", owner)
        if all(hasattr(dyn, a) for a in ("m", "Ix", "Iy", "Iz", "arm", "g")) and hasattr(owner, "alpha"):
            return _OwnerModel("quadrotor", owner)
    raise TypeError(
        "iLQR_TF (B200 engine) needs a built-in device model: pass bound methods of a CartPoleMPC / "
        "QuadrotorMPC-like object or a DeviceModel descriptor.  Arbitrary Python callables cannot run "
        "on the GPU and there is no CPU fallback.")
