"""Synthetic initial-state distributions for benchmarks and sweeps.

The sampling ranges are those of the reference's data-collection scripts
(examples/cartpole/training/training_data_collection.py:297-298,
examples/quadrotor/training/training_data_collection.py:272-276); the reference draws an unseeded Latin
hypercube, here the seed is explicit so that runs are reproducible and shards are distinct.
"""
from __future__ import annotations

import numpy as np


def sample_x0(name, B, seed=0):
    """(B, n) float64 initial states: cart-pole X, theta ~ U[-0.5, 0.5]; quadrotor position and attitude from a
    Latin hypercube over [-0.3, 0.3]^2 x [0.49, 0.51] x [-0.2, 0.2]^2 x [-0.5, 0.5], velocities and rates zero."""
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
    raise ValueError(f"unknown system {name!r}")
