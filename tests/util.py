"""Helpers shared by the tests: golden loading and error norms."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HORIZON = {"cartpole": 30, "quadrotor": 50}
TOL = {"cartpole": 1e-1, "quadrotor": 1e-3}        # cartpole_mpc.py:215 / iLQR_TF default


def golden_runs(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = []
    for i in range(int(g["n_runs"])):
        pre = f"run{i}_"
        out.append({k[len(pre):]: g[k] for k in g.files if k.startswith(pre)})
    return out


def relmax(a, b):
    """max-norm relative error of a tensor (SURVEY.md hard part 2: per tensor, not per element)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def checkpoint_dir(system):
    return os.path.join(GOLDEN, "checkpoints", system)
