#!/usr/bin/env python3
"""Fused feed-forward kernel (tf_ffn_kernel) against the two-launch path: bitwise comparison of the predictions.

    python tools/ffn_probe.py [--system quadrotor] [--traj 1200]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_solver import _wrapper, relmax  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="quadrotor")
    ap.add_argument("--traj", type=int, default=1200)
    args = ap.parse_args()
    golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    g = np.load(f"{golden}/tf_forward_{args.system}.npz")
    w = _wrapper(args.system, "rand")
    rng = np.random.default_rng(3)
    reps = max(1, args.traj // len(g["x_err"]))
    x = np.concatenate([g["x_err"] * (1.0 + 0.05 * rng.standard_normal(g["x_err"].shape)) for _ in range(reps)])
    kK = np.concatenate([g["kK"] * (1.0 + 0.05 * rng.standard_normal(g["kK"].shape)) for _ in range(reps)])
    os.environ["QX_TF_TAIL_FUSE"] = "0"
    os.environ["QX_TF_FFN_FUSE"] = "0"
    two = w.predict(x, kK)
    os.environ["QX_TF_FFN_FUSE"] = "1"
    one = w.predict(x, kK)
    again = w.predict(x, kK)
    print(f"{args.system}: {len(x)} trajectories, fused == two launches bitwise: {np.array_equal(one, two)}"
          f" (max rel diff {relmax(one, two):.2e}), run-to-run identical: {np.array_equal(one, again)}, finite: {np.isfinite(one).all()}")
    os.environ["QX_TF_TAIL_FUSE"] = "2"
    tail = w.predict(x, kK)
    again = w.predict(x, kK)
    print(f"{args.system}: layer tail in one launch == separate launches bitwise: {np.array_equal(tail, two)}"
          f" (max rel diff {relmax(tail, two):.2e}), run-to-run identical: {np.array_equal(tail, again)}, finite: {np.isfinite(tail).all()}")


if __name__ == "__main__":
    main()
