"""Closed-loop MPC sweep on one GPU (BASELINE config 5): B independent MPC instances, S control steps each.

Every control step is what ``QuadrotorMPC.control_step`` / ``CartPoleMPC.control_step`` do in the reference
(quadrotor_mpc.py:115-122, cartpole_mpc.py:326-332): a solve from the warm start of the previous step (reference
stopping rule, per trajectory), apply u_0, shift the controls, advance the plant (= the model).  Here the whole loop
stays on the device (``BatchedILQR.mpc_sweep``); only the active-trajectory counter crosses PCIe.

    python tools/mpc_sweep.py [--steps 20] > profiles/rN_mpc_sweep.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
from quattro_transformer_ilqr_b200.sampling import sample_x0

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--batches", type=int, nargs="*", default=[1024, 16384, 65536])
args = ap.parse_args()
print(f"closed-loop MPC sweep, {args.steps} control steps per instance, solves to convergence (reference tolerances), {torch.cuda.get_device_name(0)}")
for system, N, tol in (("quadrotor", 50, 1e-3), ("cartpole", 30, 1e-1)):
    for B in args.batches:
        m = 4 if system == "quadrotor" else 1
        e = BatchedILQR(DeviceModel(system), B, N)
        x0 = sample_x0(system, B, seed=5)
        u0 = np.zeros((B, N, m))
        e.mpc_sweep(x0, u0, 2, tol=tol)                       # warm-up (allocations, first launches)
        torch.cuda.synchronize()
        it0 = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e.set_problem(x0, u0)
        total_iters = torch.zeros((), dtype=torch.int64, device=e.device)
        ev0.record()
        for s in range(args.steps):
            e.solve(max_iter=100, tol=tol, check_every=4)
            total_iters += e.iters.sum()
            e.advance_receding_horizon()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        ti = int(total_iters.item())
        print(f"{system:9s} B={B:7d}: {ms / args.steps:8.3f} ms per control step of the whole batch, "
              f"{B * args.steps / ms * 1e3 / 1e6:7.3f} M control steps/s, {ti / ms * 1e3 / 1e6:7.2f} M trajectory-iterations/s, "
              f"{ti / (B * args.steps):5.2f} iterations per solve")
        del e
