#!/usr/bin/env python3
"""Probe: does replaying two classic iLQR iterations as ONE CUDA graph beat launching their 14 kernels one by one?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel  # noqa: E402
from quattro_transformer_ilqr_b200.sampling import sample_x0  # noqa: E402

B, N = 16384, 50
dev = torch.device("cuda:0")
eng = BatchedILQR(DeviceModel("quadrotor"), B, N, device=dev)
x0 = torch.as_tensor(sample_x0("quadrotor", B, seed=0)).to(dev)
u0 = torch.zeros((B, N, 4), dtype=torch.float64, device=dev)


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ten_plain():
    eng.set_problem(x0, u0)
    for _ in range(10):
        eng.iterate(1e-3, force=True)


ten_plain()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.set_problem(x0, u0)
    eng.iterate(1e-3, force=True); eng.iterate(1e-3, force=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        eng.iterate(1e-3, force=True)
        eng.iterate(1e-3, force=True)


def ten_graph():
    eng.set_problem(x0, u0)
    for _ in range(5):
        g.replay()


ten_graph()
ref_cost = None
eng.set_problem(x0, u0)
for _ in range(10):
    eng.iterate(1e-3, force=True)
c_plain = eng.cost.clone()
ten_graph()
print("same costs:", torch.equal(c_plain, eng.cost))
print(f"10 iterations, plain launches: {timed(ten_plain, 20):.4f} ms; as 5 graph replays: {timed(ten_graph, 20):.4f} ms")
