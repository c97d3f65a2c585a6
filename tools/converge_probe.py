"""Probe: cost of every batch iteration of a solve to convergence (active set shrinks) - python tools/converge_probe.py"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
from quattro_transformer_ilqr_b200.sampling import sample_x0
B, N = 16384, 50
e = BatchedILQR(DeviceModel("quadrotor"), B, N)
x0 = sample_x0("quadrotor", B, seed=0)
for rep in range(2):
    e.set_problem(x0, np.zeros((B, N, 4)))
    rows = []
    for it in range(100):
        act = int((e.status == 0).sum().item())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); e.iterate(1e-3); ev[1].record()
        torch.cuda.synchronize()
        rows.append((it, act, ev[0].elapsed_time(ev[1])))
        if int(e.n_active.item()) == 0:
            break
    if rep == 1:
        tot = sum(r[2] for r in rows)
        for it, act, ms in rows:
            print(f"it {it:2d} active {act:6d} {ms:.3f} ms")
        print(f"total {tot:.2f} ms, trajectory-iterations {int(e.iters.sum().item())}")
