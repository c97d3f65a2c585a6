"""Dev probe: per-iteration kernel times and accepted-alpha histogram for the benchmark workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quattro_transformer_ilqr_b200.sampling import sample_x0
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel

B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 50
e = BatchedILQR(DeviceModel("quadrotor"), B, N)
x0 = sample_x0("quadrotor", B, seed=0)
for rep in range(2):
    e.set_problem(x0, np.zeros((B, N, 4)))
    for it in range(12):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); e.linearize(0); ev[1].record(); e.riccati(0); ev[2].record(); e.linesearch(1e-3, force=True); ev[3].record()
        torch.cuda.synchronize()
        a = e.alpha.cpu().numpy()
        hist = [int((a == j).sum()) for j in range(-1, 6)]
        if rep == 1:
            print(f"it {it:2d} lin {ev[0].elapsed_time(ev[1]):.3f} ric {ev[1].elapsed_time(ev[2]):.3f} ls {ev[2].elapsed_time(ev[3]):.3f} ms  alpha hist(-1,0..5) {hist}")
