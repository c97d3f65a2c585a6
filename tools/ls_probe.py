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
    tot = [0.0, 0.0, 0.0]
    for it in range(12):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(); e.linearize(0); ev[1].record(); e.riccati(0); ev[2].record(); e.linesearch(1e-3, force=True); ev[3].record()
        torch.cuda.synchronize()
        a = e.alpha.cpu().numpy()
        hist = [int((a == j).sum()) for j in range(-1, 6)]
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        if it < 10:
            tot = [tot[i] + t[i] for i in range(3)]
        if rep == 1:
            print(f"it {it:2d} lin {t[0]:.3f} ric {t[1]:.3f} ls {t[2]:.3f} ms  alpha hist(-1,0..5) {hist}")
    if rep == 1:
        print(f"mean over iterations 0-9: lin {tot[0] / 10:.3f} ric {tot[1] / 10:.3f} ls {tot[2] / 10:.3f} ms, sum {sum(tot) / 10:.3f} ms")
