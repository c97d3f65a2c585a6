import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.models import sample_x0
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
B, N = 16384, 50
x0 = sample_x0("quadrotor", B, seed=0)
for two in (True, False):
    e = BatchedILQR(DeviceModel("quadrotor"), B, N, two_phase_linesearch=two)
    e.set_problem(x0, np.zeros((B, N, 4)))
    for it in range(4):
        e.iterate(tol=1e-3, force=True)
        chk = BatchedILQR(DeviceModel("quadrotor"), B, N, two_phase_linesearch=False)
        chk.set_problem(x0, e.u_batch().cpu().numpy())
        dx = (chk.x[chk.cur] - e.x[e.cur]).abs().amax(dim=(0, 1))
        dc = (chk.cost - e.cost).abs()
        a = e.alpha
        bad = (dx > 0) | (dc > 0)
        print(two, it, "mismatch count", int(bad.sum()), "max dx", float(dx.max()), "max dc", float(dc.max()),
              "alphas of bad", torch.bincount(a[bad] + 1, minlength=7).tolist())
    del e, chk
