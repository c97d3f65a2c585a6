import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from quattro_transformer_ilqr_b200.sampling import sample_x0
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
B,N=16384,50
for method in ("euler","rk4"):
    e = BatchedILQR(DeviceModel("quadrotor", integration_method=method), B, N)
    x0 = sample_x0("quadrotor", B, seed=0)
    for rep in range(2):
        e.set_problem(x0, np.zeros((B,N,4)))
        tot=[0,0,0]
        for it in range(10):
            ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record(); e.linearize(0); ev[1].record(); e.riccati(0); ev[2].record(); e.linesearch(1e-3, force=True); ev[3].record()
            torch.cuda.synchronize()
            for i in range(3): tot[i]+=ev[i].elapsed_time(ev[i+1])
    print(method, 'layout', e.exp_layout, 'mean ms: lin %.3f ric %.3f ls %.3f sum %.3f'%(tot[0]/10,tot[1]/10,tot[2]/10,sum(tot)/10))
    del e
