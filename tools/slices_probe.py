"""Probe: does running the batch as S slices on S CUDA streams (kernels of different slices overlap on the SMs) beat
one engine over the whole batch?  python tools/slices_probe.py"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
from quattro_transformer_ilqr_b200.sampling import sample_x0
from quattro_transformer_ilqr_b200.sharding import shard_bounds

B, N = 16384, 50
dev = torch.device("cuda:0")
x0 = torch.as_tensor(sample_x0("quadrotor", B, seed=0)).to(dev)
u0 = torch.zeros((B, N, 4), dtype=torch.float64, device=dev)
for S in (1, 2, 3, 4):
    bounds = [shard_bounds(B, c, S) for c in range(S)]
    engs = [BatchedILQR(DeviceModel("quadrotor"), hi - lo, N, device=dev) for lo, hi in bounds]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    def steps(k, reset_every=10):
        for i in range(k):
            for (lo, hi), e, st in zip(bounds, engs, streams):
                with torch.cuda.stream(st):
                    if i % reset_every == 0:
                        e.set_problem(x0[lo:hi], u0[lo:hi])
                    e.iterate(1e-3, force=True)
    steps(20)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    ev0.record()
    for st in streams: st.wait_stream(cur)
    steps(100)
    for st in streams: cur.wait_stream(st)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 100
    print(f"S={S}: {ms:.4f} ms/step  {B / ms / 1e3:.2f} M traj-it/s")
