"""Dev probe: can a transformer-mode iteration be captured in a CUDA graph, and what does the replay save?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel, TransformerILQR, TransformerPredictor
from quattro_transformer_ilqr_b200.sampling import sample_x0

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for system, B, N in (("cartpole", 4096, 30), ("quadrotor", 8192, 50)):
    tf = TransformerILQR(1, 1).load(os.path.join(ROOT, "tests", "golden", "checkpoints", system))
    torch.manual_seed(0)
    tf.model = TransformerPredictor(tf.state_dim, tf.control_dim, tf.d_model, tf.nhead, tf.num_decoder_layers,
                                    tf.dim_feedforward, tf.dropout, tf.max_seq_len, tf.target_len, tf.prompt_len).eval()
    tf.invalidate()
    m = 4 if system == "quadrotor" else 1
    tol = 1e-3 if system == "quadrotor" else 1e-1
    x0 = sample_x0(system, B, seed=0)
    res = {}
    for mode in ("plain", "graph"):
        e = BatchedILQR(DeviceModel(system), B, N, tf=tf)
        e.set_problem(x0, np.zeros((B, N, m)))
        for _ in range(2):
            e.iterate(tol, force=True)
        torch.cuda.synchronize()
        if mode == "graph":
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e.iterate(tol, force=True)
                e.iterate(tol, force=True)
            step = g.replay
        else:
            def step():
                e.iterate(tol, force=True)
                e.iterate(tol, force=True)
        step(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            step()
        ev1.record(); torch.cuda.synchronize()
        res[mode] = (ev0.elapsed_time(ev1) / 10, e.cost.clone())
        print(system, mode, f"{res[mode][0]:.3f} ms per iteration")
    print(system, "same costs:", torch.equal(res["plain"][1], res["graph"][1]))
