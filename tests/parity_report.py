#!/usr/bin/env python3
"""Parity report: CUDA path vs the fp64 oracle and vs golden logs of the unmodified
reference.  TEST INFRASTRUCTURE (uses oracle/); run on a GPU box:

    python tests/parity_report.py > profiles/rN_parity_report.txt

Per tensor max-norm relative errors (|a - b|_max / |b|_max), line-search decision agreement with the margin of
every disagreement, iteration-count agreement.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ilqr_batched as ob                                   # noqa: E402
from oracle.models import make_model, sample_x0                         # noqa: E402
from tests.util import HORIZON, TOL, checkpoint_dir, golden_runs, relmax  # noqa: E402


def kernels(system, B, block_path=True):
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    N = HORIZON[system]
    model = make_model(system)
    rng = np.random.default_rng(1)
    x0 = sample_x0(system, B, seed=1)
    u = 0.5 * rng.standard_normal((B, N, model.m)) + (2.45 if system == "quadrotor" else 0.0)
    e = BatchedILQR(DeviceModel(system), B, N, block_path=block_path)
    e.set_problem(x0, u)
    x_ref, c_ref = ob.rollout(model, x0, u)
    exp = ob.linearize(model, x_ref, u)
    K_ref, k_ref = ob.riccati(exp)
    e.linearize(0)
    got = {k: v.cpu().numpy() for k, v in e.expansion_batch().items()}
    e.riccati(0)
    K, k = (t.cpu().numpy() for t in e.gains_batch())
    name = f"{system}{'' if block_path or system == 'cartpole' else ' (dense records)'}"
    print(f"  {name:28s} x {relmax(e.x_batch().cpu().numpy(), x_ref):.1e}  cost {relmax(e.cost.cpu().numpy(), c_ref):.1e}  "
          f"A {relmax(got['A'], exp['A']):.1e}  B {relmax(got['B'], exp['B']):.1e}  l_x {relmax(got['lx'], exp['lx']):.1e}  "
          f"l_u {relmax(got['lu'], exp['lu']):.1e}  l_uu {relmax(got['luu'], exp['luu']):.1e}  "
          f"K {relmax(K, K_ref):.1e}  k {relmax(k, k_ref):.1e}")


def solves(system, B):
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    N = HORIZON[system]
    model = make_model(system)
    x0 = sample_x0(system, B, seed=21)
    u0 = np.zeros((B, N, model.m))
    e = BatchedILQR(DeviceModel(system), B, N)
    e.set_problem(x0, u0)
    alphas = []
    for _ in range(100):
        act = (e.status == 0).cpu().numpy()
        e.iterate(tol=TOL[system])
        alphas.append(np.where(act, e.alpha.cpu().numpy(), -9))
        if int(e.n_active.item()) == 0:
            break
    ref = ob.optimize(model, x0, u0, tol=TOL[system], keep_logs=True)
    ref_alphas = [np.where(l["active"], l["alpha_idx"], -9) for l in ref["logs"]]
    n_dec = sum(int((a != -9).sum()) for a in ref_alphas)
    n_same = sum(int(((a == b) & (b != -9)).sum()) for a, b in zip(alphas, ref_alphas))
    u, x = e.u_batch().cpu().numpy(), e.x_batch().cpu().numpy()
    eu = np.abs(u - ref["u"]).reshape(B, -1).max(1) / np.abs(ref["u"]).reshape(B, -1).max(1)
    ex = np.abs(x - ref["x"]).reshape(B, -1).max(1) / np.abs(ref["x"]).reshape(B, -1).max(1)
    print(f"  {system:10s} B={B}: batch iterations {len(alphas)} (oracle {len(ref_alphas)}); iteration counts equal "
          f"{(e.iters.cpu().numpy() == ref['iters']).mean():.4f}; line-search decisions equal {n_same}/{n_dec}; "
          f"cost {relmax(e.cost.cpu().numpy(), ref['cost']):.1e}; u median {np.median(eu):.1e} max {eu.max():.1e}; "
          f"x median {np.median(ex):.1e} max {ex.max():.1e}")


def golden(system, method):
    from quattro_transformer_ilqr_b200.mpc import CartPoleMPC, QuadrotorMPC
    runs = golden_runs(f"ilqr_{system}_classic" + ("_rk4" if method == "rk4" else ""))
    same_a = same_it = 0
    c_first, c_last, c_mid, eu, ex = [], [], [], [], []
    for r in runs:
        mpc = (CartPoleMPC(horizon=30, dt=0.01, integration_method=method, ilqr_only=True) if system == "cartpole"
               else QuadrotorMPC(horizon=50, dt=0.01, integration_method=method))
        il = mpc.ilqr
        il.x0 = r["x0"]
        u, x = il.optimize(mpc.x_ref)
        got = [(-1 if l["alpha"] is None else ob.ALPHAS.tolist().index(l["alpha"])) for l in il.logs]
        same_a += got == list(r["alpha_idx"])
        same_it += il.total_iter == int(r["total_iter"])
        if got == list(r["alpha_idx"]):
            costs = np.array([l["current_cost"] for l in il.logs])
            c_first.append(abs(costs[0] - r["cost"][0]) / abs(r["cost"][0]))
            c_last.append(abs(costs[-1] - r["cost"][-1]) / abs(r["cost"][-1]))
            c_mid.append(np.max(np.abs(costs - r["cost"]) / np.abs(r["cost"])))
            eu.append(relmax(np.array(u), r["u_final"]))
            ex.append(relmax(x, r["x_final"]))
    print(f"  {system:10s} {method:5s}: {len(runs)} reference runs; alpha sequences identical {same_a}/{len(runs)}; iteration "
          f"counts identical {same_it}/{len(runs)}; first cost {max(c_first):.1e}; final cost {max(c_last):.1e}; any cost "
          f"{max(c_mid):.1e} (reference finite-difference noise); final u {max(eu):.1e}; final x {max(ex):.1e}")


def transformer(system):
    from quattro_transformer_ilqr_b200 import TransformerILQR, TransformerPredictor
    g = np.load(os.path.join(ROOT, "tests", "golden", f"tf_forward_{system}.npz"))
    for mode in ("fp32", "rand"):
        w = TransformerILQR(1, 1).load(checkpoint_dir(system))
        if mode == "rand":
            torch.manual_seed(0)
            w.model = TransformerPredictor(w.state_dim, w.control_dim, w.d_model, w.nhead, w.num_decoder_layers,
                                           w.dim_feedforward, w.dropout, w.max_seq_len, w.target_len, w.prompt_len).eval()
            w.invalidate()
        out = {}
        for tc in (False, True):
            w.tensor_cores = tc
            out[tc] = w.predict(g["x_err"], g["kK"])
        print(f"  {system:10s} {'shipped checkpoint (fp32)' if mode == 'fp32' else 'random-init fp32         '}: exact-fp32 path vs "
              f"reference {relmax(out[False], g[f'pred_{mode}']):.1e}; tensor-core path vs reference "
              f"{relmax(out[True], g[f'pred_{mode}']):.1e}; tensor-core vs exact path {relmax(out[True], out[False]):.1e}")


def main():
    print(f"parity report - {torch.cuda.get_device_name(0)}, torch {torch.__version__}")
    print("\n1. kernels vs the fp64 oracle on the same inputs (rollout fp64; linearise / Riccati fp32)")
    kernels("cartpole", 257)
    kernels("quadrotor", 97)
    kernels("quadrotor", 97, block_path=False)
    print("\n2. solves to convergence vs the fp64 exact-derivative oracle (reference stopping rule)")
    solves("cartpole", 512)
    solves("quadrotor", 256)
    print("\n3. iLQR_TF API (B = 1) vs golden logs of the unmodified reference (finite differences, fp64)")
    for system in ("cartpole", "quadrotor"):
        for method in ("euler", "rk4"):
            golden(system, method)
    print("\n4. transformer predictions vs the reference's PyTorch CPU fp32 forward on the golden inputs")
    transformer("cartpole")
    transformer("quadrotor")


if __name__ == "__main__":
    main()
