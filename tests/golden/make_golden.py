#!/usr/bin/env python3
"""Generate golden vectors by executing the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (not on the GPU box)

Writes small .npz fixtures next to this script; the tests only read the fixtures.  Nothing from
the reference's sources is copied: the reference is imported, run, and its outputs recorded.
The shipped checkpoints (data files: fp16 state_dict + normaliser npz) are copied verbatim into
tests/golden/checkpoints/ because they are the weight wire-format contract.

Recorded per run (classic and transformer mode, both systems): x0, and for every iteration the
nominal trajectory, current cost, gains, accepted alpha, new cost; final u/x; iteration count.
"""
import os
import shutil
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "examples", "cartpole"))
sys.path.insert(0, os.path.join(REF, "examples", "quadrotor"))
sys.path.append(os.path.dirname(os.path.dirname(HERE)))

from quattro_ilqr_tf import TransformerILQR, TransformerPredictor  # noqa: E402  (the reference)
from cartpole_mpc import CartPoleMPC  # noqa: E402
from quadrotor_mpc import QuadrotorMPC  # noqa: E402
from oracle.models import sample_x0  # noqa: E402

CKPT = {
    "cartpole": os.path.join(REF, "examples/cartpole/dec3_dmodel128_nhead4_ff256_drop0.1_epoch200_promptlen5_402.7k"),
    "quadrotor": os.path.join(REF, "examples/quadrotor/dec3_dmodel128_nhead4_ff512_drop0.1_epoch200_promptlen1_616.2k"),
}


def load_wrapper(system, mode):
    """mode: 'fp16' (as shipped), 'fp32' (shipped weights up-cast), 'rand' (seed-0 fp32 init)."""
    w = TransformerILQR(state_dim=1, control_dim=1)
    w.device = torch.device("cpu")
    w.load(CKPT[system])
    if mode == "fp16":
        return w
    w.quant_mode = "none"
    if mode == "fp32":
        w.model = w.model.float()
        return w
    torch.manual_seed(0)
    w.model = TransformerPredictor(w.state_dim, w.control_dim, w.d_model, w.nhead, w.num_decoder_layers,
                                   w.dim_feedforward, w.dropout, w.max_seq_len, w.target_len, w.prompt_len).eval()
    return w


def make_mpc(system, tf=None):
    if system == "cartpole":
        return CartPoleMPC(horizon=30, dt=0.01, integration_method="euler", transformer_model=tf,
                           ilqr_only=tf is None, ilqr_tf_only=tf is not None)
    return QuadrotorMPC(horizon=50, dt=0.01, integration_method="euler", transformer_model=tf)


def run(system, x0, tf=None, method="euler"):
    mpc = make_mpc(system, tf)
    mpc.integration_method = method
    il = mpc.ilqr
    il.x0 = np.array(x0, float)
    u, x = il.optimize(mpc.x_ref)
    L = il.logs
    kk, KK = ("k_seq", "K_seq") if tf is None else ("k_seq_seg", "K_seq_seg")
    nan = float("nan")
    alphas = [1.0, 0.5, 0.25, 0.1, 0.05, 0.01]
    return dict(
        x0=np.array(x0, float),
        x=np.array([l["x_seq"] for l in L]),
        cost=np.array([l["current_cost"] for l in L]),
        k=np.array([np.array(l[kk]) for l in L]),
        K=np.array([np.array(l[KK]) for l in L]),
        alpha_idx=np.array([alphas.index(l["alpha"]) if l["found_update"] else -1 for l in L]),
        new_cost=np.array([l["new_cost"] if l["found_update"] else nan for l in L]),
        u_after=np.array([np.array(l["u_seq"]) for l in L]),      # post-update u (quirk a-Q(10))
        u_final=np.array(u), x_final=np.array(x), total_iter=np.array(il.total_iter),
    )


def save_runs(name, runs):
    flat = {"n_runs": np.array(len(runs))}
    for i, r in enumerate(runs):
        for k, v in r.items():
            flat[f"run{i}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, name), **flat)
    print("wrote", name, [int(r["total_iter"]) + 1 for r in runs])


def main():
    kat = {"cartpole": [[0, 0, 0.1, 0], [-0.5, 0, -0.5, 0]],
           "quadrotor": [[0, 0, 0.5, 0, 0, 0, 0.1, 0, 0, 0, 0, 0]]}
    n_extra = {"cartpole": 6, "quadrotor": 3}
    for system in ("cartpole", "quadrotor"):
        x0s = kat[system] + list(sample_x0(system, n_extra[system], seed=7))
        save_runs(f"ilqr_{system}_classic.npz", [run(system, x0) for x0 in x0s])
        for mode in ("fp16", "fp32", "rand"):
            save_runs(f"ilqr_{system}_tf_{mode}.npz", [run(system, x0, load_wrapper(system, mode)) for x0 in x0s])
    # rk4 classic (both MPC classes default to rk4, SURVEY hard part 9)
    save_runs("ilqr_cartpole_classic_rk4.npz", [run("cartpole", x0, method="rk4") for x0 in kat["cartpole"]])
    save_runs("ilqr_quadrotor_classic_rk4.npz", [run("quadrotor", x0, method="rk4") for x0 in kat["quadrotor"]])

    # transformer forward: predict() on trajectories / prompts taken from the classic runs
    for system in ("cartpole", "quadrotor"):
        g = np.load(os.path.join(HERE, f"ilqr_{system}_classic.npz"))
        out = {}
        xs, prompts = [], []
        for i in range(int(g["n_runs"])):
            for it in range(min(2, g[f"run{i}_x"].shape[0])):
                x = g[f"run{i}_x"][it]
                k, K = g[f"run{i}_k"][it], g[f"run{i}_K"][it]
                mpc = make_mpc(system)
                xs.append(x - mpc.x_ref + (np.eye(12)[2] * 0.5 if system == "quadrotor" else 0.0))
                prompts.append(np.concatenate([k, K.reshape(K.shape[0], -1)], axis=-1))
        out["x_err"] = np.array(xs)
        out["kK"] = np.array(prompts)          # full-horizon gains; predict() takes the last prompt_len rows
        for mode in ("fp16", "fp32", "rand"):
            w = load_wrapper(system, mode)
            out[f"pred_{mode}"] = np.array([w.predict(x, p) for x, p in zip(xs, prompts)])
            if mode == "rand":
                sd = w.model.state_dict()
                out["rand_checksum"] = np.array([float(v.double().sum()) for v in sd.values()])
                out["rand_keys"] = np.array(list(sd.keys()))
        np.savez_compressed(os.path.join(HERE, f"tf_forward_{system}.npz"), **out)
        print("wrote tf_forward", system, out["x_err"].shape, out["pred_fp32"].shape)

    for system, d in CKPT.items():
        dst = os.path.join(HERE, "checkpoints", system)
        os.makedirs(dst, exist_ok=True)
        for f in ("tf_model.pt", "tf_model_normalizer.npz"):
            shutil.copyfile(os.path.join(d, f), os.path.join(dst, f))


if __name__ == "__main__":
    main()
