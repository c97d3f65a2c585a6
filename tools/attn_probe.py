#!/usr/bin/env python3
"""Checks and times the attention kernels (qx_tf_attention) against a float64 torch reference.

    python tools/attn_probe.py [--B 8192] [--L 101]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import _lib  # noqa: E402


def reference(qkv, B, L, d, nhead):
    x = qkv.double().view(B, L, 3, nhead, 32)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)   # B, H, L, 32
    s = q @ k.transpose(-1, -2) / np.sqrt(32.0)
    mask = torch.triu(torch.ones(L, L, dtype=torch.bool, device=qkv.device), 1)
    s = s.masked_fill(mask, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * L, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8192)
    ap.add_argument("--L", type=int, default=101)
    ap.add_argument("--check", type=int, default=64, help="trajectories compared with the float64 reference")
    ap.add_argument("--impls", default="0,1")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(0)
    B, L, d, H = args.B, args.L, 128, 4
    qkv = torch.randn(B * L, 3 * d, device=dev)
    ref = reference(qkv[: args.check * L], args.check, L, d, H)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for impl in [int(i) for i in args.impls.split(",")]:
        out = torch.full((B * L, d), float("nan"), device=dev)

        def run():
            _lib.check(lib.qx_tf_attention(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), B, L, d, H, impl, st),
                       "qx_tf_attention")
        run()
        torch.cuda.synchronize()
        got = out[: args.check * L].double()
        err = (got - ref).abs()
        dbg = int(os.environ.get("QX_TF_ATTN_DBG", "0"))
        if impl == 0 and dbg:
            x = qkv[: args.check * L].double().view(args.check, L, 3, H, 32)
            q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
            sraw = q @ k.transpose(-1, -2)                                   # check, H, L, L
            if dbg == 1:
                want = sraw[..., :32]
            elif dbg == 2:
                sc = sraw / np.sqrt(32.0)
                mask = torch.triu(torch.ones(L, L, dtype=torch.bool, device=dev), 1)
                sc = sc.masked_fill(mask, float("-inf"))
                want = torch.exp(sc - sc.max(-1, keepdim=True).values)[..., :32]
            else:
                sc = sraw / np.sqrt(32.0)
                mask = torch.triu(torch.ones(L, L, dtype=torch.bool, device=dev), 1)
                sc = sc.masked_fill(mask, float("-inf"))
                want = torch.exp(sc - sc.max(-1, keepdim=True).values) @ v
            ref = want.transpose(1, 2).reshape(args.check * L, d)
            got = out[: args.check * L].double()
            err = (got - ref).abs()
        print(f"impl {impl}: max abs err {err.max().item():.3e} (ref max {ref.abs().max().item():.3f}), nan count {int(torch.isnan(out).sum())}")
        if not dbg and (err.max().item() > 1e-4 or torch.isnan(out).any()):
            e = err.view(args.check, L, H, 32)
            print("  err by head       :", [f"{e[:, :, h].max().item():.2e}" for h in range(H)])
            print("  err by row block  :", [f"{e[:, r:r + 16].max().item():.2e}" for r in range(0, L, 16)])
            print("  err by dim block  :", [f"{e[..., c:c + 8].max().item():.2e}" for c in range(0, 32, 8)])
            print("  err by trajectory :", [f"{e[b].max().item():.2e}" for b in range(min(8, args.check))])
            bad = (e > 1e-3)
            print("  bad entries:", int(bad.sum()), "of", bad.numel())
            print("  bad by row (first trajectory with any):")
            bb = int(torch.nonzero(bad.view(args.check, -1).any(1))[0])
            for h in range(H):
                rows = bad[bb, :, h].any(1).nonzero().flatten().tolist()
                cols = bad[bb, :, h].any(0).nonzero().flatten().tolist()
                print(f"    traj {bb} head {h}: bad rows {rows[:12]}{'...' if len(rows) > 12 else ''} ({len(rows)}), bad cols {cols[:12]}{'...' if len(cols) > 12 else ''} ({len(cols)})")
            print("  got[0,:4,:4]", got.view(args.check, L, d)[0, :4, :4].tolist())
            print("  ref[0,:4,:4]", ref.view(args.check, L, d)[0, :4, :4].tolist())
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"impl {impl}: {np.median(ts):.4f} ms for B={B} L={L}", flush=True)


if __name__ == "__main__":
    main()
