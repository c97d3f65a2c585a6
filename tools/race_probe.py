#!/usr/bin/env python3
"""Repeats the tensor-core kernels on identical inputs and counts runs whose output differs from the first run
(races show up as run-to-run differences; both kernels are deterministic by construction)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
P = lambda t: C.c_void_p(t.data_ptr())
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for (B, L) in [(1200, 101), (1200, 61), (8192, 101)]:
    qkv = torch.randn(B * L, 384, device=dev)
    outs = []
    bad = 0
    for i in range(reps):
        out = torch.full((B * L, 128), float("nan"), device=dev)
        _lib.check(lib.qx_tf_attention(P(qkv), P(out), B, L, 128, 4, 0, st), "attn")
        torch.cuda.synchronize()
        if i == 0:
            ref = out
            o1 = torch.empty_like(out)
            _lib.check(lib.qx_tf_attention(P(qkv), P(o1), B, L, 128, 4, 1, st), "attn1")
            print(f"attention B={B} L={L}: tcgen05 vs mma.sync max abs {float((ref - o1).abs().max()):.2e}")
        elif not torch.equal(out, ref):
            bad += 1
            d = (out - ref).abs()
            rows = torch.nonzero(d.max(1).values > 0).flatten()
            print(f"  run {i}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}, rows {rows[:6].tolist()} .. ({len(rows)} rows), "
                  f"traj {sorted(set((rows // L).tolist()))[:6]}, row-in-traj {sorted(set((rows % L).tolist()))[:10]}")
    print(f"attention B={B} L={L}: {bad}/{reps - 1} runs differ")
for (M, N, K) in [(1200 * 101, 384, 128), (1200 * 101, 128, 512), (1200 * 101, 128, 128), (8192 * 101, 512, 128)]:
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.check(lib.qx_tf_split(P(W), P(hi), P(lo), C.c_size_t(W.numel()), st), "split")
    bad = 0
    for i in range(reps):
        out = torch.full((M, N), float("nan"), device=dev)
        _lib.check(lib.qx_tf_dense(P(A), P(hi), P(lo), P(b), P(out), M, N, K, 0, st), "dense")
        torch.cuda.synchronize()
        if i == 0:
            ref = out
        elif not torch.equal(out, ref):
            bad += 1
            d = (out - ref).abs()
            d = torch.nan_to_num(d, nan=1e30)
            rows = torch.nonzero(d.max(1).values > 0).flatten()
            print(f"  run {i}: {int((d > 0).sum())} elements differ, rows {rows[:8].tolist()} ({len(rows)} rows)")
    print(f"dense M={M} N={N} K={K}: {bad}/{reps - 1} runs differ")
