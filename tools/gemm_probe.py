#!/usr/bin/env python3
"""Times one dense layer (qx_tf_dense) per shape / kernel variant with CUDA events.

    python tools/gemm_probe.py [--rows 827392]

Variants are selected through the library's debugging environment (read per call):
QX_TF_GEMM_STREAM=1 -> streaming kernel; QX_TF_GEMM_VARIANT=32|22|21 -> raw/hi and lo ring depths of the
resident-weight kernel; QX_TF_GEMM_DBG bit 0: no global stores, bit 1: no split, bit 2: no epilogue work
(timing experiments only - results are then wrong on purpose).
Variant "ln" (N = 128 shapes only): qx_tf_dense_ln, the same GEMM with residual + LayerNorm in its epilogue, timed next to
qx_tf_dense + the bytes a separate add + LN launch would move (2 x 512 B per row).
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192 * 101)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="stream,0,0:1,0:4")
    ap.add_argument("--shapes", default="384x128,128x128,512x128r,128x512", help="NxK[r] list (r = ReLU)")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(0)
    M = args.rows
    shapes = [(int(t.rstrip("r").split("x")[0]), int(t.rstrip("r").split("x")[1]), int(t.endswith("r"))) for t in args.shapes.split(",")]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for N, K, relu in shapes:
        A = torch.randn(M, K, device=dev)
        W = torch.randn(N, K, device=dev) / np.sqrt(K)
        b = torch.randn(N, device=dev)
        hi, lo = torch.empty_like(W), torch.empty_like(W)
        _lib.check(lib.qx_tf_split(C.c_void_p(W.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                   C.c_size_t(W.numel()), st), "qx_tf_split")
        Cout = torch.empty(M, N, device=dev)
        ref = None
        for var in args.variants.split(","):
            if var == "ln":
                if N != 128 or relu:
                    continue
                res = torch.randn(M, N, device=dev)
                g, be = 1.0 + 0.1 * torch.randn(N, device=dev), 0.1 * torch.randn(N, device=dev)

                def run_ln():
                    _lib.check(lib.qx_tf_dense_ln(C.c_void_p(A.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                                  C.c_void_p(b.data_ptr()), C.c_void_p(res.data_ptr()), C.c_void_p(g.data_ptr()),
                                                  C.c_void_p(be.data_ptr()), C.c_void_p(Cout.data_ptr()), M, K, st), "qx_tf_dense_ln")
                run_ln()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run_ln()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                pre = res[:4096].double() + A[:4096].double() @ W.double().T + b.double()
                exact = torch.nn.functional.layer_norm(pre, (N,), g.double(), be.double(), 1e-5)
                err = ((Cout[:4096].double() - exact).abs().max() / exact.abs().max()).item()
                gb = M * (K + 2 * N) * 4 / 1e9
                print(f"N={N:4d} K={K:4d} + residual + LayerNorm (one launch): {ms:7.4f} ms  {gb / (ms * 1e-3):6.0f} GB/s of A + res + C"
                      f"  relerr vs fp64 {err:.1e}", flush=True)
                del res
                continue
            os.environ.pop("QX_TF_GEMM_STREAM", None)
            os.environ.pop("QX_TF_GEMM_VARIANT", None)
            os.environ.pop("QX_TF_GEMM_DBG", None)
            if var == "stream":
                os.environ["QX_TF_GEMM_STREAM"] = "1"
            else:
                v, _, d = var.partition(":")
                os.environ["QX_TF_GEMM_VARIANT"] = v
                if d:
                    os.environ["QX_TF_GEMM_DBG"] = d

            def run():
                _lib.check(lib.qx_tf_dense(C.c_void_p(A.data_ptr()), C.c_void_p(hi.data_ptr()), C.c_void_p(lo.data_ptr()),
                                           C.c_void_p(b.data_ptr()), C.c_void_p(Cout.data_ptr()), M, N, K, relu, st),
                           "qx_tf_dense")
            run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            note = ""
            if ":" not in var:
                if ref is None:
                    ref = Cout.clone()
                    exact = (A[:4096].double() @ W.double().T + b.double())
                    if relu:
                        exact = exact.clamp_min(0)
                    err = ((Cout[:4096].double() - exact).abs().max() / exact.abs().max()).item()
                    note = f"relerr vs fp64 {err:.1e}"
                else:
                    note = "bitwise equal to first" if torch.equal(ref, Cout) else "DIFFERS from first"
            tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
            print(f"N={N:4d} K={K:4d} relu={relu} {var:>8s}: {ms:7.4f} ms  {tf:6.1f} TFLOP/s fp32-equivalent  {note}", flush=True)
        del A, W, Cout


if __name__ == "__main__":
    main()
