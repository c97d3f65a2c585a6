import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0"); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0); P = lambda t: C.c_void_p(t.data_ptr())
reps = int(sys.argv[1]); M, N, K = 8192 * 101, int(sys.argv[2]), int(sys.argv[3])
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
hi, lo = torch.empty_like(W), torch.empty_like(W)
_lib.check(lib.qx_tf_split(P(W), P(hi), P(lo), C.c_size_t(W.numel()), st), "split")
exact = None
bad = 0
for i in range(reps):
    out = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.qx_tf_dense(P(A), P(hi), P(lo), P(b), P(out), M, N, K, 0, st), "dense")
    torch.cuda.synchronize()
    if i == 0:
        ref = out
    elif not torch.equal(out, ref):
        bad += 1
        d = torch.nan_to_num((out - ref).abs(), nan=1e30)
        rows = torch.nonzero(d.max(1).values > 0).flatten()
        cols = torch.nonzero(d.max(0).values > 0).flatten()
        r0 = int(rows[0])
        ex = (A[r0].double() @ W.double().T + b.double())
        print(f"run {i}: {int((d > 0).sum())} differ, max {float(d.max()):.3e}; rows {rows.tolist()[:20]} cols {int(cols[0])}..{int(cols[-1])} ({len(cols)})")
        groups = 148 // (N // 128)
        c0, c1 = int(cols[0]), int(cols[-1]) + 1
        for k in (-2, -1, 1, 2):
            rr = r0 + k * 128 * groups
            if 0 <= rr < M:
                print(f"      vs correct row of tile {k:+d} in the same CTA: max diff {float((out[r0, c0:c1] - ref[rr, c0:c1]).abs().max()):.3e}")
        # does the wrong row equal A_row' W^T for another row' of the same tile (A rows swapped)?
        t0 = (r0 // 128) * 128
        cand = (ref[t0:t0 + 128, c0:c1] - out[r0, c0:c1]).abs().max(1).values
        print(f"      closest correct row within its own tile: row {int(cand.argmin())} diff {float(cand.min()):.3e}")
        print("   row", r0, "tile row", r0 % 128, "tile", r0 // 128, "err of this run vs fp64", float((out[r0].double() - ex).abs().max()), "err of first run", float((ref[r0].double() - ex).abs().max()))
print(f"dense M={M} N={N} K={K} stream={os.environ.get('QX_TF_GEMM_STREAM')}: {bad}/{reps - 1} runs differ")
