"""Long determinism stress: attention (tcgen05) at the benchmark size and the whole transformer forward (pruned last layer)."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quattro_transformer_ilqr_b200 import _lib, TransformerILQR
lib = _lib.load(); dev = torch.device("cuda:0"); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
torch.manual_seed(1)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for (B, L) in [(8192, 101), (4096, 61)]:
    qkv = torch.randn(B * L, 384, device=dev); bad = 0; ref = None
    for i in range(reps):
        out = torch.full((B * L, 128), float("nan"), device=dev)
        _lib.check(lib.qx_tf_attention(P(qkv), P(out), B, L, 128, 4, 0, st), "attn"); torch.cuda.synchronize()
        if ref is None: ref = out
        elif not torch.equal(out, ref): bad += 1
    print(f"attention B={B} L={L}: {bad}/{reps - 1} launches differ", flush=True)
for system in ("quadrotor", "cartpole"):
    g = np.load(f"tests/golden/tf_forward_{system}.npz")
    w = TransformerILQR(1, 1).load(f"tests/golden/checkpoints/{system}")
    k = int(os.environ.get("RACE_B", "2048")) // len(g["x_err"])
    x = np.concatenate([g["x_err"]] * k); kK = np.concatenate([g["kK"]] * k)
    ref = w.predict(x, kK); bad = 0
    for i in range(reps // 4):
        if not np.array_equal(w.predict(x, kK), ref): bad += 1
    print(f"forward {system} B={len(x)}: {bad}/{reps // 4} predictions differ", flush=True)
