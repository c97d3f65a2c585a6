"""GPU parity tests, kernel by kernel, through the C ABI (engine methods are 1:1 wrappers).

Each CUDA kernel is checked against the NumPy oracle on the same seeded inputs.  Tolerances:
fp64 rollouts/costs 1e-12; fp32 linearisation 2e-6; fp32 Riccati gains 1e-5 (max-norm relative per
tensor, SURVEY.md hard part 2); line-search decisions identical.
"""
import numpy as np
import pytest
import torch

from oracle import ilqr_batched as ob
from oracle.models import make_model, sample_x0
from tests.util import HORIZON, relmax

pytestmark = pytest.mark.gpu


def _engine(system, B, method="euler", block_path=True, **kw):
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    return BatchedILQR(DeviceModel(system, integration_method=method, **kw), B, HORIZON[system], block_path=block_path)


def _problem(system, B, seed, u_scale):
    rng = np.random.default_rng(seed)
    model = make_model(system)
    x0 = sample_x0(system, B, seed=seed)
    u = u_scale * rng.standard_normal((B, HORIZON[system], model.m))
    if system == "quadrotor":
        u += 2.45                                  # around hover thrust, both signs of the barrier
    return x0, u


CASES = [("cartpole", 257, 5.0), ("quadrotor", 97, 1.5)]
# the quadrotor with the Euler map and diagonal weights takes the block path (qx_blk.cu) by default;
# "-dense" forces the generic 416-float records + warp-per-trajectory Riccati on the same inputs
BACKWARD_CASES = CASES + [("quadrotor-dense", 97, 1.5)]


def _split(system):
    return system.split("-")[0], not system.endswith("-dense")


@pytest.mark.parametrize("system,B,us", CASES)
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_rollout_cost(system, B, us, method):
    x0, u = _problem(system, B, 1, us)
    e = _engine(system, B, method)
    e.set_problem(x0, u)
    x_ref, c_ref = ob.rollout(make_model(system, method=method), x0, u)
    assert relmax(e.x_batch().cpu().numpy(), x_ref) < 1e-12
    assert relmax(e.cost.cpu().numpy(), c_ref) < 1e-12


@pytest.mark.parametrize("system,B,us", BACKWARD_CASES)
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_linearize(system, B, us, method):
    from quattro_transformer_ilqr_b200 import _lib
    system, blk = _split(system)
    x0, u = _problem(system, B, 2, us)
    e = _engine(system, B, method, block_path=blk)
    assert (e.exp_layout == _lib.QX_LAYOUT_BLK) == (blk and system == "quadrotor" and method == "euler")
    e.set_problem(x0, u)
    e.linearize(0)
    got = {k: v.cpu().numpy() for k, v in e.expansion_batch().items()}
    model = make_model(system, method=method)
    ref = ob.linearize(model, *ob.rollout(model, x0, u)[:1], u)
    for key in ("A", "B", "lx", "lu", "lxx", "luu", "lux", "VxN", "VxxN"):
        if np.abs(ref[key]).max() == 0:
            assert np.abs(got[key]).max() == 0, key
        else:
            assert relmax(got[key], ref[key]) < (2e-6 if method == "euler" else 5e-6), key


def test_linearize_dense_cost_matrices():
    """Non-diagonal Q / R / Qf take the general path (L_x = (Q + Q^T) dx, L_xx = Q + Q^T)."""
    rng = np.random.default_rng(5)
    Q = np.diag([5.0, 0.1, 10.0, 0.1]) + 0.05 * rng.standard_normal((4, 4))
    Qf = np.diag([50.0, 6.0, 100.0, 0.1]) + 0.5 * rng.standard_normal((4, 4))
    x0, u = _problem("cartpole", 64, 5, 5.0)
    e = _engine("cartpole", 64, Q=Q, Qf=Qf, x_ref=np.array([0.1, 0, -0.05, 0]))
    e.set_problem(x0, u)
    e.linearize(0)
    model = make_model("cartpole", Q=Q, Qf=Qf, x_ref=np.array([0.1, 0, -0.05, 0]))
    x_ref, c_ref = ob.rollout(model, x0, u)
    assert relmax(e.cost.cpu().numpy(), c_ref) < 1e-12
    ref = ob.linearize(model, x_ref, u)
    got = {k: v.cpu().numpy() for k, v in e.expansion_batch().items()}
    for key in ("lx", "lxx", "VxN", "VxxN"):
        assert relmax(got[key], ref[key]) < 2e-6, key


@pytest.mark.parametrize("system,B,us", BACKWARD_CASES)
@pytest.mark.parametrize("t_start", [0, "segment"])
def test_riccati(system, B, us, t_start):
    """Full pass and the segment variant (backward_pass_segment) against the fp64 oracle."""
    system, blk = _split(system)
    x0, u = _problem(system, B, 3, 0.3 * us)
    N = HORIZON[system]
    ts = 0 if t_start == 0 else (N - 5 if system == "cartpole" else N - 1)
    e = _engine(system, B, block_path=blk)
    e.set_problem(x0, u)
    e.gains.fill_(float("nan"))
    e.linearize(ts)
    e.riccati(ts)
    K, k = (t.cpu().numpy() for t in e.gains_batch())
    model = make_model(system)
    x_ref, _ = ob.rollout(model, x0, u)
    K_ref, k_ref = ob.riccati(ob.linearize(model, x_ref, u), ts)
    assert np.isnan(K[:, :ts]).all() and not np.isnan(K[:, ts:]).any()        # only [t_start, N) written
    assert relmax(K[:, ts:], K_ref) < 1e-5 and relmax(k[:, ts:], k_ref) < 1e-5
    # per time step too (early steps have much smaller gains than the tail)
    for t in range(ts, N):
        assert relmax(K[:, t], K_ref[:, t - ts]) < 5e-5, t


@pytest.mark.parametrize("B", [1, 7, 8, 9, 64, 1000])
def test_riccati_block_path_sizes_and_status(B):
    """Block path: ragged octets (B not a multiple of 8), B = 1, and masked (finished) trajectories, whose
    gain records must stay untouched; compared with the fp64 oracle and with the dense CUDA path."""
    x0, u = _problem("quadrotor", B, 30 + B, 0.4)
    model = make_model("quadrotor")
    x_ref, _ = ob.rollout(model, x0, u)
    K_ref, k_ref = ob.riccati(ob.linearize(model, x_ref, u))
    out = {}
    for blk in (True, False):
        e = _engine("quadrotor", B, block_path=blk)
        e.set_problem(x0, u)
        mask = np.zeros(B, dtype=bool)
        mask[1::3] = True
        e.status.copy_(torch.as_tensor(mask.astype(np.int32)))
        e.gains.fill_(float("nan"))
        e.linearize(0)
        e.riccati(0)
        K, k = (t.cpu().numpy() for t in e.gains_batch())
        assert np.isnan(K[mask]).all() and np.isnan(k[mask]).all()
        assert not np.isnan(K[~mask]).any()
        assert relmax(K[~mask], K_ref[~mask]) < 1e-5 and relmax(k[~mask], k_ref[~mask]) < 1e-5
        out[blk] = (K[~mask], k[~mask])
    assert relmax(out[True][0], out[False][0]) < 1e-5 and relmax(out[True][1], out[False][1]) < 1e-5


def test_block_path_follows_live_parameters():
    """Non-default physics, weights, reference, dt and horizon (the reference's callers mutate these between calls):
    linearise, Riccati and one line search of the block path against the oracle built with the same values."""
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    B, N = 41, 37
    kw = dict(dt=0.02, mass=1.3, Ix=0.03, Iy=0.025, Iz=0.05, arm=0.12, gravity=9.7,
              Q=np.diag([7.0, 9, 40, 2, 1, 1.5, 12, 8, 30, 1, 2, 0.5]), R=np.diag([0.02, 0.01, 0.03, 0.015]),
              Qf=np.diag([90.0, 110, 400, 12, 8, 10, 80, 120, 450, 9, 11, 10]),
              x_ref=np.array([0.1, -0.2, 0.7, 0, 0, 0, 0.05, -0.03, 0.2, 0, 0, 0]), alpha=700.0, beta=8.0)
    model = make_model("quadrotor", **kw)
    rng = np.random.default_rng(8)
    x0 = sample_x0("quadrotor", B, seed=8)
    u = 0.4 * rng.standard_normal((B, N, 4)) + 3.2
    e = BatchedILQR(DeviceModel("quadrotor", **kw), B, N)
    from quattro_transformer_ilqr_b200 import _lib
    assert e.exp_layout == _lib.QX_LAYOUT_BLK
    e.set_problem(x0, u)
    x_ref, c_ref = ob.rollout(model, x0, u)
    assert relmax(e.x_batch().cpu().numpy(), x_ref) < 1e-12 and relmax(e.cost.cpu().numpy(), c_ref) < 1e-12
    exp = ob.linearize(model, x_ref, u)
    e.linearize(0)
    got = {k: v.cpu().numpy() for k, v in e.expansion_batch().items()}
    for key in ("A", "B", "lx", "lu", "lxx", "luu", "VxN", "VxxN"):
        assert relmax(got[key], exp[key]) < 2e-6, key
    K_ref, k_ref = ob.riccati(exp)
    e.riccati(0)
    K, k = (t.cpu().numpy() for t in e.gains_batch())
    assert relmax(K, K_ref) < 1e-5 and relmax(k, k_ref) < 1e-5
    # line search on the oracle's gains (rounded to fp32): same decisions and candidates
    e.load_gains(K_ref.astype(np.float32), k_ref.astype(np.float32))
    e.linesearch(tol=1e-3)
    xn, un, cn, idx, _ = ob.line_search(model, x_ref, u, K_ref.astype(np.float32).astype(np.float64),
                                        k_ref.astype(np.float32).astype(np.float64), c_ref)
    assert np.array_equal(e.alpha.cpu().numpy(), idx)
    assert relmax(e.cost.cpu().numpy(), cn) < 1e-12 and relmax(e.u_batch().cpu().numpy(), un) < 1e-11


def test_riccati_block_path_indefinite_quu():
    """An indefinite Q_uu (negative R) must go through the pivoted fallback like the reference's LU inverse
    (SURVEY a-Q(4)) instead of producing NaN."""
    B = 16
    R = np.diag([-900.0, 0.01, 0.01, 0.01])
    x0, u = _problem("quadrotor", B, 77, 0.2)
    model = make_model("quadrotor", R=R)
    x_ref, _ = ob.rollout(model, x0, u)
    N = HORIZON["quadrotor"]
    K_ref, k_ref = ob.riccati(ob.linearize(model, x_ref, u), N - 3)
    e = _engine("quadrotor", B, R=R)
    e.set_problem(x0, u)
    e.linearize(N - 3)
    e.riccati(N - 3)
    K, k = (t.cpu().numpy() for t in e.gains_batch())
    assert relmax(K[:, N - 3:], K_ref) < 1e-4 and relmax(k[:, N - 3:], k_ref) < 1e-4


@pytest.mark.parametrize("system,B,us", CASES + [("quadrotor-sequential", 97, 1.5), ("cartpole-sequential", 257, 3.0)])
def test_linesearch_decisions_and_candidates(system, B, us):
    """Same gains on both sides (the oracle's, rounded to fp32): identical accepted alpha,
    candidates and costs to fp64 round-off; covers alphas other than 1 and the no-accept case."""
    two_phase = not system.endswith("-sequential")       # both launch structures (staged / parallel-alpha vs one kernel)
    system = system.split("-")[0]
    x0, u = _problem(system, B, 4, 0.5 * us)
    model = make_model(system)
    x_ref, c_ref = ob.rollout(model, x0, u)
    K, k = ob.riccati(ob.linearize(model, x_ref, u))
    rng = np.random.default_rng(0)
    # spoil some gains so that smaller alphas (and outright rejection) occur
    scale = rng.choice([1.0, 3.0, 8.0, -1.0], size=B, p=[0.4, 0.25, 0.2, 0.15])
    k = (k * scale[:, None, None]).astype(np.float32)
    K = K.astype(np.float32)
    e = _engine(system, B)
    if not two_phase:
        e.ls_workspace = None
    e.set_problem(x0, u)
    e.load_gains(K, k)
    e.linesearch(tol=1e-3)
    xn, un, cn, idx, margins = ob.line_search(model, x_ref, u, K.astype(np.float64), k.astype(np.float64), c_ref)
    got_idx = e.alpha.cpu().numpy()
    assert len(set(idx.tolist())) >= 3, "test should exercise several alphas"
    assert (idx == -1).any()
    assert np.array_equal(got_idx, idx)
    assert relmax(e.cost.cpu().numpy(), cn) < 1e-12
    assert relmax(e.u_batch().cpu().numpy(), un) < 1e-11
    assert relmax(e.x_batch().cpu().numpy(), xn) < 1e-11
    done_ref = (idx < 0) | (np.abs(c_ref - cn) < 1e-3)
    assert np.array_equal(e.status.cpu().numpy() != 0, done_ref)
    assert int(e.n_active.item()) == int((~done_ref).sum())
    assert (e.iters.cpu().numpy() == 1).all()
    # finished trajectories hold their final iterate in BOTH ping-pong buffers
    fin = torch.as_tensor(done_ref, device=e.device)
    assert torch.equal(e.u[0][:, :, fin], e.u[1][:, :, fin]) and torch.equal(e.x[0][:, :, fin], e.x[1][:, :, fin])


@pytest.mark.parametrize("system,B,us", CASES)
def test_forward_pass_entry_point(system, B, us):
    """qx_forward_pass (iLQR_TF.forward_pass) for every alpha of the reference's list vs the oracle."""
    x0, u = _problem(system, B, 9, 0.5 * us)
    model = make_model(system)
    x_ref, _ = ob.rollout(model, x0, u)
    K, k = ob.riccati(ob.linearize(model, x_ref, u))
    K, k = K.astype(np.float32), k.astype(np.float32)
    e = _engine(system, B)
    e.set_problem(x0, u)
    e.load_gains(K, k)
    for alpha in ob.ALPHAS:
        xn, un, cn = ob.forward_pass(model, x_ref, u, K.astype(np.float64), k.astype(np.float64), alpha)
        gx, gu, gc = e.forward_pass(alpha)
        assert relmax(gc.cpu().numpy(), cn) < 1e-12 and relmax(gu.cpu().numpy(), un) < 1e-11
        assert relmax(gx.cpu().numpy(), xn) < 1e-11


def test_forward_pass_reproduces_the_line_search_bitwise():
    """The accept test is "<=" on costs, so every kernel that evaluates f and L must produce the same bits for the
    same inputs: forward_pass(alpha) with the alpha the line search accepted returns exactly the cost, controls
    and states the search kept (quadrotor: staged 4-lane kernels)."""
    x0, u = _problem("quadrotor", 97, 12, 0.6)
    e = _engine("quadrotor", 97)
    e.set_problem(x0, u)
    e.linearize(0)
    e.riccati(0)
    cands = {a: e.forward_pass(alpha) for a, alpha in enumerate(ob.ALPHAS)}
    e.linesearch(tol=1e-3, force=True)
    idx = e.alpha.cpu().numpy()
    assert len(set(idx.tolist())) >= 2
    xb, ub, cb = e.x_batch(), e.u_batch(), e.cost
    for a, (gx, gu, gc) in cands.items():
        sel = torch.as_tensor(idx == a, device=e.device)
        assert torch.equal(cb[sel], gc[sel]) and torch.equal(ub[sel], gu[sel]) and torch.equal(xb[sel], gx[sel])


def test_inactive_trajectories_are_untouched():
    x0, u = _problem("quadrotor", 70, 6, 0.5)
    e = _engine("quadrotor", 70)
    e.set_problem(x0, u)
    e.status[::3] = 1
    before = (e.u[e.cur].clone(), e.cost.clone())
    e.iterate(tol=1e-3)
    frozen = (e.status != 0).cpu()
    assert (e.iters.cpu()[::3] == 0).all() and (e.iters.cpu()[1::3] == 1).all()
    assert torch.equal(e.cost[::3], before[1][::3])
    assert not torch.equal(e.cost[1::3], before[1][1::3])
    assert frozen[::3].all()


@pytest.mark.parametrize("B", [1, 31, 33, 129])
def test_ragged_batch_sizes(B):
    """Tail warps / partial bulk-copy chunks: every trajectory equals its B=1 solve bit for bit."""
    x0, u = _problem("quadrotor", B, 7, 0.5)
    e = _engine("quadrotor", B)
    e.set_problem(x0, u)
    e.solve(fixed_iters=2)
    ref = ob.optimize(make_model("quadrotor"), x0, u, fixed_iters=2)
    assert np.array_equal(e.alpha.cpu().numpy(), ref["alpha_idx"])
    assert relmax(e.cost.cpu().numpy(), ref["cost"]) < 1e-5
    e1 = _engine("quadrotor", 1)
    e1.set_problem(x0[-1:], u[-1:])
    e1.solve(fixed_iters=2)
    assert torch.equal(e1.u_batch()[0], e.u_batch()[-1])       # no cross-trajectory arithmetic


def test_abi_rejects_bad_arguments():
    import ctypes as C
    from quattro_transformer_ilqr_b200 import _lib
    e = _engine("cartpole", 8)
    lib = _lib.load()
    rc = lib.qx_linesearch(C.byref(e.params), e.x[0].data_ptr(), e.u[0].data_ptr(), e.gains.data_ptr(),
                           e.cost.data_ptr(), e.status.data_ptr(), e.iters.data_ptr(), e.alpha.data_ptr(),
                           e.x[0].data_ptr(), e.u[0].data_ptr(), None, None, 1e-3, 0, 8, 30, None)
    assert rc == -1 and b"must differ" in lib.qx_last_error()
    with pytest.raises(ValueError):
        e.set_problem(np.zeros((8, 5)), np.zeros((8, 30, 1)))
    with pytest.raises(RuntimeError, match="qx_riccati failed"):
        e.riccati(t_start=30)


# ---------------------------------------------------------------------------- tensor-core operators
def _ptr(t):
    import ctypes as C
    return C.c_void_p(t.data_ptr())


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,relu", [(512, 128, 1), (384, 128, 0), (128, 512, 0), (128, 128, 0)])
def test_dense_layer_accuracy_and_run_to_run_determinism(N, K, relu):
    """qx_tf_dense (tcgen05, A operand in tensor memory) against float64 on a ragged row count, and 40 launches on
    identical inputs at the benchmark row count: the kernel is deterministic by construction, so any difference
    between two launches is a synchronisation hole (one was found this way: the raw A block was handed back to the
    TMA producer while the splitter's shared-memory reads were still in flight)."""
    import ctypes as C
    from quattro_transformer_ilqr_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(3)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.check(lib.qx_tf_split(_ptr(W), _ptr(hi), _ptr(lo), C.c_size_t(W.numel()), st), "qx_tf_split")
    assert torch.equal(hi + lo, W) or float((hi + lo - W).abs().max()) < 1e-7 * float(W.abs().max())
    for M, reps in [(1000 * 61 + 7, 2), (8192 * 101, 40)]:
        A = torch.randn(M, K, device=dev)
        first = None
        for i in range(reps):
            out = torch.full((M, N), float("nan"), device=dev)
            _lib.check(lib.qx_tf_dense(_ptr(A), _ptr(hi), _ptr(lo), _ptr(b), _ptr(out), M, N, K, relu, st), "qx_tf_dense")
            torch.cuda.synchronize()
            if first is None:
                first = out
                exact = A[:4096].double() @ W.double().T + b.double()
                tail = A[-300:].double() @ W.double().T + b.double()
                if relu:
                    exact, tail = exact.clamp_min(0), tail.clamp_min(0)
                scale = float(exact.abs().max())
                assert float((out[:4096].double() - exact).abs().max()) < 5e-6 * scale
                assert float((out[-300:].double() - tail).abs().max()) < 5e-6 * scale
            else:
                assert torch.equal(out, first), f"launch {i} differs from launch 0 (M={M})"
        del A, first, out


@pytest.mark.gpu
@pytest.mark.parametrize("L", [101, 61, 112, 37, 5])
def test_attention_operator_all_kernels_and_determinism(L):
    """qx_tf_attention: tcgen05 and mma.sync kernels against a float64 causal softmax attention, sequence lengths
    with full, ragged and single key chunks; 20 launches of the tcgen05 kernel must be bit-identical."""
    import ctypes as C
    from quattro_transformer_ilqr_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(4)
    B, d, H = 700, 128, 4
    qkv = torch.randn(B * L, 3 * d, device=dev)
    x = qkv.double().view(B, L, 3, H, 32)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    sc = (q @ k.transpose(-1, -2)) / 32.0 ** 0.5
    sc = sc.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=dev), 1), float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * L, d)
    first = None
    for impl, reps in [(0, 20), (1, 1), (2, 1)]:
        for i in range(reps):
            out = torch.full((B * L, d), float("nan"), device=dev)
            _lib.check(lib.qx_tf_attention(_ptr(qkv), _ptr(out), B, L, d, H, impl, st), "qx_tf_attention")
            torch.cuda.synchronize()
            assert float((out.double() - ref).abs().max()) < 5e-6 * float(ref.abs().max())
            if impl == 0:
                if first is None:
                    first = out
                else:
                    assert torch.equal(out, first), f"tcgen05 attention: launch {i} differs from launch 0"


@pytest.mark.gpu
@pytest.mark.parametrize("K", [128, 256, 512])
def test_dense_layer_with_layernorm_epilogue(K):
    """qx_tf_dense_ln (sub-layer GEMM + residual + LayerNorm in the epilogue; out_proj -> norm1 and linear2 -> norm2 of
    nn.TransformerEncoderLayer, transformer_model.py:112-113) against torch in float64 and in float32 on a ragged row
    count, out of place and in place on the residual stream; 30 launches on identical inputs at the benchmark row
    count are bit-identical; unsupported K is refused."""
    import ctypes as C
    from quattro_transformer_ilqr_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    torch.manual_seed(4)
    N = 128
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    gamma, beta = 1.0 + 0.1 * torch.randn(N, device=dev), 0.1 * torch.randn(N, device=dev)
    hi, lo = torch.empty_like(W), torch.empty_like(W)
    _lib.check(lib.qx_tf_split(_ptr(W), _ptr(hi), _ptr(lo), W.numel(), st), "qx_tf_split")
    for M, reps in ((128 * 37 + 45, 2), (8192 * 101, 30)):
        A = torch.randn(M, K, device=dev)
        res = 0.3 + 2.0 * torch.randn(M, N, device=dev)          # a mean the normalisation has to remove
        first = None
        for i in range(reps):
            out = torch.full((M, N), float("nan"), device=dev)
            _lib.check(lib.qx_tf_dense_ln(_ptr(A), _ptr(hi), _ptr(lo), _ptr(b), _ptr(res), _ptr(gamma), _ptr(beta), _ptr(out),
                                          M, K, st), "qx_tf_dense_ln")
            torch.cuda.synchronize()
            if first is None:
                first = out
                for sl in (slice(0, 4096), slice(M - 300, M)):
                    pre = res[sl].double() + A[sl].double() @ W.double().T + b.double()
                    exact = torch.nn.functional.layer_norm(pre, (N,), gamma.double(), beta.double(), 1e-5)
                    assert float((out[sl].double() - exact).abs().max()) < 5e-6 * float(exact.abs().max())
                    ref32 = torch.nn.functional.layer_norm(res[sl] + torch.nn.functional.linear(A[sl], W, b), (N,), gamma, beta, 1e-5)
                    assert float((out[sl] - ref32).abs().max()) < 2e-5 * float(exact.abs().max())   # torch's GEMM may run TF32-free fp32
            else:
                assert torch.equal(out, first), f"launch {i} differs from launch 0 (M={M})"
        inplace = res.clone()
        _lib.check(lib.qx_tf_dense_ln(_ptr(A), _ptr(hi), _ptr(lo), _ptr(b), _ptr(inplace), _ptr(gamma), _ptr(beta), _ptr(inplace),
                                      M, K, st), "qx_tf_dense_ln")
        torch.cuda.synchronize()
        assert torch.equal(inplace, first)
        del A, res, first, inplace
    small = torch.zeros(128, 96, device=dev)
    rc = lib.qx_tf_dense_ln(_ptr(small), _ptr(hi), _ptr(lo), _ptr(b), _ptr(small), _ptr(gamma), _ptr(beta), _ptr(torch.zeros(128, 128, device=dev)),
                            128, 96, st)
    assert rc != 0 and b"K" in lib.qx_last_error()



# ---------------------------------------------------------------------------- round 2: thread-per-candidate search
def _spoiled_search(B, seed, forced, ws=True, **model_kw):
    """One line search of the staged quadrotor kernels on the oracle's gains with some of them spoiled, so that every
    outcome occurs: alpha = 1, smaller alphas, no alpha at all."""
    x0, u = _problem("quadrotor", B, seed, 0.75)
    model = make_model("quadrotor", **model_kw)
    x_ref, c_ref = ob.rollout(model, x0, u)
    K, k = ob.riccati(ob.linearize(model, x_ref, u))
    rng = np.random.default_rng(seed)
    scale = rng.choice([1.0, 3.0, 8.0, -1.0], size=B, p=[0.4, 0.25, 0.2, 0.15])
    k = (k * scale[:, None, None]).astype(np.float32)
    K = K.astype(np.float32)
    e = _engine("quadrotor", B, **model_kw)
    if not ws:
        e.ls_workspace = None
    e.set_problem(x0, u)
    e.load_gains(K, k)
    nominal = (e.x[e.cur].clone(), e.u[e.cur].clone(), e.cost.clone())
    e.linesearch(tol=1e-3, force=forced)
    ref = ob.line_search(model, x_ref, u, K.astype(np.float64), k.astype(np.float64), c_ref)
    return e, nominal, ref


@pytest.mark.parametrize("forced", [False, True])
def test_no_accepted_alpha_keeps_the_nominal_bit_for_bit(forced):
    """Stage W gives a trajectory without an accepted alpha its nominal by rolling out alpha = 0 (no copy pass): the
    next iterate must be the nominal's exact bits - x, u and cost - forced mode or not, and in non-forced mode the
    trajectory is finished with that iterate in both buffers."""
    e, (xn, un, cn), (_, _, _, idx, _) = _spoiled_search(203, 31, forced)
    none = torch.as_tensor(idx < 0, device=e.device)
    assert int(none.sum()) >= 10
    assert np.array_equal(e.alpha.cpu().numpy(), idx)
    c = e.cur
    assert torch.equal(e.x[c][:, :, none], xn[:, :, none]) and torch.equal(e.u[c][:, :, none], un[:, :, none])
    assert torch.equal(e.cost[none], cn[none])
    assert torch.equal(e.x[1 - c][:, :, none], xn[:, :, none])           # the input buffer still holds it too
    st = e.status.cpu().numpy()
    assert (st[idx < 0] == (0 if forced else 1)).all()


def test_early_rejection_is_switched_off_for_negative_weights():
    """Partial costs only prove a rejection when every cost term is non-negative.  With a negative weight the kernels
    must run every candidate to the end - and still agree with the oracle on every decision and candidate."""
    Q = np.diag([10.0, 10, 50, 1, 1, 1, 10, 10, 50, -40.0, -40.0, 1])      # rewards body rates: costs are not monotone in t
    e, _, (xn, un, cn, idx, _) = _spoiled_search(160, 41, False, Q=Q)
    assert len(set(idx.tolist())) >= 3
    assert np.array_equal(e.alpha.cpu().numpy(), idx)
    assert relmax(e.cost.cpu().numpy(), cn) < 1e-12
    assert relmax(e.u_batch().cpu().numpy(), un) < 1e-11 and relmax(e.x_batch().cpu().numpy(), xn) < 1e-11
    # and the staged search equals the single-launch sequential search bit for bit there as well
    s, _, _ = _spoiled_search(160, 41, False, ws=False, Q=Q)
    assert torch.equal(e.cost, s.cost) and torch.equal(e.alpha, s.alpha) and torch.equal(e.u[e.cur], s.u[s.cur])


def test_staged_search_equals_sequential_search_bitwise():
    """Stages A / B / W (early rejection, costs-only probes, winner rolled out again) against the single-launch kernel
    that walks the alpha list in order: same bits in costs, alphas, iterates, status and both ping-pong buffers."""
    a, _, _ = _spoiled_search(333, 51, False)
    b, _, _ = _spoiled_search(333, 51, False, ws=False)
    assert torch.equal(a.alpha, b.alpha) and torch.equal(a.cost, b.cost) and torch.equal(a.status, b.status)
    for buf in (0, 1):
        assert torch.equal(a.x[buf], b.x[buf]) and torch.equal(a.u[buf], b.u[buf])


def test_linesearch_stage_events_time_the_three_launches():
    """qx_linesearch_stage_events: four caller-owned events recorded in front of stage A and behind A, B, W of the staged
    quadrotor search (what bench.py times the launches with); results are untouched, a cleared hook records nothing."""
    import ctypes as C
    x0, u = _problem("quadrotor", 4096, 5, 1.5)
    outs = []
    for hooked in (False, True):
        e = _engine("quadrotor", 4096)
        e.set_problem(x0, u)
        e.linearize(0); e.riccati(0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for v in ev:
            v.record()
        torch.cuda.synchronize()
        before = [ev[0].elapsed_time(v) for v in ev]
        if hooked:
            arr = (C.c_void_p * 4)(*[v.cuda_event for v in ev])
            assert e.lib.qx_linesearch_stage_events(arr, 4) == 0
        try:
            e.linesearch(1e-3)
            torch.cuda.synchronize()
        finally:
            assert e.lib.qx_linesearch_stage_events(None, 0) == 0
        after = [ev[0].elapsed_time(v) for v in ev]
        if hooked:
            d = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
            assert all(t > 0 for t in d) and d[0] > 0.01         # stage A rolls 4 096 trajectories out: tens of microseconds
        else:
            assert after == before                               # the events still hold their own records
        outs.append((e.cost.clone(), e.alpha.clone(), e.u[e.cur].clone()))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert _engine("quadrotor", 8).lib.qx_linesearch_stage_events(None, 3) != 0      # count must be 0 or 4


def test_gain_record_layout_round_trip():
    """float4 [N][13][B] gains of the quadrotor: load_gains / gains_batch are inverse, and element (t, ch, b) sits where
    include/quattro_b200.h says."""
    B, N = 37, 50
    e = _engine("quadrotor", B)
    rng = np.random.default_rng(0)
    K = rng.standard_normal((B, N, 4, 12)).astype(np.float32)
    k = rng.standard_normal((B, N, 4)).astype(np.float32)
    e.load_gains(K, k)
    K2, k2 = e.gains_batch()
    assert np.array_equal(K2.cpu().numpy(), K) and np.array_equal(k2.cpu().numpy(), k)
    flat = e.gains.cpu().numpy()
    for (t, ch, b) in [(0, 0, 0), (3, 2, 5), (49, 51, 36), (17, 4, 1), (17, 7, 1)]:
        want = k[b, t, ch] if ch < 4 else K[b, t].reshape(-1)[ch - 4]
        assert flat[((t * 13 + ch // 4) * B + b) * 4 + ch % 4] == want


@pytest.mark.parametrize("N", [1, 2, 3, 5, 64])
@pytest.mark.parametrize("system", ["quadrotor", "cartpole"])
def test_short_and_long_horizons(system, N):
    """Horizon edge cases of the staged kernels (two-deep cp.async staging, one record in flight in the Riccati
    thread): N = 1, 2, 3 leave the prefetch pipelines partly empty; N = 64 is longer than any shipped configuration.
    Every kernel against the oracle on the same inputs (a whole solve is no test here: with two or three steps the
    problem is solved after one iteration and the next accept tests are decided by fp32 round-off of the gains)."""
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    B = 67
    model = make_model(system)
    rng = np.random.default_rng(N)
    x0 = sample_x0(system, B, seed=N)
    u = (0.3 * rng.standard_normal((B, N, model.m)) + (2.45 if system == "quadrotor" else 0.0))
    e = BatchedILQR(DeviceModel(system), B, N)
    e.set_problem(x0, u)
    x_ref, c_ref = ob.rollout(model, x0, u)
    assert relmax(e.cost.cpu().numpy(), c_ref) < 1e-12 and relmax(e.x_batch().cpu().numpy(), x_ref) < 1e-12
    K_ref, k_ref = ob.riccati(ob.linearize(model, x_ref, u))
    e.backward()
    K, k = (t.cpu().numpy() for t in e.gains_batch())
    assert relmax(K, K_ref) < 1e-5 and relmax(k, k_ref) < 1e-5
    scale = rng.choice([1.0, 4.0, -1.0], size=B, p=[0.5, 0.3, 0.2])      # some spoiled gains: smaller alphas, rejections
    Kf, kf = K_ref.astype(np.float32), (k_ref * scale[:, None, None]).astype(np.float32)
    e.load_gains(Kf, kf)
    e.linesearch(tol=1e-3)
    xn, un, cn, idx, _ = ob.line_search(model, x_ref, u, Kf.astype(np.float64), kf.astype(np.float64), c_ref)
    assert np.array_equal(e.alpha.cpu().numpy(), idx)
    assert relmax(e.cost.cpu().numpy(), cn) < 1e-12 and relmax(e.u_batch().cpu().numpy(), un) < 1e-11
    assert relmax(e.x_batch().cpu().numpy(), xn) < 1e-11
