"""CPU tests: the oracle against the golden vectors recorded from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import ilqr_batched as ob
from oracle.ilqr_port import ILQRPort
from oracle.models import make_model, sample_x0
from oracle.transformer import TransformerOracle, sinusoid_pe
from tests.util import HORIZON, TOL, checkpoint_dir, golden_runs, relmax


def _port(system, run, **kw):
    model = make_model(system)
    N = HORIZON[system]
    p = ILQRPort(model, run["x0"], [np.zeros(model.m) for _ in range(N)], N, tol=TOL[system], **kw)
    u, x = p.optimize(model.x_ref)
    return p, np.array(u), x


@pytest.mark.parametrize("system,idx", [("cartpole", 0), ("cartpole", 1), ("cartpole", 4), ("quadrotor", 1)])
def test_scalar_port_is_bitwise_the_reference(system, idx):
    """oracle-FD: same finite differences, same operation order -> identical floats."""
    run = golden_runs(f"ilqr_{system}_classic")[idx]
    p, u, x = _port(system, run)
    assert p.total_iter == int(run["total_iter"])
    assert np.array_equal(np.array([l["current_cost"] for l in p.logs]), run["cost"])
    assert np.array_equal(np.array([np.array(l["K_seq"]) for l in p.logs]), run["K"])
    assert np.array_equal(np.array([np.array(l["k_seq"]) for l in p.logs]), run["k"])
    assert np.array_equal(u, run["u_final"]) and np.array_equal(x, run["x_final"])


def test_known_answers_from_baseline_md():
    """BASELINE.md section 4 / SURVEY.md section 4 known-answer values."""
    r = golden_runs("ilqr_cartpole_classic")[0]
    assert r["cost"][0] == pytest.approx(2.717429502282358, rel=1e-12)
    assert r["cost"][1] == pytest.approx(1.3473424119126785, rel=1e-12)
    assert r["u_final"][0, 0] == pytest.approx(-6.755067, abs=1e-6)
    r = golden_runs("ilqr_cartpole_classic")[1]
    assert r["cost"][:3] == pytest.approx([116.54505574619608, 80.47242396282621, 80.35179335570987], rel=1e-12)
    r = golden_runs("ilqr_quadrotor_classic")[0]
    assert r["cost"][:4] == pytest.approx([2997.738432876904, 2113.314492958847, 1779.5520273040233, 1413.2685744742648], rel=1e-12)
    assert list(r["alpha_idx"]) == [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    assert r["u_final"][0] == pytest.approx([4.064699, 0.711285, 0.711285, 4.064699], abs=1e-6)


@pytest.mark.parametrize("system", ["cartpole", "quadrotor"])
def test_batched_oracle_matches_reference_runs(system):
    """oracle-AD vs the reference: same alphas / iteration counts, costs 1e-5, u/x within the
    reference's own finite-difference noise (documented tail 1e-4, SURVEY.md hard part 2)."""
    runs = golden_runs(f"ilqr_{system}_classic")
    model = make_model(system)
    N = HORIZON[system]
    x0 = np.array([r["x0"] for r in runs])
    out = ob.optimize(model, x0, np.zeros((len(runs), N, model.m)), tol=TOL[system], keep_logs=True)
    for i, r in enumerate(runs):
        assert out["iters"][i] == int(r["total_iter"]) + 1
        alphas = [int(l["alpha_idx"][i]) for l in out["logs"] if l["active"][i]]
        assert alphas == list(r["alpha_idx"])
        costs = np.array([l["cost"][i] for l in out["logs"] if l["active"][i]])
        # iteration 0 has identical inputs; later iterates inherit the reference's FD noise, which
        # sensitive mid-run iterations amplify (measured up to 7.4e-4 mid-run on a 20-iteration quadrotor case; first and final costs agree to 1e-5)
        assert abs(costs[0] - r["cost"][0]) <= 1e-12 * abs(r["cost"][0])
        assert np.max(np.abs(costs - r["cost"]) / np.abs(r["cost"])) < 2e-3
        assert abs(costs[-1] - r["cost"][-1]) <= 1e-5 * abs(r["cost"][-1])
        assert relmax(out["u"][i], r["u_final"]) < 1e-4
        assert relmax(out["x"][i], r["x_final"]) < 1e-4
        assert relmax(out["logs"][0]["k"][i], r["k"][0]) < 5e-4      # FD noise on l_uu (cart-pole 2R = 0.002)


@pytest.mark.parametrize("system", ["cartpole", "quadrotor"])
def test_batched_oracle_rk4(system):
    runs = golden_runs(f"ilqr_{system}_classic_rk4")
    model = make_model(system, method="rk4")
    N = HORIZON[system]
    x0 = np.array([r["x0"] for r in runs])
    out = ob.optimize(model, x0, np.zeros((len(runs), N, model.m)), tol=TOL[system], keep_logs=True)
    for i, r in enumerate(runs):
        assert out["iters"][i] == int(r["total_iter"]) + 1
        assert relmax(out["u"][i], r["u_final"]) < 1e-4


@pytest.mark.parametrize("system", ["cartpole", "quadrotor"])
def test_analytic_derivatives_match_finite_differences(system):
    """Exact Jacobians / cost expansions vs central differences of the model functions."""
    model = make_model(system)
    rng = np.random.default_rng(3)
    x = sample_x0(system, 5, seed=3) + 0.05 * rng.standard_normal((5, model.n))
    u = rng.standard_normal((5, model.m)) * (1.0 if system == "quadrotor" else 5.0)
    for method in ("euler", "rk4"):
        model.method = method
        A, B = model.jacobians(x, u)
        eps = 1e-6
        for i in range(model.n):
            d = np.zeros(model.n); d[i] = eps
            assert np.allclose(A[:, :, i], (model.f(x + d, u) - model.f(x - d, u)) / (2 * eps), atol=2e-8)
        for j in range(model.m):
            d = np.zeros(model.m); d[j] = eps
            assert np.allclose(B[:, :, j], (model.f(x, u + d) - model.f(x, u - d)) / (2 * eps), atol=2e-8)
    lx, lu, lxx, luu, lux = model.cost_expansion(x, u)
    eps = 1e-5
    for j in range(model.m):
        d = np.zeros(model.m); d[j] = eps
        assert np.allclose(lu[:, j], (model.cost(x, u + d) - model.cost(x, u - d)) / (2 * eps), rtol=1e-6, atol=1e-6)
        g = lambda uu: model.cost_expansion(x, uu)[1]
        assert np.allclose(luu[:, :, j], (g(u + d) - g(u - d)) / (2 * eps), rtol=1e-6, atol=1e-6)
    # the softplus barrier curvature at u = 0 quoted in SURVEY.md 8a-M
    if system == "quadrotor":
        assert model.cost_expansion(x[:1], np.zeros((1, 4)))[3][0, 0, 0] == pytest.approx(846.59, abs=0.05)


def _torch_predictor(system, mode):
    from quattro_transformer_ilqr_b200 import TransformerILQR, TransformerPredictor
    w = TransformerILQR(1, 1).load(checkpoint_dir(system))
    if mode == "rand":
        torch.manual_seed(0)
        w.model = TransformerPredictor(w.state_dim, w.control_dim, w.d_model, w.nhead, w.num_decoder_layers,
                                       w.dim_feedforward, w.dropout, w.max_seq_len, w.target_len, w.prompt_len).eval()
    return w


def oracle_from_wrapper(w, dtype=np.float32):
    sd = {k: v.detach().cpu().numpy() for k, v in w.model.state_dict().items()}
    nz = w.normalizer
    return TransformerOracle(sd, w.nhead, w.num_decoder_layers, w.target_len, w.prompt_len,
                             nz.x_mean, nz.x_std, nz.u_mean, nz.u_std, dtype=dtype)


@pytest.mark.parametrize("system", ["cartpole", "quadrotor"])
@pytest.mark.parametrize("mode", ["fp32", "rand"])
def test_transformer_oracle_matches_reference_predict(system, mode, golden_dir):
    g = np.load(f"{golden_dir}/tf_forward_{system}.npz")
    w = _torch_predictor(system, mode)
    if mode == "rand":   # same RNG stream as the reference class
        cs = np.array([float(v.double().sum()) for v in w.model.state_dict().values()])
        assert np.array_equal(cs, g["rand_checksum"])
    orc = oracle_from_wrapper(w)
    pred = orc.predict_batch(g["x_err"], g["kK"])
    assert relmax(pred, g[f"pred_{mode}"]) < 2e-5
    # and the fp16 reference path is within its own quantisation gap of the fp32 one
    if mode == "fp32":
        assert relmax(pred, g["pred_fp16"]) < 5e-2      # measured: 1.8e-2 (quadrotor), the reference fp16-vs-fp32 gap


def test_positional_encoding_matches_checkpoint_buffer():
    w = _torch_predictor("quadrotor", "rand")
    pe = w.model.state_dict()["pos_encoder.pe"][0].numpy()
    assert np.allclose(sinusoid_pe(pe.shape[0], pe.shape[1]), pe, atol=2e-5)   # fp32 exp/sin: numpy vs torch differ by 7.6e-6


def test_transformer_oracle_matches_torch_module():
    """The NumPy restatement against torch's nn.TransformerEncoder on random inputs."""
    w = _torch_predictor("cartpole", "rand")
    orc = oracle_from_wrapper(w, dtype=np.float64)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 31, 4)).astype(np.float32)
    p = rng.standard_normal((3, 5, 5)).astype(np.float32)
    with torch.no_grad():
        ref = w.model.double()(torch.tensor(x).double(), torch.tensor(p).double()).numpy()
    assert relmax(orc.forward(x, p), ref) < 1e-10


@pytest.mark.parametrize("system", ["cartpole", "quadrotor"])
def test_tf_mode_port_matches_reference(system):
    """Transformer-mode control flow of the scalar port, with the NumPy transformer as predictor,
    against the reference run with the same (fp32 up-cast) weights."""
    runs = golden_runs(f"ilqr_{system}_tf_fp32")
    w = _torch_predictor(system, "fp32")
    orc = oracle_from_wrapper(w)
    model = make_model(system)
    N = HORIZON[system]
    r = runs[0]
    p = ILQRPort(model, r["x0"], [np.zeros(model.m) for _ in range(N)], N, tol=TOL[system], tf=orc)
    if system == "quadrotor":
        p.state_offset[2] = 0.5
    p.optimize(model.x_ref)
    assert p.total_iter == int(r["total_iter"])
    alphas = [(-1 if l["alpha"] is None else ob.ALPHAS.tolist().index(l["alpha"])) for l in p.logs]
    assert alphas == list(r["alpha_idx"])
    assert relmax([l["current_cost"] for l in p.logs], r["cost"]) < 1e-5


def test_int8_checkpoint_round_trip(tmp_path, monkeypatch):
    """quant_mode="int8" (transformer_ilqr.py:226-227, :297-298): save writes torch's dynamically-quantised state_dict,
    load de-quantises it; the loaded fp32 module must carry exactly the de-quantised int8 weights of the nn.Linear
    modules and the untouched fp32 tensors of everything else.  Host logic only (no GPU)."""
    import copy
    import os
    import torch
    from quattro_transformer_ilqr_b200 import TransformerILQR
    from quattro_transformer_ilqr_b200.transformer_model import TransformerPredictor
    w = TransformerILQR(4, 5, prompt_len=5, d_model=128, nhead=4, num_decoder_layers=3, dim_feedforward=256,
                        max_seq_len=100, quant_mode="int8")
    w.target_len = 25
    torch.manual_seed(3)
    w.model = TransformerPredictor(4, 5, 128, 4, 3, 256, 0.1, 100, 25, 5).eval()
    nz = w.normalizer
    nz.x_mean, nz.x_std = np.zeros(4, np.float32), np.ones(4, np.float32)
    nz.u_mean, nz.u_std = np.zeros(5, np.float32), np.ones(5, np.float32)
    monkeypatch.chdir(tmp_path)
    d = w.save("int8test")
    sd_file = torch.load(os.path.join(d, "tf_model.pt"), weights_only=True)
    assert any(k.endswith("._packed_params._packed_params") for k in sd_file)          # the reference's wire format
    r = TransformerILQR(1, 1).load(d)
    assert r.quant_mode == "int8" and r.target_len == 25 and r.prompt_len == 5
    ref_q = torch.quantization.quantize_dynamic(copy.deepcopy(w.model), {torch.nn.Linear}, dtype=torch.qint8)
    got, orig = r.model.state_dict(), w.model.state_dict()
    for name in ("state_embed", "control_embed", "output_linear", "transformer_decoder.layers.1.linear1",
                 "transformer_decoder.layers.2.linear2"):
        mod = ref_q
        for part in name.split("."):
            mod = getattr(mod, part)
        assert torch.equal(got[name + ".weight"], mod.weight().dequantize())
        assert torch.equal(got[name + ".bias"], mod.bias())
        assert not torch.equal(got[name + ".weight"], orig[name + ".weight"])           # really quantised
    for name in ("transformer_decoder.layers.0.self_attn.in_proj_weight", "transformer_decoder.layers.0.norm1.weight",
                 "target_embedding", "transformer_decoder.layers.2.self_attn.out_proj.weight"):
        assert torch.equal(got[name], orig[name])                                       # not nn.Linear: stays fp32
