"""CPU tests: the C-ABI library builds, loads and exports every symbol include/quattro_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "quattro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qx_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from quattro_transformer_ilqr_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from quattro_transformer_ilqr_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 13
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_layout_queries(lib):
    from quattro_transformer_ilqr_b200 import _lib
    assert lib.qx_abi_version() == 1
    assert lib.qx_record_floats(_lib.QX_MODEL_QUADROTOR) == 416        # 1 664 B per (t, b), SURVEY.md 8d
    assert lib.qx_record_floats(_lib.QX_MODEL_CARTPOLE) == 46
    assert lib.qx_layout(_lib.QX_MODEL_QUADROTOR) == _lib.QX_LAYOUT_REC
    assert lib.qx_layout(_lib.QX_MODEL_CARTPOLE) == _lib.QX_LAYOUT_SOA
    # bytes per trajectory quoted in SURVEY.md 8(d): expansions + terminal = 83 824 B, gains 10 400 B
    assert lib.qx_expansion_bytes(1, 1, 50) + lib.qx_terminal_bytes(1, 1) == 83824
    assert lib.qx_gains_bytes(1, 1, 50) == 10400
    assert lib.qx_layout(7) < 0 and b"unknown model_id" in lib.qx_last_error()
    # line-search workspaces: one int32 winner hint per cart-pole trajectory (128-B granules); the quadrotor keeps a
    # queue, two flag arrays, the nominal's cost and the five probe costs per trajectory (no candidate is parked)
    assert lib.qx_linesearch_workspace_bytes(_lib.QX_MODEL_CARTPOLE, 4096, 30) == 4096 * 4
    assert lib.qx_linesearch_workspace_bytes(_lib.QX_MODEL_CARTPOLE, 1, 30) == 128
    assert lib.qx_linesearch_workspace_bytes(_lib.QX_MODEL_QUADROTOR, 1024, 50) == 32 + 3 * 4096 + 6 * 8192
    assert lib.qx_linesearch_workspace_bytes(_lib.QX_MODEL_CARTPOLE, 0, 30) == 0


def test_block_path_queries(lib):
    """The compact block records (QX_LAYOUT_BLK): 192 B per (t, b) in octets of trajectories, 64-B terminal; only
    the quadrotor with the Euler map and diagonal weights qualifies."""
    from quattro_transformer_ilqr_b200 import DeviceModel
    assert lib.qx_blk_expansion_bytes(16384, 50) == 16384 * 50 * 192
    assert lib.qx_blk_expansion_bytes(9, 50) == 16 * 50 * 192          # ragged batch: padded to whole octets
    assert lib.qx_blk_terminal_bytes(97) == 97 * 64
    assert lib.qx_blk_expansion_bytes(0, 50) == 0
    ok = DeviceModel("quadrotor").params()
    assert lib.qx_blk_supported(ctypes.byref(ok)) == 1
    assert lib.qx_blk_supported(ctypes.byref(DeviceModel("quadrotor", integration_method="rk4").params())) == 0
    Q = np.diag([10.0] * 12); Q[0, 1] = Q[1, 0] = 0.5
    assert lib.qx_blk_supported(ctypes.byref(DeviceModel("quadrotor", Q=Q).params())) == 0
    assert lib.qx_blk_supported(ctypes.byref(DeviceModel("cartpole").params())) == 0
    # argument validation needs no GPU: the wrong model is refused before any launch
    bad = DeviceModel("cartpole").params()
    assert lib.qx_riccati_blk(ctypes.byref(bad), None, None, None, None, 8, 50, 0, 1e-6, None) == -1


def test_params_struct_matches_header():
    from quattro_transformer_ilqr_b200 import _lib
    # 6 int32 + (1 + 8 + 144 + 16 + 144 + 12 + 2) doubles
    assert ctypes.sizeof(_lib.QxParams) == 6 * 4 + 327 * 8
    assert ctypes.sizeof(_lib.QxTfConfig) == 10 * 4


def test_argument_validation_without_gpu(lib):
    from quattro_transformer_ilqr_b200 import _lib
    p = _lib.QxParams()
    assert lib.qx_rollout_cost(ctypes.byref(p), None, None, None, None, None, 4, 4, None) == -1
    assert b"null pointer" in lib.qx_last_error()


def test_device_model_registry():
    from quattro_transformer_ilqr_b200 import DeviceModel, resolve
    from quattro_transformer_ilqr_b200.device_models import _OwnerModel

    class Dyn:
        m_cart, m_pole, length, gravity = 1.0, 0.1, 0.15, 9.81

    class Owner:
        dynamics = Dyn()
        dt, integration_method = 0.01, "euler"
        Q, R, Qf, x_ref = np.eye(4), np.eye(1), np.eye(4), np.zeros(4)

        def f(self, x, u):
            return x

    o = Owner()
    dm = resolve(o.f, o.f, o.f)
    assert isinstance(dm, _OwnerModel) and dm.name == "cartpole"
    p = dm.params()
    assert p.diag_cost == 1 and p.n == 4 and p.m == 1 and p.Q[5] == 1.0
    o.x_ref = np.array([1.0, 0, 0, 0])          # callers mutate x_ref between solves
    o.Q = np.ones((4, 4))
    p = dm.params()
    assert p.x_ref[0] == 1.0 and p.diag_cost == 0
    o.integration_method = "bogus"
    with pytest.raises(ValueError, match="Unknown integration method"):
        dm.params()
    with pytest.raises(TypeError, match="no CPU fallback"):
        resolve(lambda x, u: x)
    q = DeviceModel("quadrotor", mass=1.2)
    assert q.params().phys[0] == 1.2 and q.params().barrier_alpha == 1000.0
    with pytest.raises(ValueError):
        DeviceModel("pendulum")


def test_transformer_weight_packing(lib):
    from quattro_transformer_ilqr_b200 import TransformerILQR
    from tests.util import checkpoint_dir
    for system, n_params in (("cartpole", 402693), ("quadrotor", 616244)):
        w = TransformerILQR(1, 1).load(checkpoint_dir(system))
        assert sum(p.numel() for p in w.model.parameters()) == n_params      # "402.7k" / "616.2k"
        flat = w.pack_weights()
        assert flat.dtype.is_floating_point and flat.numel() == lib.qx_tf_weight_floats(ctypes.byref(w.config()))


def test_no_cpu_fallback():
    import torch
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedILQR(DeviceModel("cartpole"), 4, 30)


def test_ilqr_constructor_errors():
    from quattro_transformer_ilqr_b200 import DeviceModel, iLQR_TF
    dm = DeviceModel("cartpole")
    with pytest.raises(ValueError, match="tf_window must be less than the horizon"):
        iLQR_TF(dm, None, None, np.zeros(4), [np.zeros(1)] * 5, 5, tf_window=5)
    s = iLQR_TF(dm, None, None, np.zeros(4), [np.zeros(1)] * 30, 30)
    assert s.get_state_offset().shape == (4,) and len(s.get_time()) == 3


def test_bench_arms_share_the_config_object_and_the_profile_parser_knows_every_launch():
    """bench.py: both arms print the same `config` (the driver compares them); the reference arm's line carries the keys
    of the contract; the committed ncu summary has a block for each of the five launches `roofline` may name."""
    import json
    import subprocess
    import sys
    import bench
    for wl in bench.WORKLOADS:
        assert bench.workload_config(wl, 2) == bench.workload_config(wl, 2)
        c = bench.workload_config(wl, 4)
        assert c["global_batch"] == 4 * bench.WORKLOADS[wl][2] and set(c) == {"workload", "global_batch", "parallelism", "l2"}
    for k in ("linearize", "riccati", "ls_phaseA", "ls_phaseB", "ls_phaseW"):
        ref = bench.ncu_reference(k, True, "quadrotor")
        assert ref is not None and ref["traffic"] > 0 and ref["duration_us"] > 0, k
    assert bench.ncu_reference("ls_phaseB", True, "quadrotor")["fp64_pipe_pct"] > 10
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cartpole_classic"],
                         capture_output=True, text=True, timeout=600, cwd=bench.ROOT)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.workload_config("cartpole_classic", 1)
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["value"] == line["value"] > 0
