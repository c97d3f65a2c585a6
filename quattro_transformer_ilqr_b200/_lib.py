"""ctypes binding of libquattro_sm100a.so (the C ABI declared in include/quattro_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If it has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``csrc/build.sh``) loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquattro_sm100a.so")

QX_MODEL_CARTPOLE, QX_MODEL_QUADROTOR = 0, 1
QX_EULER, QX_RK4 = 0, 1
QX_LAYOUT_SOA, QX_LAYOUT_REC, QX_LAYOUT_BLK = 0, 1, 2
QX_MAX_N, QX_MAX_M, QX_NUM_ALPHA = 12, 4, 6
QX_LS_FORCE_ACTIVE = 1
ALPHAS = (1.0, 0.5, 0.25, 0.1, 0.05, 0.01)


class QxParams(C.Structure):
    _fields_ = [
        ("model_id", C.c_int32), ("integrator", C.c_int32), ("n", C.c_int32), ("m", C.c_int32),
        ("diag_cost", C.c_int32), ("_pad", C.c_int32),
        ("dt", C.c_double), ("phys", C.c_double * 8),
        ("Q", C.c_double * (QX_MAX_N * QX_MAX_N)), ("R", C.c_double * (QX_MAX_M * QX_MAX_M)),
        ("Qf", C.c_double * (QX_MAX_N * QX_MAX_N)), ("x_ref", C.c_double * QX_MAX_N),
        ("barrier_alpha", C.c_double), ("barrier_beta", C.c_double),
    ]


class QxTfConfig(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("state_dim", "control_dim", "d_model", "nhead", "num_layers", "dim_ff",
                                         "target_len", "prompt_len", "horizon", "m")]


# name -> (restype, argtypes); must list every symbol declared in include/quattro_b200.h
_P = C.c_void_p
SIGNATURES = {
    "qx_abi_version": (C.c_int, []),
    "qx_last_error": (C.c_char_p, []),
    "qx_layout": (C.c_int, [C.c_int]),
    "qx_record_floats": (C.c_int, [C.c_int]),
    "qx_expansion_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "qx_terminal_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "qx_gains_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "qx_rollout_cost": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "qx_linearize": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "qx_riccati": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "qx_blk_supported": (C.c_int, [C.POINTER(QxParams)]),
    "qx_blk_expansion_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "qx_blk_terminal_bytes": (C.c_size_t, [C.c_int]),
    "qx_linearize_blk": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "qx_riccati_blk": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "qx_linesearch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "qx_linesearch": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_double, C.c_uint32,
                                C.c_int, C.c_int, _P]),
    "qx_linesearch_stage_events": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "qx_forward_pass": (C.c_int, [C.POINTER(QxParams), _P, _P, _P, C.c_double, _P, _P, _P, C.c_int, C.c_int, _P]),
    "qx_tf_weight_floats": (C.c_size_t, [C.POINTER(QxTfConfig)]),
    "qx_tf_workspace_bytes": (C.c_size_t, [C.POINTER(QxTfConfig), C.c_int]),
    "qx_tf_split_floats": (C.c_size_t, [C.POINTER(QxTfConfig)]),
    "qx_tf_prepare": (C.c_int, [C.POINTER(QxTfConfig), _P, _P, _P]),
    "qx_tf_forward": (C.c_int, [C.POINTER(QxTfConfig), _P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P]),
    "qx_tf_attention": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "qx_tf_split": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
    "qx_tf_dense": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "qx_tf_dense_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
}

_lib = None


def load():
    """Load the shared library once; raise if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.qx_abi_version() != 1:
        raise RuntimeError("libquattro_sm100a.so ABI version mismatch")
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {load().qx_last_error().decode()}")
