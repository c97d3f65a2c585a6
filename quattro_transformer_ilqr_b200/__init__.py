"""B200-native batched iLQR engine behind the ``quattro_ilqr_tf`` API.

Public surface (same names as the reference package, quattro_ilqr_tf/__init__.py:1-3):
``iLQR_TF``, ``TransformerILQR``, ``TransformerPredictor``, ``DataNormalizer`` — plus the batched
engine (``BatchedILQR``), the device-model descriptor (``DeviceModel``) and the MPC front-ends
(``quattro_transformer_ilqr_b200.mpc``).
"""
from .device_models import DeviceModel, resolve
from .engine import BatchedILQR
from .quattro_ilqr_tf import iLQR_TF
from .transformer_model import DataNormalizer, PositionalEncoding, TransformerPredictor
from .transformer_ilqr import TransformerILQR

__all__ = ["iLQR_TF", "TransformerILQR", "TransformerPredictor", "DataNormalizer", "PositionalEncoding",
           "BatchedILQR", "DeviceModel", "resolve"]
