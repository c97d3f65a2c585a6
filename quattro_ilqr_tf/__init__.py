"""Drop-in alias: ``import quattro_ilqr_tf`` resolves to the B200 engine with the reference's names
(reference package surface: quattro_ilqr_tf/__init__.py:1-3)."""
from quattro_transformer_ilqr_b200 import (BatchedILQR, DataNormalizer, DeviceModel, TransformerILQR,  # noqa: F401
                                           TransformerPredictor, iLQR_TF)
