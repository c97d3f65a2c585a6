"""Alias of quattro_transformer_ilqr_b200.transformer_model (the reference's module path, e.g. the examples do
``from quattro_ilqr_tf.quattro_ilqr_tf import iLQR_TF``)."""
from quattro_transformer_ilqr_b200.transformer_model import DataNormalizer, PositionalEncoding, TransformerPredictor  # noqa: F401
