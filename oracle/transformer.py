"""CPU oracle: NumPy restatement of the Quattro gain-predicting transformer.  TEST INFRASTRUCTURE ONLY.

Follows (paths relative to /root/reference):
  quattro_ilqr_tf/transformer_model.py:37-50   DataNormalizer.transform_x/u, inverse_transform_u
  quattro_ilqr_tf/transformer_model.py:60-80   sinusoidal positional encoding (buffer `pos_encoder.pe`)
  quattro_ilqr_tf/transformer_model.py:97-138  TransformerPredictor.forward
  quattro_ilqr_tf/transformer_ilqr.py:311-325  TransformerILQR.predict
The encoder layers are torch's nn.TransformerEncoderLayer (torch 2.6.0 pinned by the reference's
requirements.txt:23; not part of /root/reference) with its defaults: post-LayerNorm, ReLU,
LayerNorm eps 1e-5 with biased variance, softmax(QK^T/sqrt(head_dim) + causal -inf mask), dropout
inactive in eval.  Pinned against the real torch module in tests/test_oracle.py and against
reference outputs in tests/golden/.

Weights are given as a dict of NumPy arrays keyed like the reference's ``state_dict``.
"""
from __future__ import annotations

import numpy as np


def sinusoid_pe(max_len, d_model):
    """transformer_model.py:64-71, computed in float32 like the reference buffer."""
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, d_model, 2, dtype=np.float32) * np.float32(-np.log(np.float32(10000.0)) / d_model))
    pe = np.zeros((max_len, d_model), np.float32)
    pe[:, 0::2] = np.sin(pos * div)
    pe[:, 1::2] = np.cos(pos * div)
    return pe


def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


class TransformerOracle:
    def __init__(self, weights, nhead, num_layers, target_len, prompt_len,
                 x_mean=None, x_std=None, u_mean=None, u_std=None, dtype=np.float32):
        self.dtype = dtype
        self.w = {k: np.asarray(v).astype(dtype) for k, v in weights.items()}
        self.nhead, self.num_layers = nhead, num_layers
        self.target_len, self.prompt_len = target_len, prompt_len
        self.x_mean, self.x_std, self.u_mean, self.u_std = x_mean, x_std, u_mean, u_std
        self.d = self.w["state_embed.weight"].shape[0]

    def forward(self, x_norm, prompt_norm):
        """(B, T, n), (B, P, c) -> (B, target_len, c); transformer_model.py:122-138."""
        w, dt = self.w, self.dtype
        x_norm = np.asarray(x_norm).astype(dt)
        prompt_norm = np.asarray(prompt_norm).astype(dt)
        B = x_norm.shape[0]
        xe = x_norm @ w["state_embed.weight"].T + w["state_embed.bias"]
        ue = prompt_norm @ w["control_embed.weight"].T + w["control_embed.bias"]
        tgt = np.broadcast_to(w["target_embedding"], (B,) + w["target_embedding"].shape)
        h = np.concatenate([xe, ue, tgt], axis=1)
        L = h.shape[1]
        h = h + w["pos_encoder.pe"].reshape(-1, self.d)[:L]
        mask = np.triu(np.ones((L, L), bool), 1)
        hd = self.d // self.nhead
        for l in range(self.num_layers):
            p = f"transformer_decoder.layers.{l}."
            qkv = h @ w[p + "self_attn.in_proj_weight"].T + w[p + "self_attn.in_proj_bias"]
            q, k, v = np.split(qkv, 3, axis=-1)
            sp = lambda z: z.reshape(B, L, self.nhead, hd).transpose(0, 2, 1, 3)
            q, k, v = sp(q), sp(k), sp(v)
            s = (q @ k.transpose(0, 1, 3, 2)) * dt(1.0 / np.sqrt(hd))
            s = np.where(mask, -np.inf, s)
            s = s - s.max(-1, keepdims=True)
            e = np.exp(s)
            a = e / e.sum(-1, keepdims=True)
            o = (a @ v).transpose(0, 2, 1, 3).reshape(B, L, self.d)
            o = o @ w[p + "self_attn.out_proj.weight"].T + w[p + "self_attn.out_proj.bias"]
            h = _layer_norm(h + o, w[p + "norm1.weight"], w[p + "norm1.bias"])
            ff = np.maximum(h @ w[p + "linear1.weight"].T + w[p + "linear1.bias"], 0)
            ff = ff @ w[p + "linear2.weight"].T + w[p + "linear2.bias"]
            h = _layer_norm(h + ff, w[p + "norm2.weight"], w[p + "norm2.bias"])
        return h[:, -self.target_len:] @ w["output_linear.weight"].T + w["output_linear.bias"]

    def predict_batch(self, x_seq, kK_seq):
        """Batched transformer_ilqr.py:311-325: normalise (fp64 -> fp32), forward, de-normalise."""
        x_norm = ((np.asarray(x_seq) - self.x_mean) / self.x_std).astype(np.float32)
        kK_norm = ((np.asarray(kK_seq) - self.u_mean) / self.u_std)[:, -self.prompt_len:].astype(np.float32)
        pred = self.forward(x_norm, kK_norm).astype(np.float32)
        return pred * self.u_std + self.u_mean

    def predict(self, x_seq, kK_seq):
        return self.predict_batch(x_seq[None], kK_seq[None])[0]
