"""Weight containers of the Quattro transformer (mirror of quattro_ilqr_tf/transformer_model.py).

``TransformerPredictor`` is kept as a ``torch.nn.Module`` with the reference's sub-module names and
construction order, so that (i) ``state_dict`` keys are the reference's checkpoint wire format
 and shipped checkpoints load unchanged, (ii) ``torch.manual_seed(s)`` yields the
same random initialisation as the reference class.  Inference on the hot path does NOT go through
``forward`` — ``TransformerILQR`` packs the weights once and runs the CUDA kernels
(csrc/qx_tf_*.cu); ``forward`` exists for training (``fit``) only.

  DataNormalizer        transformer_model.py:15-50
  PositionalEncoding    transformer_model.py:55-80
  TransformerPredictor  transformer_model.py:85-138
"""
from __future__ import annotations

import torch
import torch.nn as nn


class DataNormalizer:
    def __init__(self, eps=1e-6):
        self.eps = eps
        self.x_mean = self.x_std = self.u_mean = self.u_std = None

    def fit(self, x_data, u_data):
        self.x_mean = x_data.mean(axis=(0, 1))
        self.x_std = x_data.std(axis=(0, 1)) + self.eps
        self.u_mean = u_data.mean(axis=(0, 1))
        self.u_std = u_data.std(axis=(0, 1)) + self.eps

    def transform_x(self, x):
        return (x - self.x_mean) / self.x_std

    def transform_u(self, u):
        return (u - self.u_mean) / self.u_std

    def inverse_transform_u(self, u):
        return u * self.u_std + self.u_mean


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=100):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
        freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-torch.log(torch.tensor(10000.0)) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(position * freq)
        table[:, 1::2] = torch.cos(position * freq)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, x):
        return self.dropout(x + self.pe[:, : x.size(1)])


class TransformerPredictor(nn.Module):
    def __init__(self, state_dim, control_dim, d_model=64, nhead=8, num_decoder_layers=3, dim_feedforward=128,
                 dropout=0.1, max_seq_len=100, target_len=20, prompt_len=10):
        super().__init__()
        self.d_model, self.target_len, self.prompt_len = d_model, target_len, prompt_len
        # creation order matters: it fixes the RNG stream of the random initialisation
        self.state_embed = nn.Linear(state_dim, d_model)
        self.control_embed = nn.Linear(control_dim, d_model)
        self.pos_encoder = PositionalEncoding(d_model, dropout, max_seq_len)
        self.output_linear = nn.Linear(d_model, control_dim)
        layer = nn.TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, batch_first=True)
        self.transformer_decoder = nn.TransformerEncoder(layer, num_decoder_layers)
        self.target_embedding = nn.Parameter(torch.zeros(target_len, d_model))
        nn.init.normal_(self.target_embedding, std=0.02)

    def forward(self, x_seq, u_prompt):
        tokens = torch.cat([self.state_embed(x_seq), self.control_embed(u_prompt),
                            self.target_embedding.unsqueeze(0).expand(x_seq.size(0), -1, -1)], dim=1)
        tokens = self.pos_encoder(tokens)
        L = tokens.size(1)
        mask = torch.triu(torch.ones(L, L, device=tokens.device), diagonal=1).bool()
        out = self.transformer_decoder(tokens, mask=mask)
        return self.output_linear(out[:, -self.target_len:, :])
