"""Sinusoidal positional encoding buffer -- same module name / state_dict key ('pe', shape
(max_len, 1, d_model)) as model/positional_encoding.py:8-23 of the reference.  The addition itself
is fused into the GEMM epilogue of Linear_chord / Linear_vis (engine.chord_stream / encode_memory);
forward() is kept for callers that use the module directly."""
import math

import torch
import torch.nn as nn


class PositionalEncoding(nn.Module):
    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout_p = dropout
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1).contiguous())

    def forward(self, x):
        if self.training and self.dropout_p > 0:
            raise NotImplementedError("dropout > 0 in training mode is not built yet (use dropout=0.0 or eval())")
        return x + self.pe[: x.size(0), :]
