"""Transformer encoder stack (reference: models/TFEnc.py:19-100), 'standard' type.

Same constructor / forward(src, src_mask) -> (x, att) / expand_time(); same parameter names including the
never-executed template layer `enc` (TFEnc.py:51-58), which exists so that state_dicts interchange.
The 'universal' / ACT variants are broken upstream (Act.py:28) and are rejected."""
import copy

import torch
import torch.nn as nn

from b200st import functional as BF
from modules.layers import (TransformerEncoderLayer, _gen_position_signal, PositionSignal, position_signal, _next_ln,
                            _ln_args)


class Encoder(nn.Module):

    def __init__(self, dim_model=200, dim_feedforward=512, num_heads=8, num_layers=6, act=False,
                 dropout=0.2, transformer_type='standard'):
        super().__init__()
        if act or transformer_type != 'standard':
            raise NotImplementedError("only transformer_type='standard', act=False is implemented")
        upperbound_seq_len = 500
        self.layer_signal = _gen_position_signal(num_layers, dim_model)
        self.time_signal = _gen_position_signal(upperbound_seq_len, dim_model)
        self.dim_model = dim_model
        self.dim_feedforward = dim_feedforward
        self.d_k = int(dim_model / num_heads)
        self.d_v = int(dim_model / num_heads)
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.act = act
        self.transformer_type = transformer_type
        self.enc = TransformerEncoderLayer(dim_model, num_heads, dim_feedforward, self.d_k, self.d_v, dropout)
        self.enc_layers = _get_clones(self.enc, num_layers)
        self.norm = nn.LayerNorm(dim_model, eps=1e-6)
        self._pe = PositionSignal()

    def expand_time(self, max_seq_len):
        self.time_signal = _gen_position_signal(max_seq_len, self.dim_model)

    def forward(self, src, src_mask=None):
        assert src.shape[1] <= self.time_signal.shape[1], 'call expand_time() for longer sequences'
        x = BF.add_posenc(src, position_signal(self).on(self.time_signal, src.device))     # TFEnc.py:82-83
        att = None
        n = len(self.enc_layers)
        for i, layer in enumerate(self.enc_layers):
            # the LayerNorm that normalises this layer's output next: the following layer's pre-norm, or the final norm
            nxt = self.enc_layers[i + 1].slf_attn.layer_norm if i + 1 < n else self.norm
            with _next_ln(layer, _ln_args(nxt)):
                x, att = layer(x, slf_attn_mask=src_mask)
        x = BF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x, att


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])
