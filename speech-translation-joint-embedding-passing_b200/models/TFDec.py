"""Transformer decoder stack (reference: models/TFDec.py:19-141), 'standard' type.

forward(tgt, memory, tgt_mask, src_mask, ...) -> (x, att_decslf, att_encdec).  The final LayerNorm keeps
PyTorch's default eps 1e-5 — unlike every other LayerNorm on the path (TFDec.py:58 vs TFEnc.py:61)."""
import copy

import torch
import torch.nn as nn

from b200st import functional as BF
from modules.layers import (TransformerDecoderLayer, _gen_position_signal, PositionSignal, position_signal, _next_ln,
                            _ln_args)


class Decoder(nn.Module):

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
        self.dec = TransformerDecoderLayer(dim_model, num_heads, dim_feedforward, self.d_k, self.d_v, dropout)
        self.dec_layers = _get_clones(self.dec, num_layers)
        self.norm = nn.LayerNorm(dim_model)                                   # eps 1e-5 (TFDec.py:58)
        self._pe = PositionSignal()

    def expand_time(self, max_seq_len):
        self.time_signal = _gen_position_signal(max_seq_len, self.dim_model)

    def forward(self, tgt, memory, tgt_mask=None, src_mask=None, decode_speedup=False,
                cache_decslf=None, cache_encdec=None):
        if decode_speedup:
            raise NotImplementedError('decode_speedup is never used by Seq2seq (SURVEY.md §2.1 #6)')
        assert tgt.shape[1] <= self.time_signal.shape[1], 'call expand_time() for longer sequences'
        x = BF.add_posenc(tgt, position_signal(self).on(self.time_signal, tgt.device))     # TFDec.py:85-86
        att_decslf = att_encdec = None
        n = len(self.dec_layers)
        for i, layer in enumerate(self.dec_layers):
            nxt = self.dec_layers[i + 1].decslf_attn.layer_norm if i + 1 < n else self.norm     # see TFEnc.Encoder.forward
            with _next_ln(layer, _ln_args(nxt)):
                x, att_decslf, att_encdec = layer(x, memory, decslf_attn_mask=tgt_mask,
                                                  encdec_attn_mask=src_mask)
        x = BF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return x, att_decslf, att_encdec


def _get_clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])
