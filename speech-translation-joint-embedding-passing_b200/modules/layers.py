"""Transformer blocks on b200st kernels — same classes, constructor/forward signatures and parameter
names as the reference's modules/layers.py, so checkpoints and callers are interchangeable.

Reference behaviour reproduced on purpose: LayerNorm is applied to the QUERY input only and
K/V are projected from the raw input (layers.py:153-160); masked scores are set to the finite -1e9
(layers.py:224); every LayerNorm here uses eps 1e-6.
"""
import math

import torch
import torch.nn as nn

from b200st import functional as BF
from b200st.kernels import K
from b200st.hostutil import PAD


def _p(module):
    """Active dropout probability of an nn.Dropout: its p in training mode, 0 otherwise."""
    return float(module.p) if module.training else 0.0


# Which LayerNorm normalises a sub-layer's OUTPUT next (the following sub-layer's pre-norm, or the stack's final norm)?
# The stack / layer forward passes note it on the sub-layer module for the duration of the call (a plain __dict__ entry, not
# a registered sub-module: state_dicts and pickles are unaffected), so that the sub-layer's last GEMM can normalise in its
# epilogue (b200st.functional.mha_block / ffn_block, csrc/gemm_ln.cu).  Forward signatures stay the reference's.
_NEXT_LN = '_b200st_next_ln'


def _ln_args(ln):
    return None if ln is None else (ln.weight, ln.bias, ln.eps)


class _next_ln:
    def __init__(self, module, ln_args):
        self.module, self.ln_args = module, ln_args

    def __enter__(self):
        self.module.__dict__[_NEXT_LN] = self.ln_args

    def __exit__(self, *exc):
        self.module.__dict__.pop(_NEXT_LN, None)
        return False


class TransformerEncoderLayer(nn.Module):
    """layers.py:23-63"""

    def __init__(self, dim_model, nhead, dim_feedforward, d_k, d_v, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiheadAttention(nhead, dim_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(dim_model, dim_feedforward, dropout=dropout)

    def forward(self, src, slf_attn_mask=None, prior_weight=None):
        with _next_ln(self.slf_attn, _ln_args(self.pos_ffn.layer_norm)):
            y, att = self.slf_attn(src, src, src, mask=slf_attn_mask, prior_weight=prior_weight)
        with _next_ln(self.pos_ffn, self.__dict__.get(_NEXT_LN)):
            return self.pos_ffn(y), att


class TransformerDecoderLayer(nn.Module):
    """layers.py:66-112"""

    def __init__(self, dim_model, nhead, dim_feedforward, d_k, d_v, dropout=0.1):
        super().__init__()
        self.decslf_attn = MultiheadAttention(nhead, dim_model, d_k, d_v, dropout=dropout)
        self.encdec_attn = MultiheadAttention(nhead, dim_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(dim_model, dim_feedforward, dropout=dropout)

    def forward(self, dec_input, enc_output, decslf_attn_mask=None, encdec_attn_mask=None,
                decode_speedup=False, cache_decslf=None, cache_encdec=None):
        if decode_speedup:
            raise NotImplementedError('decode_speedup is never used by Seq2seq (SURVEY.md §2.1 #6)')
        with _next_ln(self.decslf_attn, _ln_args(self.encdec_attn.layer_norm)):
            y, att_decslf = self.decslf_attn(dec_input, dec_input, dec_input, mask=decslf_attn_mask)
        with _next_ln(self.encdec_attn, _ln_args(self.pos_ffn.layer_norm)):
            y, att_encdec = self.encdec_attn(y, enc_output, enc_output, mask=encdec_attn_mask)
        with _next_ln(self.pos_ffn, self.__dict__.get(_NEXT_LN)):
            return self.pos_ffn(y), att_decslf, att_encdec


class MultiheadAttention(nn.Module):
    """layers.py:120-197.  Parameter names: w_qs, w_ks, w_vs, fc (no bias), layer_norm (eps 1e-6)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.fc = nn.Linear(n_head * d_v, d_model, bias=False)
        self.attention = ScaledDotProductAttention(temperature=d_k ** 0.5)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)

    def forward(self, q, k, v, mask=None, prior_weight=None, decode_speedup=False, cache=None):
        if prior_weight is not None or decode_speedup:
            raise NotImplementedError('prior_weight / decode_speedup paths are unused by Seq2seq')
        assert self.d_k == self.d_v, 'the fused attention core assumes d_k == d_v (always true upstream)'
        # dropout on the fc output (layers.py:194) and the hard-wired p = 0.1 attention dropout (layers.py:207,226)
        p_fc, p_attn = _p(self.dropout), _p(self.attention.dropout)
        tag = getattr(self, '_b200st_tag', '')
        if mask is not None and mask.dtype == torch.bool:
            mask = mask.view(torch.uint8) if mask.is_contiguous() else mask.to(torch.uint8)
        if k is v:           # every call site of the reference (self- and cross-attention): one fused sub-layer node
            return BF.mha_block(q, k, mask, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps,
                                self.w_qs.weight, self.w_ks.weight, self.w_vs.weight, self.fc.weight, self.n_head,
                                self.attention.temperature, p_attn=p_attn, p_fc=p_fc, tag=tag,
                                next_ln=self.__dict__.get(_NEXT_LN))
        residual = q
        qn = BF.layer_norm(q, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        qp = BF.linear(qn, self.w_qs.weight)
        kp = BF.linear(k, self.w_ks.weight)
        vp = BF.linear(v, self.w_vs.weight)
        o, attn = BF.mha_core(qp, kp, vp, mask, self.n_head, self.attention.temperature, p_attn=p_attn, tag=tag)
        if p_fc > 0:
            out = BF.dropout(BF.linear(o, self.fc.weight), p_fc, True, tag + '.fc', residual=residual)
        else:
            out = BF.linear(o, self.fc.weight, residual=residual)  # fc, += residual
        return out, attn


class ScaledDotProductAttention(nn.Module):
    """layers.py:200-229 — kept as a module for its attributes (temperature, dropout); the computation
    lives in the fused attention kernel called by MultiheadAttention."""

    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)

    def forward(self, q, k, v, mask=None, prior_weight=None):
        # q,k,v: [B, H, L, d] as in the reference; mask [B,1,1|Lq,Lk]
        if prior_weight is not None:
            raise NotImplementedError('prior_weight is unused by Seq2seq')
        B, H, Lq, d = q.shape
        qq = q.transpose(1, 2).reshape(B, Lq, H * d)
        kk = k.transpose(1, 2).reshape(B, k.size(2), H * d)
        vv = v.transpose(1, 2).reshape(B, v.size(2), H * d)
        m = None if mask is None else mask.reshape(B, -1, k.size(2)).to(torch.uint8)
        o, attn = BF.mha_core(qq, kk, vv, m, H, self.temperature, p_attn=_p(self.dropout),
                              tag=getattr(self, '_b200st_tag', ''))
        return o.view(B, Lq, H, d).transpose(1, 2), attn


class PositionwiseFeedForward(nn.Module):
    """layers.py:232-252: x + w_2(relu(w_1(LN(x))))."""

    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.layer_norm = nn.LayerNorm(d_in, eps=1e-6)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        return BF.ffn_block(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps,
                            self.w_1.weight, self.w_1.bias, self.w_2.weight, self.w_2.bias, p=_p(self.dropout),
                            tag=getattr(self, '_b200st_tag', ''), next_ln=self.__dict__.get(_NEXT_LN))


# ---- helpers (layers.py:260-309) ---------------------------------------------------------------
def _get_zero_mask(seq):
    return (seq != 0).unsqueeze(-2)


def _get_pad_mask(seq):
    return (seq != PAD).unsqueeze(-2)


def _get_subsequent_mask(max_length):
    return (1 - torch.triu(torch.ones((1, max_length, max_length)), diagonal=1)).type(torch.bool)


def _gen_position_signal(max_len, d_model):
    """Sinusoid table [1, max_len, d_model]: sin on even, cos on odd columns (layers.py:292-309).
    Built with the same torch expression as the reference so the constants are bit-identical."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).clone().detach()


def position_signal(module):
    """The PositionSignal cache of an Encoder / Decoder; created on first use for instances that never ran this
    package's constructor (whole-module pickles written by the reference's classes, checkpoint.py:76)."""
    pe = module.__dict__.get('_pe')
    if pe is None:
        pe = module.__dict__['_pe'] = PositionSignal()
    return pe


class PositionSignal:
    """Device-resident copy of a time-signal table: the reference re-uploads it on every call
    (`.type_as`, TFEnc.py:82-83); here it is uploaded once per device and fused into one add kernel."""

    def __init__(self):
        self._dev = {}

    def __getstate__(self):          # whole-module pickles (checkpoint.py:76) must not carry device tensors of a cache
        return {'_dev': {}}

    def on(self, table, device):
        key = (str(device), table.data_ptr(), table.size(1))
        hit = self._dev.get(key)
        if hit is None:
            hit = table[0].to(device=device, dtype=torch.float32).contiguous()
            self._dev = {key: hit}
        return hit
