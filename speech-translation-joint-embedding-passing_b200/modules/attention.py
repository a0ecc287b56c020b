"""LAS single-head attention — mirror of the reference's modules/attention.py AttentionLayer.

Only mode='bilinear' is ever constructed upstream (Seq2seq.py:150, Dec.py:86-93); the other modes
(bahdanau / hybrid / dot_prod) are never instantiated and are rejected here.
Semantics kept: score = q . (W k_j) (attention.py:190-193), masked positions filled with the finite -1e12
(attention.py:250-252), softmax over keys, context = P . values (attention.py:268-273).
"""
import torch
import torch.nn as nn

from b200st import functional as BF
from b200st.kernels import K
from b200st import runtime as rt


class _BilinearAttention(torch.autograd.Function):
    """Stand-alone (non-hoisted) attention call for users of AttentionLayer.forward; the training loop
    inside Dec uses the fused decoder function with the key projection hoisted."""

    @staticmethod
    def forward(ctx, query, keys, values, weight, klens):
        k = K()
        b, tq, nq = query.shape
        tk = keys.size(1)
        assert tq == 1, 'b200st AttentionLayer supports the single-query (decoding) form'
        keys_c, vals_c = keys.contiguous(), values.contiguous()
        wk = k.gemm(keys_c.view(b * tk, -1), rt.operand(weight), trans_b=True).view(b, tk, nq)
        q2 = query.reshape(b, nq).contiguous()
        cx, probs = k.las_attn_fwd(q2, wk, vals_c, klens)
        ctx.save_for_backward(q2, keys_c, vals_c, wk, probs, weight)
        ctx.same_kv = keys.data_ptr() == values.data_ptr()
        return cx.view(b, 1, -1), probs.view(b, 1, tk).to(query.dtype)

    @staticmethod
    def backward(ctx, dctx, _dp):
        k = K()
        q2, keys, vals, wk, probs, weight = ctx.saved_tensors
        b, tk, nq = wk.shape
        dt = q2.dtype
        dscore, dq = k.las_attn_bwd(dctx.reshape(b, -1).contiguous(), wk, vals, probs)
        dsc = k.cast(dscore, dt).view(b, 1, tk)
        prb = k.cast(probs, dt).view(b, 1, tk)
        d_wk = k.gemm(dsc, q2.view(b, 1, nq), trans_a=True)                 # [b, tk, nq]
        d_vals = k.gemm(prb, dctx.reshape(b, 1, -1).contiguous(), trans_a=True)
        d_keys = k.gemm(d_wk.view(b * tk, nq), rt.operand(weight)).view(keys.shape)
        dw = k.gemm(d_wk.view(b * tk, nq), keys.view(b * tk, -1), trans_a=True, out_dtype=torch.float32)
        return dq.view(b, 1, nq), d_keys, d_vals, dw, None


class AttentionLayer(nn.Module):

    def __init__(self, query_size, key_size, value_size=None, mode='bahdanau', dropout=0.0,
                 batch_first=True, bias=True, query_transform=False, output_transform=False,
                 output_nonlinearity='tanh', output_size=None, hidden_size=1, hard_att=False):
        super().__init__()
        if mode != 'bilinear' or query_transform or output_transform or hard_att or not batch_first:
            raise NotImplementedError(
                "b200st AttentionLayer implements mode='bilinear' without query/output transforms, the only "
                "configuration the joint-ST model constructs (Seq2seq.py:150, Dec.py:86-93)")
        value_size = value_size or key_size
        self.mode = mode
        self.query_size, self.key_size, self.value_size = query_size, key_size, value_size
        self.batch_first = batch_first
        self.mask = None
        self.hidden_size = hidden_size
        self.hard_att = hard_att
        self.linear_att_w = nn.Linear(self.key_size, self.query_size, bias=False)   # attention.py:68
        self.output_size = value_size
        self.dropout = nn.Dropout(dropout)
        self.output_nonlinearity = output_nonlinearity

    def set_mask(self, mask):
        """mask: bool [b, t_k], True over padded keys (attention.py:82-89)."""
        self.mask = mask

    def forward(self, query, keys, values=None, prev_c=None, use_gpu=True):
        single = query.dim() == 2
        if single:
            query = query.unsqueeze(1)
        values = keys if values is None else values
        klens = None
        if self.mask is not None:
            # the decoder's mask is always "j >= len" (Dec.py:173-181): recover the length per row
            klens = (~self.mask.bool()).sum(dim=1).to(torch.int32)
        ctx, probs = _BilinearAttention.apply(query, keys, values, self.linear_att_w.weight, klens)
        if single:
            ctx, probs = ctx.squeeze(1), probs.squeeze(1)
        return ctx, probs, None
