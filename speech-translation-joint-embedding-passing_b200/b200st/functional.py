"""Autograd functions that orchestrate the b200st kernels (forward and hand-written backward).

Everything numerical happens inside libb200st.so; this file only sequences launches, owns the saved
buffers and tells autograd which gradient belongs to which input.  Citations are to the reference files
whose PyTorch library calls each function replaces.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd import Function

from .kernels import K
from . import runtime as rt

PAD, UNK, BOS, EOS, SPC = 0, 1, 2, 3, 4   # utils/config.py:7


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Gradients arrive from autograd with arbitrary strides; kernels want dense rows."""
    return None if t is None else (t if t.is_contiguous() else t.contiguous())


# ------------------------------------------------------------------------------------------------
# Linear (+bias, +ReLU, +residual): nn.Linear call sites, layers.py:131-134,158-160,192-195,238-252
# ------------------------------------------------------------------------------------------------
class _Linear(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu, residual):
        assert not (relu and residual is not None)
        k = weight.size(1)
        x2 = x.reshape(-1, k)
        w = rt.operand(weight)
        r2 = None if residual is None else residual.reshape(-1, weight.size(0))
        y = K().gemm(x2, w, trans_b=True, bias=bias, relu=relu, residual=r2)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x2, weight, y if relu else None)
        return y.view(*x.shape[:-1], weight.size(0))

    @staticmethod
    def backward(ctx, dy):
        x2, weight, y = ctx.saved_tensors
        n = weight.size(0)
        dy2 = _c(dy).reshape(-1, n)
        dz = K().relu_bwd(dy2, y) if ctx.relu else dy2
        w = rt.operand(weight)
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = K().gemm(dz, w).view(*dy.shape[:-1], weight.size(1))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = K().colsum(dz)
        if ctx.needs_input_grad[1]:
            side = rt.side_streams(dz.device, 1, pool='dw') if rt.can_defer(weight) else [None]
            with rt.fork(side[0]):      # weight gradient: off the dX chain, joined at the end of backward
                dw = K().gemm(dz, x2, trans_a=True, out_dtype=torch.float32)
            rt.defer(side[0], (dz, x2), [(weight, dw)])
        if ctx.has_res and ctx.needs_input_grad[4]:
            dres = dy
        return dx, dw, db, None, dres


def linear(x, weight, bias=None, relu=False, residual=None):
    return _Linear.apply(x, weight, bias, relu, residual)


# ------------------------------------------------------------------------------------------------
# nn.Dropout call sites (Enc.py:159-212, Seq2seq.py:195-209, layers.py:182-249).  The mask is never stored: it is a
# function of (step seed, site, element index) that the backward launch recomputes (csrc/philox.cuh).
# ------------------------------------------------------------------------------------------------
class _Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, site, residual):
        rng = rt.current_rng(x.device)
        ctx.p, ctx.site, ctx.has_res = p, site, residual is not None
        ctx.save_for_backward(rng)
        return K().dropout(_c(x), p, rng, site, residual=None if residual is None else _c(residual))

    @staticmethod
    def backward(ctx, dy):
        (rng,) = ctx.saved_tensors
        dy = _c(dy)
        return K().dropout(dy, ctx.p, rng, ctx.site), None, None, (dy if ctx.has_res else None)


def dropout(x, p, training=True, tag='', residual=None):
    """dropout(x) (+ residual).  Identity (or a plain add) when p == 0 or not training."""
    if not training or p <= 0:
        return x if residual is None else _Add.apply(x, residual)
    return _Dropout.apply(x, p, rt.next_site(tag, p), residual)


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return K().add(_c(a), _c(b))

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


# ------------------------------------------------------------------------------------------------
# LayerNorm: layers.py:139,153,240,245; TFEnc.py:61,89; TFDec.py:58,127
# ------------------------------------------------------------------------------------------------
def _ln_backward(k, dy, x, ln_w, ln_b, mean, rstd, add=None):
    """LayerNorm backward -> (dx, dgamma, dbeta).  Inside a graph capture the per-CTA dgamma / dbeta partial sums are
    stored instead of reduced with same-address atomics, and their column sums run on a side stream whose join is
    deferred to the end of backward (they are parameter gradients: nothing in backward reads them)."""
    D = x.size(-1)
    side = rt.side_streams(x.device, 1, pool='dw') if rt.can_defer(ln_w, ln_b) else [None]
    if side[0] is not None:
        dx, part = k.layernorm_bwd_partial(dy, x, ln_w, mean, rstd, add=add)
        if part is not None:
            with rt.fork(side[0]):
                dgamma = k.colsum(part[:, :D])
                dbeta = k.colsum(part[:, D:])
            rt.defer(side[0], (part,), [(ln_w, dgamma), (ln_b, dbeta)])
            return dx, dgamma, dbeta
    dln = torch.zeros((2, D), dtype=torch.float32, device=x.device)
    dx = k.layernorm_bwd(dy, x, ln_w, mean, rstd, dln[0], dln[1], add=add)
    return dx, dln[0], dln[1]


def _dx_ln_backward(k, dyp, w, x, ln_w, ln_b, mean, rstd, add=None):
    """LayerNorm backward of  dy = dyp @ w  (the input-gradient GEMM in front of a pre-norm) -> (dx [+ add], dgamma, dbeta).
    One launch (b200st_gemm_lnbwd: the LayerNorm backward is the GEMM's epilogue) where the shapes allow, the GEMM and the
    LayerNorm-backward kernel otherwise.  The dgamma / dbeta column sums of the per-row-block partials run on a side stream
    whose join is deferred to the end of backward when the gradients can be adopted (graph capture)."""
    if not k.gemm_lnbwd_ok(dyp, w, x, add):
        return _ln_backward(k, k.gemm(dyp, w), x, ln_w, ln_b, mean, rstd, add=add)
    D = x.size(-1)
    dx, part = k.gemm_lnbwd(dyp, w, x, ln_w, mean, rstd, add=add)
    side = rt.side_streams(x.device, 1, pool='dw') if rt.can_defer(ln_w, ln_b) else [None]
    with rt.fork(side[0]):
        dgamma = k.colsum(part[:, :D])
        dbeta = k.colsum(part[:, D:])
    rt.defer(side[0], (part,), [(ln_w, dgamma), (ln_b, dbeta)])
    return dx, dgamma, dbeta


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        xc = _c(x)
        pre = rt.take_prenorm(xc, weight, bias, eps)          # the producing GEMM's epilogue may already have normalised x
        y, mean, rstd = pre if pre is not None else K().layernorm_fwd(xc, weight, bias, eps)
        ctx.save_for_backward(xc, weight, bias, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        dx, dgamma, dbeta = _ln_backward(K(), _c(dy), x, weight, bias, mean, rstd)
        return dx, dgamma, dbeta, None


def layer_norm(x, weight, bias, eps):
    return _LayerNorm.apply(x, weight, bias, eps)


# ------------------------------------------------------------------------------------------------
# Scaled-dot-product attention core: layers.py:162-170,213-229
# ------------------------------------------------------------------------------------------------
class _MHACore(Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, n_head, temperature, p_attn, site):
        rng = rt.current_rng(q.device) if p_attn > 0 else None
        o, p = K().mha_fwd(q, k, v, mask, n_head, temperature, dropout=(p_attn, rng, site))
        ctx.n_head, ctx.temperature, ctx.drop = n_head, temperature, (p_attn, site)
        ctx.save_for_backward(q, k, v, p, rng)
        ctx.mark_non_differentiable(p)
        return o, p

    @staticmethod
    def backward(ctx, do, _dp):
        q, k, v, p, rng = ctx.saved_tensors
        dq, dk, dv = K().mha_bwd(_c(do), q, k, v, p, ctx.n_head, ctx.temperature,
                                 dropout=(ctx.drop[0], rng, ctx.drop[1]))
        return dq, dk, dv, None, None, None, None, None


def mha_core(q, k, v, mask, n_head, temperature, p_attn=0.0, tag=''):
    """q [B,Lq,H*d], k/v [B,Lk,H*d] (dense last dim), mask uint8/bool [B,1|Lq,Lk] or None.  p_attn > 0: attention
    dropout on the probabilities (layers.py:226); the returned probabilities are the un-dropped ones."""
    site = rt.next_site(tag + '.attn', p_attn) if p_attn > 0 else 0
    return _MHACore.apply(q, k, v, mask, n_head, temperature, p_attn, site)


# ------------------------------------------------------------------------------------------------
# Whole residual sub-layers as ONE autograd node each.  Same kernels as the pieces above; what the fusion buys is
# launch count on a path that is latency-bound (12 layers x ~30 small kernels): K and V come out of one GEMM on the
# concatenated w_ks|w_vs, their input gradient out of one GEMM, the skip-connection gradient is added inside the
# LayerNorm-backward kernel and (self-attention) inside the K|V input-gradient GEMM's epilogue, and the two
# LayerNorm parameter gradients share one zero-filled buffer.
# ------------------------------------------------------------------------------------------------
class _MHABlock(Function):
    """out = fc(attention(w_qs(LN(q)), w_ks(kv), w_vs(kv))) + q;  MultiHeadAttention.forward, layers.py:148-197
    (LayerNorm on the query input only, K/V from the raw input, no biases, dropout p = 0)."""

    @staticmethod
    def forward(ctx, q, kv, mask, ln_w, ln_b, eps, w_q, w_k, w_v, w_fc, n_head, temperature, drop, next_ln=None):
        # drop = (p_attn, site_attn, p_fc, site_fc): attention dropout on the probabilities (layers.py:226) and
        # dropout on the fc output before the residual add (layers.py:194-195); all zeros = the fused fast path
        k = K()
        p_attn, s_attn, p_fc, s_fc = drop
        rng = rt.current_rng(q.device) if (p_attn > 0 or p_fc > 0) else None
        B, Lq, D = q.shape
        Lk = kv.size(1)
        self_attn = kv is q
        q2 = _c(q).reshape(-1, D)
        kv2 = q2 if self_attn else _c(kv).reshape(-1, D)
        HD = w_q.size(0)
        side = rt.side_streams(q.device, 1, pool='fwd')
        with rt.fork(side[0]):          # K | V come from the RAW input: independent of the LayerNorm -> Q chain
            kvp = k.gemm(kv2, rt.operand_cat(w_k, w_v), trans_b=True)             # [B*Lk, 2*HD] = K | V
        pre = rt.take_prenorm(q2, ln_w, ln_b, eps)           # LayerNorm(q) from the previous sub-layer's GEMM epilogue
        qn, mean, rstd = pre if pre is not None else k.layernorm_fwd(q2, ln_w, ln_b, eps)
        qp = k.gemm(qn, rt.operand(w_q), trans_b=True)
        rt.join(side[0])
        kv3 = kvp.view(B, Lk, 2 * HD)
        o, p = k.mha_fwd(qp.view(B, Lq, HD), kv3[:, :, :HD], kv3[:, :, HD:], mask, n_head, temperature,
                         dropout=(p_attn, rng, s_attn))
        o2 = o.view(-1, HD)
        wfc = rt.operand(w_fc)
        if next_ln is not None and k.gemm_ln_ok(o2, wfc, q2):
            # fc (+ dropout) + skip connection + the NEXT sub-layer's LayerNorm in one launch; handed over through rt
            out, yn, mean_n, rstd_n = k.gemm_ln(o2, wfc, None, q2, next_ln[0], next_ln[1], next_ln[2],
                                                dropout=(p_fc, rng, s_fc) if p_fc > 0 else None)
            rt.offer_prenorm(out, next_ln[0], next_ln[1], next_ln[2], yn, mean_n, rstd_n)
        elif p_fc > 0:
            out = k.dropout(k.gemm(o2, wfc, trans_b=True), p_fc, rng, s_fc, residual=q2)
        else:
            out = k.gemm(o2, wfc, trans_b=True, residual=q2)
        ctx.self_attn, ctx.n_head, ctx.temperature, ctx.dims = self_attn, n_head, temperature, (B, Lq, Lk, D, HD)
        ctx.kv_shape, ctx.drop = kv.shape, drop
        ctx.save_for_backward(q2, None if self_attn else kv2, qn, mean, rstd, qp, kvp, p, o2, ln_w, w_q, w_k, w_v, w_fc,
                              rng, ln_b)
        ctx.mark_non_differentiable(p)
        return out.view(B, Lq, D), p

    @staticmethod
    def backward(ctx, dout, _dp):
        k = K()
        q2, kv2, qn, mean, rstd, qp, kvp, p, o2, ln_w, w_q, w_k, w_v, w_fc, rng, ln_b = ctx.saved_tensors
        B, Lq, Lk, D, HD = ctx.dims
        p_attn, s_attn, p_fc, s_fc = ctx.drop
        if kv2 is None:
            kv2 = q2
        dout2 = _c(dout).reshape(-1, D)
        dfc = k.dropout(dout2, p_fc, rng, s_fc) if p_fc > 0 else dout2      # gradient of the fc output
        do = k.gemm(dfc, rt.operand(w_fc))
        # weight gradients: off the dX chain on side streams, joined at the end of backward (rt.defer).  Every returned
        # gradient is a fresh dense tensor (or a row block of one) so that autograd adopts it instead of copying it
        side = rt.side_streams(q2.device, 2, pool='dw') if rt.can_defer(w_fc, w_q, w_k, w_v) else [None, None]
        with rt.fork(side[0]):
            dw_fc = k.gemm(dfc, o2, trans_a=True, out_dtype=torch.float32)
        rt.defer(side[0], (dfc, o2), [(w_fc, dw_fc)])
        dqp = torch.empty_like(qp)
        dkvp = torch.empty_like(kvp)
        kv3, dkv3 = kvp.view(B, Lk, 2 * HD), dkvp.view(B, Lk, 2 * HD)
        k.mha_bwd(do.view(B, Lq, HD), qp.view(B, Lq, HD), kv3[:, :, :HD], kv3[:, :, HD:], p, ctx.n_head,
                  ctx.temperature, dq=dqp.view(B, Lq, HD), dk=dkv3[:, :, :HD], dv=dkv3[:, :, HD:],
                  dropout=(p_attn, rng, s_attn))
        with rt.fork(side[1]):
            dw_q = k.gemm(dqp, qn, trans_a=True, out_dtype=torch.float32)
            # K | V weight gradients from ONE GEMM on the concatenated gate gradients: the two halves are contiguous ROW
            # blocks of its output, which autograd adopts as they are (a row-slice view is dense; column slices are not)
            dw_kv = k.gemm(dkvp, kv2, trans_a=True, out_dtype=torch.float32)
            dw_k, dw_v = dw_kv[:HD], dw_kv[HD:]
        rt.defer(side[1], (dqp, qn, dkvp, kv2), [(w_q, dw_q), (w_k, dw_k), (w_v, dw_v)])
        # dX of the K | V projections runs on a parallel branch beside the dQ GEMM; for self-attention its epilogue also
        # adds the skip-connection gradient, and the LayerNorm-backward kernel adds that sum to its own dx: the chain is
        # fc-dX -> attention backward -> max(dQ, dK|V GEMM) -> LayerNorm backward (4 dependent kernels, not 5)
        wkv = rt.operand_cat(w_k, w_v)
        side_x = rt.side_streams(q2.device, 1, pool='fwd')
        with rt.fork(side_x[0]):
            dkv = k.gemm(dkvp, wkv, residual=dout2 if ctx.self_attn else None)
        rt.join(side_x[0])
        # dQ GEMM with the LayerNorm backward (+ the skip-connection / K|V gradient) as its epilogue
        if ctx.self_attn:
            dq, dgamma, dbeta = _dx_ln_backward(k, dqp, rt.operand(w_q), q2, ln_w, ln_b, mean, rstd, add=dkv)
            dkv = None
        else:
            dq, dgamma, dbeta = _dx_ln_backward(k, dqp, rt.operand(w_q), q2, ln_w, ln_b, mean, rstd, add=dout2)
            dkv = dkv.view(ctx.kv_shape)
        return (dq.view(B, Lq, D), dkv, None, dgamma, dbeta, None, dw_q, dw_k, dw_v, dw_fc, None, None,
                None, None)


def mha_block(q, kv, mask, ln_w, ln_b, eps, w_q, w_k, w_v, w_fc, n_head, temperature, p_attn=0.0, p_fc=0.0, tag='',
              next_ln=None):
    """Returns (out [B, Lq, D], attention probabilities [B, H, Lq, Lk]).  Pass the SAME tensor object as q and kv
    for self-attention: its two gradient contributions are then merged inside the GEMM epilogue.
    p_attn / p_fc: dropout on the attention probabilities / on the fc output (training mode values, 0 otherwise).
    next_ln = (weight, bias, eps) of the LayerNorm that will normalise the OUTPUT of this sub-layer next (the following
    sub-layer's pre-norm): when given (and dropout is off, bf16, d_model = 512) it is computed in the fc GEMM's epilogue."""
    drop = (p_attn, rt.next_site(tag + '.attn', p_attn) if p_attn > 0 else 0,
            p_fc, rt.next_site(tag + '.fc', p_fc) if p_fc > 0 else 0)
    return _MHABlock.apply(q, kv, mask, ln_w, ln_b, eps, w_q, w_k, w_v, w_fc, n_head, temperature, drop, next_ln)


class _FFNBlock(Function):
    """x + w_2(relu(w_1(LN(x))));  PositionwiseFeedForward.forward, layers.py:232-252."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, w1, b1, w2, b2, p, site, next_ln=None):
        k = K()
        D = x.size(-1)
        x2 = _c(x).reshape(-1, D)
        pre = rt.take_prenorm(x2, ln_w, ln_b, eps)           # LayerNorm(x) from the previous sub-layer's GEMM epilogue
        y, mean, rstd = pre if pre is not None else k.layernorm_fwd(x2, ln_w, ln_b, eps)
        h = k.gemm(y, rt.operand(w1), trans_b=True, bias=b1, relu=True)
        rng = None
        if p > 0:            # x + dropout(w_2(.)) (layers.py:248-250)
            rng = rt.current_rng(x.device)
            if next_ln is not None and k.gemm_ln_ok(h, rt.operand(w2), x2):
                out, yn, mean_n, rstd_n = k.gemm_ln(h, rt.operand(w2), b2, x2, next_ln[0], next_ln[1], next_ln[2],
                                                    dropout=(p, rng, site))
                rt.offer_prenorm(out, next_ln[0], next_ln[1], next_ln[2], yn, mean_n, rstd_n)
            else:
                out = k.dropout(k.gemm(h, rt.operand(w2), trans_b=True, bias=b2), p, rng, site, residual=x2)
        elif next_ln is not None and k.gemm_ln_ok(h, rt.operand(w2), x2):
            out, yn, mean_n, rstd_n = k.gemm_ln(h, rt.operand(w2), b2, x2, next_ln[0], next_ln[1], next_ln[2])
            rt.offer_prenorm(out, next_ln[0], next_ln[1], next_ln[2], yn, mean_n, rstd_n)
        else:
            out = k.gemm(h, rt.operand(w2), trans_b=True, bias=b2, residual=x2)
        ctx.drop = (p, site)
        ctx.save_for_backward(x2, y, mean, rstd, h, ln_w, w1, w2, rng, ln_b, b1, b2)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dout):
        k = K()
        x2, y, mean, rstd, h, ln_w, w1, w2, rng, ln_b, b1, b2 = ctx.saved_tensors
        D = x2.size(1)
        dout2 = _c(dout).reshape(-1, D)
        d2 = k.dropout(dout2, ctx.drop[0], rng, ctx.drop[1]) if ctx.drop[0] > 0 else dout2   # gradient of w_2's output
        dz = k.gemm(d2, rt.operand(w2), relu_gate=h)                              # ReLU backward in the epilogue
        side = rt.side_streams(x2.device, 2, pool='dw') if rt.can_defer(w1, w2, b1, b2) else [None, None]
        with rt.fork(side[0]):              # weight / bias gradients: off the dX chain, joined at the end of backward
            dw2 = k.gemm(d2, h, trans_a=True, out_dtype=torch.float32)
            db2 = k.colsum(d2)
        rt.defer(side[0], (d2, h), [(w2, dw2), (b2, db2)])
        with rt.fork(side[1]):
            dw1 = k.gemm(dz, y, trans_a=True, out_dtype=torch.float32)
            db1 = k.colsum(dz)
        rt.defer(side[1], (dz, y), [(w1, dw1), (b1, db1)])
        dx, dgamma, dbeta = _dx_ln_backward(k, dz, rt.operand(w1), x2, ln_w, ln_b, mean, rstd, add=dout2)
        return dx.view(dout.shape), dgamma, dbeta, None, dw1, db1, dw2, db2, None, None, None


def ffn_block(x, ln_w, ln_b, eps, w1, b1, w2, b2, p=0.0, tag='', next_ln=None):
    """next_ln: see mha_block."""
    return _FFNBlock.apply(x, ln_w, ln_b, eps, w1, b1, w2, b2, p, rt.next_site(tag + '.ffn', p) if p > 0 else 0, next_ln)


# ------------------------------------------------------------------------------------------------
# Embedding lookups: Seq2seq.py:188,207; Dec.py:166,223
# ------------------------------------------------------------------------------------------------
class _Embedding(Function):
    @staticmethod
    def forward(ctx, ids, table, padding_idx):
        ids = _c(ids)
        out = K().embedding_fwd(ids, table, rt.compute_dtype())
        ctx.padding_idx = padding_idx
        ctx.save_for_backward(ids, table)
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, table = ctx.saved_tensors
        dtable = torch.zeros_like(table)
        K().embedding_bwd(ids.reshape(-1), _c(dout).reshape(-1, table.size(1)), dtable, ctx.padding_idx)
        return None, dtable, None


def embedding(ids, table, padding_idx=PAD):
    return _Embedding.apply(ids, table, padding_idx)


class _AddPosEnc(Function):
    @staticmethod
    def forward(ctx, x, pe):
        return K().add_posenc(_c(x), pe)

    @staticmethod
    def backward(ctx, dy):
        return dy, None


def add_posenc(x, pe):
    """x + time_signal[:, :L] (TFEnc.py:82-83, TFDec.py:85-86); pe: fp32 [max_len, D] on the device."""
    return _AddPosEnc.apply(x, pe)


# ------------------------------------------------------------------------------------------------
# The embedding-passing mix: Seq2seq._get_src_emb, Seq2seq.py:183-199
#   emb_src = enc_emb_proj(cat(E_static[src], e_dyn)),  Linear(E + D -> D, no bias)
# ------------------------------------------------------------------------------------------------
class _Mix(Function):
    @staticmethod
    def forward(ctx, ids, table, dyn, weight, p, site):
        b, s = ids.shape
        ids = _c(ids)
        dyn2 = _c(dyn).reshape(b * s, -1)
        cat = K().mix_gather_concat(ids.reshape(-1), table, dyn2)
        rng = None
        if p > 0:            # embedding_dropout on the concatenation (Seq2seq.py:195)
            rng = rt.current_rng(ids.device)
            K().dropout(cat, p, rng, site, out=cat)
        out = K().gemm(cat, rt.operand(weight), trans_b=True)
        ctx.drop = (p, site)
        ctx.save_for_backward(ids, table, cat, weight, rng)
        return out.view(b, s, weight.size(0))

    @staticmethod
    def backward(ctx, dy):
        ids, table, cat, weight, rng = ctx.saved_tensors
        e = table.size(1)
        d_out = weight.size(0)
        dy2 = _c(dy).reshape(-1, d_out)
        w = rt.operand(weight)
        p, site = ctx.drop
        ld = weight.size(1)                                       # E + D: the mask is indexed over the whole concat row
        dtable = ddyn = dw = None
        if ctx.needs_input_grad[1]:
            dstatic = K().gemm(dy2, w[:, :e])                     # [n, E]
            if p > 0:
                K().dropout(dstatic, p, rng, site, out=dstatic, ld_mask=ld, col_off=0)
            dtable = torch.zeros_like(table)
            K().embedding_bwd(ids.reshape(-1), dstatic, dtable, PAD)   # padding_idx=PAD, Seq2seq.py:106-107
        if ctx.needs_input_grad[2]:
            ddyn = K().gemm(dy2, w[:, e:])
            if p > 0:
                K().dropout(ddyn, p, rng, site, out=ddyn, ld_mask=ld, col_off=e)
            ddyn = ddyn.view(ids.size(0), ids.size(1), -1)
        if ctx.needs_input_grad[3]:
            dw = K().gemm(dy2, cat, trans_a=True, out_dtype=torch.float32)       # cat holds the dropped values
        return None, dtable, ddyn, dw, None, None


def mix(ids, table, dyn, weight, p=0.0, tag='mix'):
    return _Mix.apply(ids, table, dyn, weight, p, rt.next_site(tag, p) if p > 0 else 0)


# ------------------------------------------------------------------------------------------------
# log-softmax (+arg-max) and the masked NLL: Seq2seq.py:254-255; loss.py:116-132
# ------------------------------------------------------------------------------------------------
class _LogSoftmax(Function):
    @staticmethod
    def forward(ctx, x):
        x2 = _c(x).reshape(-1, x.size(-1))
        y, am = K().log_softmax_fwd(x2, want_argmax=True)
        ctx.save_for_backward(y)
        am = am.view(*x.shape[:-1], 1)
        ctx.mark_non_differentiable(am)
        return y.view(x.shape), am

    @staticmethod
    def backward(ctx, dy, _dam):
        (y,) = ctx.saved_tensors
        dx = K().log_softmax_bwd(_c(dy).reshape(y.shape), y)
        return dx.view(dy.shape)


def log_softmax_argmax(x):
    """Returns (log_softmax(x, -1), argmax(x, -1, keepdim=True)) in one pass over the vocabulary."""
    return _LogSoftmax.apply(x)


class _MaskedNLL(Function):
    @staticmethod
    def forward(ctx, logp, target, mask):
        lp = logp if (logp.dim() == 2 and logp.stride(1) == 1) else _c(logp)
        mask = None if mask is None else _c(mask)
        target = _c(target)
        loss = K().masked_nll_fwd(lp, target, mask)
        ctx.save_for_backward(target, mask)
        ctx.shape, ctx.dtype = tuple(logp.shape), logp.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        target, mask = ctx.saved_tensors
        g32 = _c(g).reshape(1).to(torch.float32)
        d = K().masked_nll_bwd(g32, target, mask, ctx.shape[0], ctx.shape[1], ctx.dtype)
        return d, None, None


def masked_nll_sum(logp, target, mask):
    """sum over rows with mask!=0 of -logp[r, target[r]] (loss.py:130-132).  Scalar fp32."""
    return _MaskedNLL.apply(logp, target, mask)


_unit_upstream = [False]     # did the last eager backward of the fused loss see an upstream gradient of exactly 1?


class _FusedSoftmaxNLL(Function):
    """K17: softmax + masked NLL + gradient in one pass; gradient is formed in forward."""

    @staticmethod
    def forward(ctx, logits, target, mask, scale, eps):
        x2 = logits.reshape(-1, logits.size(-1))
        loss, d = K().softmax_nll_fused(x2, _c(target).reshape(-1), None if mask is None else _c(mask).reshape(-1),
                                        scale, eps)
        ctx.save_for_backward(d)
        ctx.shape = logits.shape
        return (loss * scale).reshape(())

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        # The upstream gradient g is 1 in the training step (loss.backward() on the already normalised loss); anything
        # else is a scalar rescale of the stored gradient.  Multiplying regardless costs a full pass over [rows, V]
        # (128 MB at configs[2], ~60 us on the critical path), and g is a device value: an eager pass reads it (one
        # sync, eager only) and remembers whether it was exactly 1; a graph capture -- which replays the same Python
        # path its eager warm-up ran -- skips the multiply only if that check passed.
        dl = d.view(ctx.shape)
        if torch.cuda.is_available() and g.is_cuda and torch.cuda.is_current_stream_capturing():
            unit = _unit_upstream[0]
            if unit:          # keep the captured value of g: whoever replays the graph verifies it after the first replay
                rt.unit_grad_probes.append(g.detach().reshape(-1)[:1].clone())
        else:
            unit = bool((g == 1).all())
            _unit_upstream[0] = unit
        if not unit:
            dl = dl * g.to(dl.dtype)
        return dl, None, None, None, None


def fused_softmax_nll(logits, target, mask, scale, eps=0.0):
    """mean-style loss `scale * sum_mask(lse - logit[target])` with dlogits produced in the same kernel.
    `scale` is a 1-element fp32 device tensor (e.g. 1 / #non-PAD)."""
    return _FusedSoftmaxNLL.apply(logits, target, mask, scale, eps)


# ------------------------------------------------------------------------------------------------
# One bidirectional packed LSTM layer of the pyramidal encoder: Enc.py:150-167 (x4)
# ------------------------------------------------------------------------------------------------
import os as _os
# SM budget of the PERSISTENT weight-gradient GEMMs that are deferred onto side streams and end up running under a BLSTM
# backward recurrence (the upper BLSTM layers' own, and the LAS decoder's): measured inside the replayed step graph
# (scripts/graph_timeline.py, profiles/r02_graph_timeline.txt) those GEMMs, streaming 130-260 MB each at full speed, slowed
# every backward recurrence by ~25 % (1444 instead of 1160 us for the bottom layer) -- the recurrence's per-step loads of the
# saved state are prefetched two steps ahead and a saturated memory system pushes their latency past that.  On <= 24 SMs the
# GEMMs still finish inside the recurrence they run under and the recurrence keeps its stand-alone speed: -0.4 ms per step.
DEFERRED_SM_BUDGET = int(_os.environ.get('B200ST_DEFERRED_SM_BUDGET', '16'))
LAS_DEFERRED_SM_BUDGET = int(_os.environ.get('B200ST_LAS_DEFERRED_SM_BUDGET', '16'))


class _BLSTMLayer(Function):
    @staticmethod
    def forward(ctx, x, lens, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, pair,
                batch_first_out):
        # x: [T, B, I] time-major, compute dtype; lens int32 [B] valid frames at this layer.
        k = K()
        T, B, I = x.shape
        H = w_hh_f.size(1)
        x2 = x.reshape(T * B, I)
        xproj = torch.empty((2, T, B, 4 * H), dtype=x.dtype, device=x.device)
        k.gemm(x2, rt.operand(w_ih_f), trans_b=True, bias=k.add(b_ih_f, b_hh_f), out=xproj[0].view(T * B, 4 * H))
        k.gemm(x2, rt.operand(w_ih_r), trans_b=True, bias=k.add(b_ih_r, b_hh_r), out=xproj[1].view(T * B, 4 * H))
        if batch_first_out:
            assert pair == 1
            out = torch.empty((B, T, 2 * H), dtype=x.dtype, device=x.device)
            ld_t, ld_b = 2 * H, T * 2 * H
        else:
            assert T % pair == 0
            out = torch.empty((T // pair, B, pair * 2 * H), dtype=x.dtype, device=x.device)
            ld_t, ld_b = B * pair * 2 * H, pair * 2 * H
        need_grad = any(ctx.needs_input_grad)
        hs, acts, cs = k.blstm_fwd(xproj, w_hh_f, w_hh_r, lens, out, ld_t, ld_b, pair, save=need_grad)
        ctx.geom = (T, B, I, H, pair, ld_t, ld_b)
        ctx.save_for_backward(x2, lens, hs, acts, cs, w_ih_f, w_hh_f, w_ih_r, w_hh_r, b_ih_f, b_hh_f, b_ih_r, b_hh_r)
        return out

    @staticmethod
    def backward(ctx, dout):
        k = K()
        x2, lens, hs, acts, cs, w_ih_f, w_hh_f, w_ih_r, w_hh_r, b_ih_f, b_hh_f, b_ih_r, b_hh_r = ctx.saved_tensors
        T, B, I, H, pair, ld_t, ld_b = ctx.geom
        dg = k.blstm_bwd(_c(dout), ld_t, ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens, x2.dtype)
        dgf, dgr = dg[0].view(T * B, 4 * H), dg[1].view(T * B, 4 * H)
        f32 = torch.float32
        dev = dgf.device
        dw_ih_f, dw_ih_r = torch.empty_like(w_ih_f), torch.empty_like(w_ih_r)
        dw_hh_f, dw_hh_r = torch.empty_like(w_hh_f), torch.empty_like(w_hh_r)
        db_f = torch.empty(4 * H, dtype=f32, device=dev)
        db_r = torch.empty(4 * H, dtype=f32, device=dev)
        side = (rt.side_streams(dev, 2, pool='dw')
                if rt.can_defer(w_ih_f, w_hh_f, w_ih_r, w_hh_r, b_ih_f, b_hh_f, b_ih_r, b_hh_r) else [None, None])
        # Weight AND bias gradients leave the dependent chain (nothing in backward reads them): split-K GEMMs and column sums
        # over T*B rows on the side streams, under the next layer's recurrence.  b_ih and b_hh get the same values -- ONE pass
        # over the gate gradients (132 MB at the bottom layer) and a 4 KB copy, because autograd adopts a returned gradient only
        # if nothing else references it.
        # A layer above the bottom one is followed by the next recurrence kernel: its deferred persistent GEMMs are launched
        # with a small SM budget (see DEFERRED_SM_BUDGET) so that they neither hold the SMs the recurrence's clusters need nor
        # saturate the memory system under it.
        budget = DEFERRED_SM_BUDGET if (side[0] is not None and ctx.needs_input_grad[0]) else 0
        old_budget = k.set_gemm_sm_budget(budget)
        try:
            with rt.fork(side[0]):
                k.gemm(dgf, hs[0, :T].reshape(T * B, H), trans_a=True, out=dw_hh_f)
                k.gemm(dgr, hs[1, 1:].reshape(T * B, H), trans_a=True, out=dw_hh_r)
                k.colsum(dgf, out=db_f)
                db_f2 = k.cast(db_f, f32, out=torch.empty_like(db_f))
            with rt.fork(side[1]):
                k.gemm(dgf, x2, trans_a=True, out=dw_ih_f)
                k.gemm(dgr, x2, trans_a=True, out=dw_ih_r)
                k.colsum(dgr, out=db_r)
                db_r2 = k.cast(db_r, f32, out=torch.empty_like(db_r))
        finally:
            k.set_gemm_sm_budget(old_budget)
        # no join here: the gradients keep running under the NEXT layer's recurrence (rt.defer)
        rt.defer(side[0], (dg, hs), [(w_hh_f, dw_hh_f), (w_hh_r, dw_hh_r), (b_ih_f, db_f), (b_hh_f, db_f2)])
        rt.defer(side[1], (dg, x2), [(w_ih_f, dw_ih_f), (w_ih_r, dw_ih_r), (b_ih_r, db_r), (b_hh_r, db_r2)])
        dx = None
        if ctx.needs_input_grad[0]:
            # both directions' contributions in ONE launch (two-segment K loop into the same accumulator)
            dx = k.gemm2(dgf, rt.operand(w_ih_f), dgr, rt.operand(w_ih_r)).view(T, B, I)
        return dx, None, dw_ih_f, dw_hh_f, db_f, db_f2, dw_ih_r, dw_hh_r, db_r, db_r2, None, None


def blstm_layer(x_tm, lens, weights_f, weights_r, pair, batch_first_out=False):
    """weights_*: (w_ih, w_hh, b_ih, b_hh) nn.Parameters of one direction of torch.nn.LSTM."""
    return _BLSTMLayer.apply(x_tm, lens, *weights_f, *weights_r, pair, batch_first_out)


# ------------------------------------------------------------------------------------------------
# LAS attention-LSTM decoder loop: Dec.forward / forward_step / decode, Dec.py:130-233,320-438
# ------------------------------------------------------------------------------------------------
class _LASDecoder(Function):
    """One autograd node for the whole S-step loop; backward is hand-written BPTT.

    inputs : enc [B,Tk,2H] (keys = values), klens int32[B] | None, ids_tf int64[B,S+1] | None (teacher
             forcing tokens; None = free running from BOS), emb_table, w_att, w_ffn, w_out, b_out,
             then (w_ih, w_hh, b_ih, b_hh) for each of the `n_layers` uni-LSTMs.
    outputs: embs [B,S,D] (the dynamic embedding = pre-softmax `cell_value`, Dec.py:433),
             logps [B,S,V] | empty, symbols int64 [B,S], lengths int32 [B].
    """

    @staticmethod
    def forward(ctx, enc, klens, ids_tf, n_steps, need_logps, drop, emb_table, w_att, w_ffn, w_out, b_out,
                *lstm_params):
        # drop = (p_emb, p): embedding dropout on the input tokens' embeddings (Dec.py:166 -- every teacher-forcing
        # token, only the BOS embedding when free running, Dec.py:199,223) and dropout on every LSTM layer's output and
        # on the attention context (Dec.py:403,419,429); the recurrent h/c and the residual sum itself stay un-dropped.
        k = K()
        p_emb, p_drop = drop
        dt = enc.dtype
        dev = enc.device
        enc = _c(enc)
        B, Tk, H2 = enc.shape
        S = n_steps
        E = emb_table.size(1)
        D = w_ffn.size(0)
        V = w_out.size(0)
        n_layers = len(lstm_params) // 4
        lp = [lstm_params[4 * i:4 * i + 4] for i in range(n_layers)]
        wih = [rt.operand(p[0]) for p in lp]
        whh = [rt.operand(p[1]) for p in lp]
        bias = [k.add(p[2], p[3]) for p in lp]
        wf = rt.operand(w_ffn)
        wo = rt.operand(w_out)
        # step-invariant bilinear key projection, hoisted out of the loop (attention.py:192)
        wk = k.gemm(enc.view(B * Tk, H2), rt.operand(w_att), trans_b=True).view(B, Tk, D)

        z = lambda *s, dtype=dt: torch.zeros(s, dtype=dtype, device=dev)
        e = lambda *s, dtype=dt: torch.empty(s, dtype=dtype, device=dev)
        if (rt.las_persistent() and dt == torch.bfloat16 and D == 512 and H2 == 512 and n_layers == 3 and p_emb == 0 and
                p_drop == 0 and 1 <= Tk <= 512 and B <= 128 and V <= 128 * 80 and hasattr(k, 'las_decoder_fwd')
                and enc.is_cuda and torch.cuda.get_device_properties(dev).multi_processor_count >= 128):
            return _LASDecoder._forward_persistent(ctx, k, enc, klens, ids_tf, S, need_logps, drop, emb_table, w_att, w_ffn,
                                                   w_out, b_out, lstm_params, lp, wih, whh, bias, wf, wo, wk)
        CV = z(S + 1, B, D)                                   # CV[s+1] = cell_value of step s; CV[0] = 0
        Hst = [z(S + 1, B, D) for _ in range(n_layers)]
        Cst = [z(S + 1, B, D, dtype=torch.float32) for _ in range(n_layers)]
        ACT = [e(S, B, 4 * D, dtype=torch.float32) for _ in range(n_layers)]
        RES = [e(S, B, D) if 0 < i < n_layers - 1 else None for i in range(n_layers)]   # Dec.py:417-418
        CTX = e(S, B, H2)
        PROBS = e(S, B, Tk, dtype=torch.float32)
        EMB = e(S, B, E)
        LOGITS = e(S, B, V)
        SYM = e(S, B, dtype=torch.int64)
        lengths = torch.full((B,), S + 1, dtype=torch.int32, device=dev)       # Dec.py:163
        if ids_tf is None:
            IDS = e(S + 1, B, dtype=torch.int64)
            IDS[0].fill_(BOS)                                                  # Dec.py:158-160,199
            ids_in = IDS[:S]
            sym_dst = IDS[1:]
        else:
            ids_in = ids_tf.t()[:S].contiguous()
            sym_dst = SYM
        # Per-step critical path (5 GEMMs): emb W_a -> cell0 -> x W_ih1 -> cell1 -> x W_ih2 -> cell2 -> attention ->
        # ctx W_fa -> vocabulary -> arg-max.  Everything that only depends on the PREVIOUS step (h_i W_hh_i, cv W_b) or
        # is independent of the attention (dec_out W_fb) is forked onto side streams as soon as its input exists.
        side = rt.side_streams(dev, n_layers + 1)
        Gx = [e(B, 4 * D) for _ in range(n_layers)]          # x-part (+ bias) of the gate pre-activations
        Gh = [e(B, 4 * D) for _ in range(n_layers)]          # h_{s-1} W_hh^T
        Gcv = e(B, 4 * D)                                     # cell_value_{s-1} W_ih0[:, E:]^T
        CVb = e(B, D)                                         # dec_out W_ffn[:, 2H:]^T
        fused_feed = ids_tf is None          # free running: the arg-max kernel also writes the next step's embedding
        if fused_feed:
            k.embedding_fwd(ids_in[0], emb_table, dt, out=EMB[0])
        else:                                # teacher forcing: every input token is known up front
            k.embedding_fwd(ids_in.reshape(-1), emb_table, dt, out=EMB.view(S * B, -1))
        rng = rt.current_rng(dev) if (p_emb > 0 or p_drop > 0) else None
        site_emb = 0
        if p_emb > 0:                        # EMB holds the dropped embeddings from here on
            site_emb = rt.next_site('las.dec.emb', p_emb)
            tgt_e = EMB[0] if fused_feed else EMB.view(S * B, -1)
            k.dropout(tgt_e, p_emb, rng, site_emb, out=tgt_e)
        XD = [e(S, B, D) for _ in range(n_layers)] if p_drop > 0 else None     # dropped layer outputs
        CTXD = e(S, B, H2) if p_drop > 0 else None                               # dropped attention contexts
        site_l = [[0] * n_layers for _ in range(S)]
        site_att = [0] * S
        # Free running: the fed-back token's first-layer gate contribution emb W_ih0[:, :E]^T + b is a row of the table
        # TOK = E W_ih0[:, :E]^T + b ([V, 4D], one GEMM per forward, on a side branch): the arg-max kernel gathers that
        # row for the next step, which takes one GEMM off every step's dependent chain.  (Gated on the vocabulary size:
        # the table costs V / (S B) times the FLOPs of the per-step products it replaces.)
        TOK = None
        if fused_feed and S > 1 and V * 4 * D * 2 <= (256 << 20) and V % 8 == 0:
            with rt.fork(side[0]):      # joined by the first `rt.join(side[0])` of step 1
                TOK = k.gemm(rt.operand(emb_table), wih[0][:, :E], trans_b=True, bias=bias[0])
        for s in range(S):
            x = EMB[s]
            for i in range(n_layers):
                # gates = x W_ih^T + b_ih + b_hh + h W_hh^T   (torch.nn.LSTM step, Dec.py:393-415)
                if i == 0 and TOK is not None and s > 0:
                    pass            # Gx[0] was gathered from TOK by the previous step's arg-max kernel
                elif i == 0:    # x = cat(emb, prev cell_value) (Dec.py:383) without materialising the concat
                    k.gemm(x, wih[0][:, :E], trans_b=True, bias=bias[0], out=Gx[0])
                else:
                    k.gemm(x, wih[i], trans_b=True, bias=bias[i], out=Gx[i])
                if s > 0:
                    rt.join(side[i])
                res_in = x if RES[i] is not None else None
                _, _, _, out_res = k.lstm_cell_fwd(Gx[i], Cst[i][s], residual=res_in, h_out=Hst[i][s + 1],
                                                   c_out=Cst[i][s + 1], acts_out=ACT[i][s],
                                                   res_out=RES[i][s] if RES[i] is not None else None,
                                                   gates_b=Gh[i] if s > 0 else None,
                                                   gates_c=Gcv if (i == 0 and s > 0) else None)
                if s + 1 < S:
                    with rt.fork(side[i]):                    # next step's recurrent part, off the critical path
                        k.gemm(Hst[i][s + 1], whh[i], trans_b=True, out=Gh[i])
                x = out_res if out_res is not None else Hst[i][s + 1]
                if p_drop > 0:
                    site_l[s][i] = rt.next_site(f'las.dec.s{s}.l{i}', p_drop)
                    x = k.dropout(x, p_drop, rng, site_l[s][i], out=XD[i][s])
            dec_out = x
            with rt.fork(side[n_layers]):
                k.gemm(dec_out, wf[:, H2:], trans_b=True, out=CVb)
            k.las_attn_fwd(dec_out, wk, enc, klens, ctx_out=CTX[s], probs_out=PROBS[s])
            # cell_value = acous_ffn(cat(context, dec_out)) (Dec.py:431-433), again without the concat
            rt.join(side[n_layers])
            ctx_s = CTX[s]
            if p_drop > 0:
                site_att[s] = rt.next_site(f'las.dec.s{s}.att', p_drop)
                ctx_s = k.dropout(CTX[s], p_drop, rng, site_att[s], out=CTXD[s])
            k.gemm(ctx_s, wf[:, :H2], trans_b=True, residual=CVb, out=CV[s + 1])
            if s + 1 < S:
                with rt.fork(side[0]):
                    k.gemm(CV[s + 1], wih[0][:, E:], trans_b=True, out=Gcv)
            k.gemm(CV[s + 1], wo, trans_b=True, bias=b_out, out=LOGITS[s])       # Dec.py:434
            feed = (emb_table, EMB[s + 1]) if (fused_feed and s + 1 < S) else None
            feed2 = (TOK, Gx[0]) if (TOK is not None and s + 1 < S) else None
            if s == 0 and feed2 is not None:
                rt.join(side[0])        # the table GEMM was forked before the loop
            k.argmax_rows(LOGITS[s], sym_dst[s], lengths=lengths, step=s, embed=feed, embed2=feed2)   # Dec.py:331 + 334-341
        for st in side:
            rt.join(st)
        if ids_tf is None:
            SYM = IDS[1:]
        embs = k.transpose01(CV[1:])                                             # [B,S,D]
        if need_logps:
            logp_tm, _ = k.log_softmax_fwd(LOGITS.view(S * B, V))
            logps = k.transpose01(logp_tm.view(S, B, V))
        else:
            logp_tm = None
            logps = torch.empty(0, dtype=dt, device=dev)
        symbols = SYM.t().contiguous()
        ctx.geom = (B, Tk, H2, S, E, D, V, n_layers)
        ctx.need_logps = need_logps
        ctx.has_klens = klens is not None
        ctx.n_fixed = 10
        ctx.drop = (p_emb, p_drop, site_emb, site_l, site_att, fused_feed)
        saved = [enc, klens, ids_in.contiguous(), wk, CV, CTX, PROBS, EMB, logp_tm, emb_table, w_att,
                 w_ffn, w_out]
        saved += Hst + Cst + ACT + [r for r in RES if r is not None]
        if p_drop > 0:
            saved += XD + [CTXD]
        saved.append(rng)
        saved += list(lstm_params)
        ctx.res_layers = [i for i in range(n_layers) if RES[i] is not None]
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(symbols, lengths)
        return embs, logps, symbols, lengths

    @staticmethod
    def _forward_persistent(ctx, k, enc, klens, ids_tf, S, need_logps, drop, emb_table, w_att, w_ffn, w_out, b_out,
                            lstm_params, lp, wih, whh, bias, wf, wo, wk):
        """Same outputs and saved buffers as the step-by-step path below, produced by ONE persistent launch
        (csrc/las_decoder.cu): bf16, D = 512, 3 layers, no dropout."""
        dt, dev, f32 = enc.dtype, enc.device, torch.float32
        B, Tk, H2 = enc.shape
        E, D, V, n_layers = emb_table.size(1), w_ffn.size(0), w_out.size(0), 3
        z = lambda *s, dtype=dt: torch.zeros(s, dtype=dtype, device=dev)
        e = lambda *s, dtype=dt: torch.empty(s, dtype=dtype, device=dev)
        CV = e(S + 1, B, D); CV[0].zero_()
        Hst = [e(S + 1, B, D) for _ in range(3)]
        Cst = [e(S + 1, B, D, dtype=f32) for _ in range(3)]
        for t in Hst + Cst:
            t[0].zero_()
        ACT = [e(S, B, 4 * D, dtype=f32) for _ in range(3)]
        RES = [None, e(S, B, D), None]
        CTX, PROBS = e(S, B, H2), e(S, B, Tk, dtype=f32)
        LOGITS = e(S, B, V) if need_logps else None
        lengths = torch.full((B,), S + 1, dtype=torch.int32, device=dev)       # Dec.py:163
        teacher = ids_tf is not None
        if teacher:      # every input token is known: its first-layer gate contribution for all steps is ONE GEMM
            ids_in = ids_tf.t()[:S].contiguous()
            SYM = e(S, B, dtype=torch.int64)
            EMB = k.embedding_fwd(ids_in.reshape(-1), emb_table, dt).view(S, B, E)
            gx0 = k.gemm(EMB.view(S * B, E), wih[0][:, :E], trans_b=True, bias=bias[0])
        else:            # free running: row of E W_ih0[:, :E]^T + b gathered by the fed-back token (Dec.py:199,223,383)
            IDS = e(S + 1, B, dtype=torch.int64)
            IDS[0].fill_(BOS)                                                  # Dec.py:158-160,199
            SYM = IDS[1:]
            gx0 = k.gemm(rt.operand(emb_table), wih[0][:, :E], trans_b=True, bias=bias[0])
        k.las_decoder_fwd(wk, enc, klens, gx0, [wih[0][:, E:], wih[1], wih[2]], whh, [None, bias[1], bias[2]], wf, wo, b_out,
                          CV, Hst, Cst, ACT, RES[1], CTX, PROBS, LOGITS, SYM, lengths, teacher)
        if not teacher:
            ids_in = IDS[:S]
            EMB = k.embedding_fwd(ids_in.reshape(-1), emb_table, dt).view(S, B, E)     # what backward multiplies dG0 with
        embs = k.transpose01(CV[1:])                                             # [B,S,D]
        if need_logps:
            logp_tm, _ = k.log_softmax_fwd(LOGITS.view(S * B, V))
            logps = k.transpose01(logp_tm.view(S, B, V))
        else:
            logp_tm = None
            logps = torch.empty(0, dtype=dt, device=dev)
        symbols = SYM.t().contiguous()
        ctx.geom = (B, Tk, H2, S, E, D, V, n_layers)
        ctx.need_logps = need_logps
        ctx.has_klens = klens is not None
        ctx.n_fixed = 10
        ctx.drop = (0.0, 0.0, 0, [[0] * n_layers for _ in range(S)], [0] * S, not teacher)
        saved = [enc, klens, ids_in.contiguous(), wk, CV, CTX, PROBS, EMB, logp_tm, emb_table, w_att, w_ffn, w_out]
        saved += Hst + Cst + ACT + [RES[1]]
        saved.append(None)
        saved += list(lstm_params)
        ctx.res_layers = [1]
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(symbols, lengths)
        return embs, logps, symbols, lengths

    @staticmethod
    def backward(ctx, d_embs, d_logps, _ds, _dl):
        k = K()
        B, Tk, H2, S, E, D, V, n_layers = ctx.geom
        sv = list(ctx.saved_tensors)
        (enc, klens, ids_in, wk, CV, CTX, PROBS, EMB, logp_tm, emb_table, w_att, w_ffn, w_out) = sv[:13]
        p = 13
        Hst = sv[p:p + n_layers]; p += n_layers
        Cst = sv[p:p + n_layers]; p += n_layers
        ACT = sv[p:p + n_layers]; p += n_layers
        RES = [None] * n_layers
        for i in ctx.res_layers:
            RES[i] = sv[p]; p += 1
        p_emb, p_drop, site_emb, site_l, site_att, fused_feed = ctx.drop
        XD = CTXD = None
        if p_drop > 0:
            XD = sv[p:p + n_layers]; p += n_layers
            CTXD = sv[p]; p += 1
        rng = sv[p]; p += 1
        lstm_params = sv[p:]
        lp = [lstm_params[4 * i:4 * i + 4] for i in range(n_layers)]
        dt, dev, f32 = enc.dtype, enc.device, torch.float32
        wih = [rt.operand(q[0]) for q in lp]
        whh = [rt.operand(q[1]) for q in lp]
        wf = rt.operand(w_ffn)
        e = lambda *s, dtype=dt: torch.empty(s, dtype=dtype, device=dev)

        # DCV[s] accumulates the total gradient w.r.t. cell_value of step s
        if d_embs is not None:
            DCV = k.transpose01(_c(d_embs))                                      # [S,B,D]
        else:
            DCV = torch.zeros((S, B, D), dtype=dt, device=dev)
        dw_out = db_out = None
        if ctx.need_logps and d_logps is not None and d_logps.numel() > 0:
            dlogits = k.log_softmax_bwd(k.transpose01(_c(d_logps)).view(S * B, V), logp_tm)
            k.gemm(dlogits, rt.operand(w_out), residual=DCV.view(S * B, D), out=DCV.view(S * B, D))
            dw_out = k.gemm(dlogits, CV[1:].reshape(S * B, D), trans_a=True, out_dtype=f32)
            db_out = k.colsum(dlogits)

        DG = [e(S, B, 4 * D) for _ in range(n_layers)]
        DCTX = e(S, B, H2)
        # Without context dropout the context's projection is folded into the attention values once, in front of the loop:
        # dP = (dcv W_fa) V^T = dcv (V W_fa^T)^T, so the attention backward reads dcv directly and the GEMM dcv W_fa leaves every
        # step's dependent chain (8 -> 7 kernels per step); DCTX (needed for d_enc only) is ONE GEMM over all steps after the loop.
        fold_ctx = p_drop == 0
        VW = k.gemm(enc.reshape(B * Tk, H2), wf[:, :H2], trans_b=True).view(B, Tk, D) if fold_ctx else None
        DSC = e(S, B, Tk, dtype=f32)
        DEMB = e(S, B, E)
        # Per-step critical path (4 GEMMs, 3 with fold_ctx): dcv W_fa -> attention bwd -> cell2 bwd -> dG2 W_ih2 -> cell1 bwd
        # -> dG1 W_ih1 -> cell0 bwd -> dG0 W_ih0[:, E:] (into DCV[s-1]).  The recurrent products dG_i W_hh_i (needed one step later),
        # dcv W_fb (needed after the attention backward) and dG0 W_ih0[:, :E] (needed after the loop) run on side streams.
        side = rt.side_streams(dev, n_layers + 1)
        DHN = [e(B, D) for _ in range(n_layers)]             # dh flowing to the previous step, per layer
        D_OUT = e(B, D)
        have_dhn = [False] * n_layers
        dc_next: List[Optional[torch.Tensor]] = [None] * n_layers
        for s in reversed(range(S)):
            dcv = DCV[s]
            # cell_value = ctx Wf[:, :2H]^T + dec_out Wf[:, 2H:]^T
            with rt.fork(side[n_layers]):
                k.gemm(dcv, wf[:, H2:], out=D_OUT)
            if fold_ctx:
                _, dq_att = k.las_attn_bwd(dcv, wk, VW, PROBS[s], dscore_out=DSC[s])
            else:
                k.gemm(dcv, wf[:, :H2], out=DCTX[s])
                # gradient of the dropped context -> gradient of the context
                k.dropout(DCTX[s], p_drop, rng, site_att[s], out=DCTX[s])
                _, dq_att = k.las_attn_bwd(DCTX[s], wk, enc, PROBS[s], dscore_out=DSC[s])
            rt.join(side[n_layers])
            # dec_out = y_{n-1};  y_i = h_i (+ y_{i-1} on residual layers, Dec.py:417-418)
            dy_parts = [D_OUT, dq_att]
            for i in reversed(range(n_layers)):
                if p_drop > 0:                   # dy_parts is the gradient of the DROPPED layer output x_{i+1}
                    dsum = dy_parts[0] if len(dy_parts) == 1 else k.add(dy_parts[0], dy_parts[1])
                    dy_parts = [k.dropout(dsum, p_drop, rng, site_l[s][i])]
                if have_dhn[i]:
                    rt.join(side[i])
                _, dc_next[i] = k.lstm_cell_bwd(dy_parts + [DHN[i] if have_dhn[i] else None], dc_next[i], ACT[i][s],
                                                Cst[i][s], Cst[i][s + 1], dt, dgates_out=DG[i][s])
                dgi = DG[i][s]
                with rt.fork(side[i]):
                    if s > 0:
                        k.gemm(dgi, whh[i], out=DHN[i])
                        have_dhn[i] = True
                    if i == 0:
                        k.gemm(dgi, wih[0][:, :E], out=DEMB[s])
                if i > 0:
                    if RES[i] is not None:      # the skip connection carries dy_i straight to y_{i-1}
                        skip = dy_parts[0] if len(dy_parts) == 1 else k.add(dy_parts[0], dy_parts[1])
                        dy_parts = [k.gemm(dgi, wih[i], residual=skip)]
                    else:
                        dy_parts = [k.gemm(dgi, wih[i])]
                elif s > 0:
                    k.gemm(dgi, wih[0][:, E:], residual=DCV[s - 1], out=DCV[s - 1])
        for st in side:
            rt.join(st)
        if fold_ctx:
            k.gemm(DCV.view(S * B, D), wf[:, :H2], out=DCTX.view(S * B, H2))

        SB = S * B
        dec_out_stack = RES[n_layers - 1] if RES[n_layers - 1] is not None else Hst[n_layers - 1][1:]
        if XD is not None:
            dec_out_stack = XD[n_layers - 1]
        # keys / values first -- d_enc is what the BLSTM backward is waiting for:
        # d wk[b] = dscore[:, b]^T dec_out[:, b];  d vals[b] = probs[:, b]^T dctx[:, b]
        d_wk = k.las_stack_grad(DSC, dec_out_stack.contiguous())                 # [B,Tk,D]   (one launch each: the fp32
        d_enc = k.las_stack_grad(PROBS, DCTX)                                    # [B,Tk,2H]   weights are read as they are)
        k.gemm(d_wk.view(B * Tk, D), rt.operand(w_att), residual=d_enc.view(B * Tk, H2),
               out=d_enc.view(B * Tk, H2))
        # every weight gradient of the loop is ONE GEMM over S*B rows; none of them is read again in backward: side
        # stream, joined at the end of backward (rt.defer).  Each returned gradient is its own whole tensor.
        flat_params = [q for layer in lp for q in layer]
        side = rt.side_streams(dev, 1, pool='dw') if rt.can_defer(emb_table, w_att, w_ffn, *flat_params) else [None]
        expect = []
        old_budget = k.set_gemm_sm_budget(LAS_DEFERRED_SM_BUDGET if side[0] is not None else 0)
        # (under the top BLSTM layer's backward recurrence: see DEFERRED_SM_BUDGET)
        try:
            with rt.fork(side[0]):
                grads_lstm = []
                for i in range(n_layers):
                    dg2 = DG[i].view(SB, 4 * D)
                    if i == 0:
                        dw_ih = torch.empty_like(lp[0][0])
                        k.gemm(dg2, EMB.view(SB, E), trans_a=True, out=dw_ih[:, :E])
                        k.gemm(dg2, CV[:S].reshape(SB, D), trans_a=True, out=dw_ih[:, E:])
                    else:
                        below = XD[i - 1] if XD is not None else (RES[i - 1] if RES[i - 1] is not None else Hst[i - 1][1:])
                        dw_ih = k.gemm(dg2, below.reshape(SB, D), trans_a=True, out_dtype=f32)
                    dw_hh = k.gemm(dg2, Hst[i][:S].reshape(SB, D), trans_a=True, out_dtype=f32)
                    db_i = k.colsum(dg2)
                    db_h = k.cast(db_i, f32, out=torch.empty_like(db_i))     # b_ih and b_hh share the values; autograd wants two tensors
                    grads_lstm += [dw_ih, dw_hh, db_i, db_h]
                    expect += list(zip(lp[i], (dw_ih, dw_hh, db_i, db_h)))
                dcv2 = DCV.view(SB, D)
                dw_ffn = torch.empty_like(w_ffn)
                k.gemm(dcv2, (CTXD if CTXD is not None else CTX).view(SB, H2), trans_a=True, out=dw_ffn[:, :H2])
                k.gemm(dcv2, dec_out_stack.reshape(SB, D), trans_a=True, out=dw_ffn[:, H2:])
                if p_emb > 0:
                    d_e = DEMB[0] if fused_feed else DEMB.view(SB, E)
                    k.dropout(d_e, p_emb, rng, site_emb, out=d_e)
                d_table = torch.zeros_like(emb_table)
                k.embedding_bwd(ids_in.reshape(-1), DEMB.view(SB, E), d_table, PAD)  # Dec.py:80-81 padding_idx
                dw_att = k.gemm(d_wk.view(B * Tk, D), enc.view(B * Tk, H2), trans_a=True, out_dtype=f32)
                expect += [(w_ffn, dw_ffn), (emb_table, d_table), (w_att, dw_att)]
        finally:
            k.set_gemm_sm_budget(old_budget)
        rt.defer(side[0], (DG, EMB, CV, Hst, RES, XD, DCV, CTX, CTXD, DEMB, ids_in, d_wk, enc, dec_out_stack), expect)
        return (d_enc, None, None, None, None, None, d_table, dw_att, dw_ffn, dw_out, db_out, *grads_lstm)


def las_decoder(enc, klens, ids_tf, n_steps, need_logps, emb_table, w_att, w_ffn, w_out, b_out,
                lstm_params, p_emb=0.0, p_drop=0.0):
    flat = [t for layer in lstm_params for t in layer]
    return _LASDecoder.apply(enc, klens, ids_tf, n_steps, need_logps, (p_emb, p_drop), emb_table, w_att, w_ffn,
                             w_out, b_out, *flat)
