"""Incremental (key/value-cached) Transformer decoder step for inference.

The reference re-runs the whole decoder stack on the growing prefix at every step of `forward_eval` /
`forward_translate` (Seq2seq.py:260-304, 307-393: O(L^2) decoder work, `decode_speedup` is never switched on).  Only
the newest position's output is used, and its self-attention needs the keys/values of the earlier positions — which
earlier steps already produced.  `DecoderCache` keeps them:

  * per layer one self-attention cache [n_hyp, max_len, 2*H*d] (K | V columns), filled one position per step straight
    from the K|V projection GEMM's epilogue (row-strided output view, no copy);
  * per layer the cross-attention K|V of the encoder output [B, S, 2*H*d], projected ONCE per utterance and shared by
    the beams of that utterance (the reference tiles the encoder output beam_width times and re-projects it each step);
  * beam re-ordering (Seq2seq.py:381-384) never moves the caches: an int32 ancestry table [max_len, n_hyp] records, for
    every hypothesis and position, the slot that holds that position, and the attention kernel (`b200st_mha_decode`)
    gathers through it.

Same arithmetic per row as the full decoder (LayerNorm on the query input only, K/V from the raw input, masked scores
set to -1e9, PAD tokens of the prefix masked as keys, final LayerNorm eps 1e-5), so greedy / beam token ids are the
ones the recompute path produces.
"""
from __future__ import annotations

import torch

from . import runtime as rt
from .kernels import K

PAD = 0


class DecoderCache:
    def __init__(self, model, enc_outputs: torch.Tensor, src_mask: torch.Tensor, beam_width: int, max_len: int):
        k = K()
        self.model = model
        dec = model.dec_tgt
        self.layers = list(dec.dec_layers)
        B, S, D = enc_outputs.shape
        self.B, self.S, self.D, self.k = B, S, D, beam_width
        self.n_hyp = B * beam_width
        self.max_len = max_len
        assert max_len <= dec.time_signal.shape[1], 'call expand_time() for longer sequences'
        dev, dt = enc_outputs.device, enc_outputs.dtype
        self.pe = dec._pe.on(dec.time_signal, dev)                       # fp32 [max_len, D] on the device
        if self.pe.dim() == 3:
            self.pe = self.pe[0]
        self.src_mask = src_mask.contiguous()                            # uint8 [B, 1, S]
        enc2 = enc_outputs.contiguous().view(B * S, D)
        self.kv_cross, self.kv_self = [], []
        for layer in self.layers:
            a = layer.encdec_attn
            self.kv_cross.append(k.gemm(enc2, rt.operand_cat(a.w_ks.weight, a.w_vs.weight), trans_b=True)
                                 .view(B, S, -1))
            hd2 = 2 * layer.decslf_attn.w_ks.weight.size(0)
            self.kv_self.append(torch.empty((self.n_hyp, max_len, hd2), dtype=dt, device=dev))
        # ancestry: slot holding position t of hypothesis b; identity until a beam re-ordering permutes it
        self.anc = None
        if beam_width > 1:
            self.anc = torch.arange(self.n_hyp, dtype=torch.int32, device=dev).expand(max_len, self.n_hyp).contiguous()
        # keys that are PAD tokens are masked (tgt_mask = pad & causal, Seq2seq.py:204-205)
        self.tokmask = torch.ones((self.n_hyp, max_len), dtype=torch.uint8, device=dev)

    def reorder(self, rows: torch.Tensor, n_pos: int):
        """Hypothesis h continues hypothesis rows[h] (its first n_pos positions)."""
        if self.anc is not None:
            self.anc[:n_pos] = self.anc[:n_pos][:, rows]
        self.tokmask[:, :n_pos] = self.tokmask[rows, :n_pos]

    def step(self, tok: torch.Tensor, pos: int) -> torch.Tensor:
        """Feeds token `tok` [n_hyp] int64 at position `pos`; returns the decoder output of that position [n_hyp, D]
        (after the final LayerNorm), i.e. row `pos` of what Decoder.forward would return for the prefix."""
        k, m = K(), self.model
        dt = rt.compute_dtype()
        self.tokmask[:, pos] = tok.ne(PAD)
        x = k.embedding_fwd(tok.contiguous(), m.dec_embedder.weight, dt)
        if m.dec_emb_proj_flag:
            x = k.gemm(x, rt.operand(m.dec_emb_proj.weight), trans_b=True)
        x = k.add_posenc(x.view(self.n_hyp, 1, self.D), self.pe[pos:pos + 1]).view(self.n_hyp, self.D)
        Lk = pos + 1
        for li, layer in enumerate(self.layers):
            a = layer.decslf_attn
            HD = a.w_qs.weight.size(0)
            qn, _, _ = k.layernorm_fwd(x, a.layer_norm.weight, a.layer_norm.bias, a.layer_norm.eps, save_stats=False)
            qp = k.gemm(qn, rt.operand(a.w_qs.weight), trans_b=True)
            cache = self.kv_self[li]
            k.gemm(x, rt.operand_cat(a.w_ks.weight, a.w_vs.weight), trans_b=True, out=cache[:, pos])
            o = k.mha_decode(qp, cache[:, :, :HD], cache[:, :, HD:], Lk, a.n_head, a.attention.temperature,
                             anc=self.anc, mask=self.tokmask)
            x = k.gemm(o, rt.operand(a.fc.weight), trans_b=True, residual=x)
            a = layer.encdec_attn
            qn, _, _ = k.layernorm_fwd(x, a.layer_norm.weight, a.layer_norm.bias, a.layer_norm.eps, save_stats=False)
            qp = k.gemm(qn, rt.operand(a.w_qs.weight), trans_b=True)
            kv = self.kv_cross[li]
            o = k.mha_decode(qp, kv[:, :, :HD], kv[:, :, HD:], self.S, a.n_head, a.attention.temperature,
                             bdiv=self.k, mask=self.src_mask, mask_bdiv=self.k)
            x = k.gemm(o, rt.operand(a.fc.weight), trans_b=True, residual=x)
            f = layer.pos_ffn
            y, _, _ = k.layernorm_fwd(x, f.layer_norm.weight, f.layer_norm.bias, f.layer_norm.eps, save_stats=False)
            h = k.gemm(y, rt.operand(f.w_1.weight), trans_b=True, bias=f.w_1.bias, relu=True)
            x = k.gemm(h, rt.operand(f.w_2.weight), trans_b=True, bias=f.w_2.bias, residual=x)
        dec = m.dec_tgt
        x, _, _ = k.layernorm_fwd(x, dec.norm.weight, dec.norm.bias, dec.norm.eps, save_stats=False)
        return x

    def step_logps(self, tok: torch.Tensor, pos: int):
        """(log-probabilities [n_hyp, V] in the compute dtype, arg-max [n_hyp, 1]) of the token after position pos."""
        k = K()
        x = self.step(tok, pos)
        logits = k.gemm(x, rt.operand(self.model.out_tgt.weight), trans_b=True)
        logp, am = k.log_softmax_fwd(logits, want_argmax=True)
        return logp, am.view(-1, 1)
