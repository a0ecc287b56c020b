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


def position_signal(module):
    from modules.layers import position_signal as f
    return f(module)

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
        if max_len > dec.time_signal.shape[1]:       # the API default max_seq_len = 900 exceeds the 500-row table the
            dec.expand_time(max_len)                 # constructor builds; rows < 500 are unchanged (layers.py:293-309)
        dev, dt = enc_outputs.device, enc_outputs.dtype
        self.pe = position_signal(dec).on(dec.time_signal, dev)                       # fp32 [max_len, D] on the device
        if self.pe.dim() == 3:
            self.pe = self.pe[0]
        self.src_mask = src_mask.contiguous()                            # uint8 [B, 1, S]
        self.enc = enc_outputs.contiguous()
        self.kv_cross, self.kv_self = [], []
        for layer in self.layers:
            hd2c = 2 * layer.encdec_attn.w_ks.weight.size(0)
            self.kv_cross.append(torch.empty((B, S, hd2c), dtype=dt, device=dev))
            hd2 = 2 * layer.decslf_attn.w_ks.weight.size(0)
            self.kv_self.append(torch.empty((self.n_hyp, max_len, hd2), dtype=dt, device=dev))
        self.refresh_cross()
        # ancestry: slot holding position t of hypothesis b; identity until a beam re-ordering permutes it
        self.anc = None
        if beam_width > 1:
            self.anc = torch.arange(self.n_hyp, dtype=torch.int32, device=dev).expand(max_len, self.n_hyp).contiguous()
        # keys that are PAD tokens are masked (tgt_mask = pad & causal, Seq2seq.py:204-205)
        self.tokmask = torch.ones((self.n_hyp, max_len), dtype=torch.uint8, device=dev)

    def refresh_cross(self):
        """(Re)project the encoder output into every layer's cross-attention K|V buffer (in place)."""
        k = K()
        enc2 = self.enc.view(self.B * self.S, self.D)
        for layer, kv in zip(self.layers, self.kv_cross):
            a = layer.encdec_attn
            k.gemm(enc2, rt.operand_cat(a.w_ks.weight, a.w_vs.weight), trans_b=True, out=kv.view(self.B * self.S, -1))

    def reset(self, enc_outputs: torch.Tensor, src_mask: torch.Tensor):
        """Re-use the buffers for a new batch of the same shape (everything a captured step graph points at stays put)."""
        self.enc.copy_(enc_outputs)
        self.src_mask.copy_(src_mask)
        self.refresh_cross()
        if self.anc is not None:
            self.anc.copy_(torch.arange(self.n_hyp, dtype=torch.int32, device=self.anc.device).expand_as(self.anc))
        self.tokmask.fill_(1)

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


BOS, EOS = 2, 3


class BeamSearch:
    """The step loop of forward_translate (`_prep_translate` + `_step_translate`, Seq2seq.py:307-393,720-739) over a
    `DecoderCache`, with every piece of state in static device buffers so that a decode position is a fixed launch
    sequence: positions >= 3 are captured once into one CUDA graph each and replayed (the eager loop is bound by host
    launch overhead: ~90 launches per token).  Same arithmetic and selection rule as the reference, including the
    length penalty `score / len^alpha`, the EOS masking and the final `reshape(batch, -1)[:, :max_seq_len]` slice quirk
    (Seq2seq.py:738); one host read per step remains (the reference's all-EOS early exit, Seq2seq.py:388-393)."""

    def __init__(self, model, enc_outputs, src_mask, beam_width, penalty_factor, max_len, graphs=True):
        dev = enc_outputs.device
        self.B, self.k, self.max_len, self.penalty = enc_outputs.size(0), beam_width, max_len, penalty_factor
        self.n_hyp = self.B * beam_width
        self.cache = DecoderCache(model, enc_outputs.clone(), src_mask.clone(), beam_width, max_len)
        self.preds = torch.zeros((self.n_hyp, max_len), dtype=torch.int64, device=dev)
        self.scores = torch.zeros(self.n_hyp, device=dev)
        self.eos = torch.zeros(self.n_hyp, dtype=torch.bool, device=dev)
        self.len_map = torch.ones(self.n_hyp, device=dev)
        self.n_done = torch.zeros((), dtype=torch.int64, device=dev)
        self.done_u = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.offs = torch.arange(0, self.n_hyp * beam_width, beam_width * beam_width, device=dev).float().reshape(self.B, 1)
        self.graphs = {}
        self.use_graphs = graphs and dev.type == 'cuda'
        self.epoch = None
        self.pool = None
        self._fresh = True

    def _reset(self, enc_outputs, src_mask):
        if not self._fresh:
            self.cache.reset(enc_outputs, src_mask)
        self._fresh = False
        self.preds.zero_()
        self.preds[:, 0] = BOS
        self.scores.zero_()
        self.eos.zero_()
        self.len_map.fill_(1.0)
        self.n_done.zero_()
        self.ticket.zero_()

    def _step(self, i: int):
        """One decode position: incremental decoder -> vocabulary projection -> top-k log-probabilities -> the reference's
        selection rule and re-ordering (Seq2seq.py:337-393), the last two as ONE kernel each (csrc/beam.cu)."""
        k = K()
        x = self.cache.step(self.preds[:, i - 1], i - 1)
        logits = k.gemm(x, rt.operand(self.cache.model.out_tgt.weight), trans_b=True)
        score, pred = k.topk_logsoftmax(logits, self.k)
        k.beam_select(self.scores, score, pred, self.eos, self.len_map, self.penalty, i, i == 1, self.preds, self.cache.anc,
                      self.cache.tokmask, self.k, self.done_u, self.ticket, self.n_done)

    def run(self, enc_outputs, src_mask):
        if self.use_graphs:
            rt.refresh_all()                     # the graphs read cached weight copies without re-casting them
            if self.epoch != rt.cache_epoch() and self.graphs:
                self.graphs = {}                 # the copies were re-allocated (compute dtype switch): recapture
        self._reset(enc_outputs, src_mask)
        width = 1
        for i in range(1, self.max_len):
            if not self.use_graphs or i < 3:
                self._step(i)                    # eager: also creates every cached weight operand before any capture
            else:
                g = self.graphs.get(i)
                if g is None:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self.pool):
                        self._step(i)
                    self.pool = self.pool or g.pool()
                    self.graphs[i] = g
                    self.epoch = rt.cache_epoch()
                    # the capture itself did not execute the step: replay it below
                g.replay()
            width = i + 1
            if int(self.n_done) == self.n_hyp:   # the reference's early exit, one host read per step
                break
        return self.preds[:, :width].reshape(self.B, -1)[:, :self.max_len].contiguous()
