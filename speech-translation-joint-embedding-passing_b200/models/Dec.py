"""LAS attention-LSTM decoder (produces the *dynamic embedding*) on b200st kernels.

Mirror of the reference's models/Dec.py (class Dec): same constructor, `forward(acous_outputs, acous_lens,
tgt, hidden, is_training, teacher_forcing_ratio, beam_width, use_gpu, lm_mode, lm_model)` returning
`(sequence_embs [B,S,D], sequence_logps [B,S,V], sequence_symbols [B,S,1], lengths ndarray[B])`, same
parameter names (`embedder`, `acous_att.linear_att_w`, `acous_ffn`, `acous_out`, `dec.l{i}`).

What changed underneath: the whole S-step loop is one autograd node
(b200st.functional.las_decoder) — the step-invariant key projection W.K is computed once instead of every
step (attention.py:192), the concat inputs of the first LSTM and of acous_ffn are consumed as two
accumulating GEMMs instead of being materialised, arg-max feedback and the EOS/PAD length rule
(Dec.py:320-341) stay on the device (one D2H copy at the end instead of one per step), and backward is a
hand-written BPTT.  N-gram LM fusion (`add_lm`, Dec.py:236-317) needs an external pickled LM and is out of
scope: only lm_mode='null' is accepted.
"""
import random

import numpy as np
import torch
import torch.nn as nn

from b200st import functional as BF
from b200st import runtime as rt
from modules.attention import AttentionLayer
from b200st.hostutil import PAD, EOS, BOS
from b200st.hostutil import check_device

KEY_ATTN_SCORE = 'attention_score'
KEY_ATTN_OUT = 'attention_out'
KEY_LENGTH = 'length'
KEY_SEQUENCE = 'sequence'


class Dec(nn.Module):

    def __init__(self, vocab_size, embedding_size=200, acous_hidden_size=256, acous_att_mode='bahdanau',
                 hidden_size_dec=200, hidden_size_shared=200, num_unilstm_dec=4, embedding_dropout=0,
                 dropout=0.0, residual=True, batch_first=True, max_seq_len=32, embedder=None,
                 word2id=None, id2word=None, hard_att=False):
        super().__init__()
        if not residual or not batch_first or hard_att:
            raise NotImplementedError('b200st Dec implements residual=True, batch_first=True, hard_att=False '
                                      '(what Seq2seq constructs, Seq2seq.py:160-168)')
        if hidden_size_dec != hidden_size_shared:
            raise NotImplementedError('hidden_size_dec must equal hidden_size_shared (Seq2seq.py:151-152)')
        self.acous_hidden_size = acous_hidden_size
        self.acous_att_mode = acous_att_mode
        self.hidden_size_dec = hidden_size_dec
        self.hidden_size_shared = hidden_size_shared
        self.num_unilstm_dec = num_unilstm_dec
        self.hard_att = hard_att
        self.residual = residual
        self.max_seq_len = max_seq_len
        self.vocab_size = vocab_size
        self.embedding_size = embedding_size
        self.word2id, self.id2word = word2id, id2word
        self.embedding_dropout = nn.Dropout(embedding_dropout)
        self.dropout = nn.Dropout(dropout)
        if embedder is not None:
            self.embedder = embedder
        else:
            self.embedder = nn.Embedding(vocab_size, embedding_size, sparse=False, padding_idx=PAD)
        self.acous_hidden_size_att = 0
        self.acous_key_size = acous_hidden_size * 2
        self.acous_value_size = acous_hidden_size * 2
        self.acous_query_size = hidden_size_dec
        self.acous_att = AttentionLayer(self.acous_query_size, self.acous_key_size,
                                        value_size=self.acous_value_size, mode=acous_att_mode,
                                        dropout=dropout, query_transform=False, output_transform=False,
                                        hidden_size=self.acous_hidden_size_att, hard_att=False)
        self.acous_ffn = nn.Linear(acous_hidden_size * 2 + hidden_size_dec, hidden_size_shared, bias=False)
        self.acous_out = nn.Linear(hidden_size_shared, vocab_size, bias=True)
        self.dec = nn.Module()                                             # Dec.py:108-118
        self.dec.add_module('l0', torch.nn.LSTM(embedding_size + hidden_size_shared, hidden_size_dec,
                                                num_layers=1, batch_first=batch_first, bias=True,
                                                dropout=dropout, bidirectional=False))
        for i in range(1, num_unilstm_dec):
            self.dec.add_module('l' + str(i), torch.nn.LSTM(hidden_size_dec, hidden_size_dec, num_layers=1,
                                                            batch_first=batch_first, bias=True,
                                                            dropout=dropout, bidirectional=False))

    def check_var(self, var_name, var_val_set=None):
        if not hasattr(self, var_name):
            setattr(self, var_name, var_val_set if var_val_set is not None else None)

    def _lstm_params(self):
        out = []
        for i in range(self.num_unilstm_dec):
            m = getattr(self.dec, 'l' + str(i))
            out.append((m.weight_ih_l0, m.weight_hh_l0, m.bias_ih_l0, m.bias_hh_l0))
        return out

    def forward_device(self, acous_outputs, klens, tgt=None, teacher_forcing_ratio=0.0, need_logps=True):
        """Device-resident variant used by Seq2seq: klens int32[B] (valid keys) or None; returns
        (embs, logps|empty, symbols int64 [B,S,1], lengths int32 device tensor [B])."""
        if tgt is None:
            n_steps = self.max_seq_len - 1                                  # Dec.py:158-162,205
        else:
            n_steps = tgt.size(1) - 1
        # one draw per forward, exactly like Dec.py:196
        use_teacher_forcing = True if random.random() < teacher_forcing_ratio else False
        ids_tf = tgt if (use_teacher_forcing and tgt is not None) else None
        if not use_teacher_forcing and tgt is not None:
            # free running but with a caller-provided first token row (Dec.py:199 uses emb_tgt[:,0])
            assert bool((tgt[:, 0] == BOS).all()), 'free-running decode starts from BOS'
        embs, logps, symbols, lengths = BF.las_decoder(
            acous_outputs, klens, ids_tf, n_steps, need_logps, self.embedder.weight,
            self.acous_att.linear_att_w.weight, self.acous_ffn.weight, self.acous_out.weight,
            self.acous_out.bias, self._lstm_params(),
            p_emb=float(self.embedding_dropout.p) if self.training else 0.0,
            p_drop=float(self.dropout.p) if self.training else 0.0)
        return embs, logps, symbols.unsqueeze(2), lengths

    def forward(self, acous_outputs, acous_lens=None, tgt=None, hidden=None, is_training=False,
                teacher_forcing_ratio=0.0, beam_width=1, use_gpu=False, lm_mode='null', lm_model=None):
        if lm_mode != 'null':
            raise NotImplementedError("LM fusion (Dec.add_lm) is out of scope; use lm_mode='null'")
        klens = None
        if acous_lens is not None:                                          # Dec.py:173-181
            from .Enc import padded_lengths
            ln, _ = padded_lengths(acous_lens, acous_outputs.size(0), 0, acous_outputs.device)
            klens = ln // 8
        embs, logps, symbols, lengths = self.forward_device(
            acous_outputs, klens, tgt=tgt, teacher_forcing_ratio=teacher_forcing_ratio, need_logps=True)
        return embs, logps, symbols, lengths.cpu().numpy().astype(np.int64)
