"""Listen-attend-spell wrapper (reference: models/Las.py:17-123): Enc then Dec, same signature."""
import torch
import torch.nn as nn

from .Dec import Dec
from .Enc import Enc, padded_lengths


class LAS(nn.Module):

    def __init__(self, vocab_size, embedding_size=200, acous_dim=26, acous_hidden_size=256,
                 acous_att_mode='bahdanau', hidden_size_dec=200, hidden_size_shared=200,
                 num_unilstm_dec=4, acous_norm=False, spec_aug=False, batch_norm=False,
                 enc_mode='pyramid', embedding_dropout=0, dropout=0.0, residual=True, batch_first=True,
                 max_seq_len=32, embedder=None, word2id=None, id2word=None, hard_att=False):
        super().__init__()
        self.encoder = Enc(acous_dim=acous_dim, acous_hidden_size=acous_hidden_size,
                           acous_norm=acous_norm, spec_aug=spec_aug, batch_norm=batch_norm,
                           enc_mode=enc_mode, dropout=dropout, batch_first=batch_first)
        self.decoder = Dec(vocab_size=vocab_size, embedding_size=embedding_size,
                           acous_hidden_size=acous_hidden_size, acous_att_mode=acous_att_mode,
                           hidden_size_dec=hidden_size_dec, hidden_size_shared=hidden_size_shared,
                           num_unilstm_dec=num_unilstm_dec, embedding_dropout=embedding_dropout,
                           dropout=dropout, residual=residual, batch_first=batch_first,
                           max_seq_len=max_seq_len, embedder=embedder, word2id=word2id, id2word=id2word,
                           hard_att=hard_att)

    def check_var(self, var_name, var_val_set=None):
        if not hasattr(self, var_name):
            setattr(self, var_name, var_val_set if var_val_set is not None else None)

    def forward_device(self, acous_feats, acous_lens=None, tgt=None, is_training=False,
                       teacher_forcing_ratio=0.0, need_logps=True):
        """Like forward() but `lengths` stays an int32 device tensor (no host sync)."""
        if teacher_forcing_ratio > 0.1:
            assert tgt is not None
        lens_dev, host = padded_lengths(acous_lens, acous_feats.size(0), acous_feats.size(1),
                                        acous_feats.device)
        if host is not None:
            assert max(host) == acous_feats.size(1), 'padded max length must equal the feature length'
        acous_outputs = self.encoder(acous_feats, acous_lens=acous_lens, is_training=is_training,
                                     lens_dev=lens_dev)
        klens = (lens_dev // 8) if acous_lens is not None else None
        return self.decoder.forward_device(acous_outputs, klens, tgt=tgt,
                                           teacher_forcing_ratio=teacher_forcing_ratio,
                                           need_logps=need_logps)

    def forward(self, acous_feats, acous_lens=None, tgt=None, hidden=None, is_training=False,
                teacher_forcing_ratio=0.0, beam_width=1, use_gpu=False, lm_mode='null', lm_model=None):
        if lm_mode != 'null':
            raise NotImplementedError("LM fusion is out of scope; use lm_mode='null'")
        if self.training:
            from b200st import runtime as rt
            rt.new_step()
        embs, logps, symbols, lengths = self.forward_device(
            acous_feats, acous_lens=acous_lens, tgt=tgt, is_training=is_training,
            teacher_forcing_ratio=teacher_forcing_ratio, need_logps=True)
        return embs, logps, symbols, lengths.cpu().numpy().astype('int64')
