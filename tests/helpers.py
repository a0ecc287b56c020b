"""Shared test helpers: build the product model from a golden fixture / oracle params, run a train step."""
import numpy as np
import torch


def build_model(cfg, params, device='cpu', mode='ST'):
    from models.Seq2seq import Seq2seq
    m = Seq2seq(cfg.enc_vocab_size, cfg.dec_vocab_size, share_embedder=False,
                enc_embedding_size=cfg.enc_embedding_size, dec_embedding_size=cfg.dec_embedding_size,
                max_seq_len_src=cfg.max_seq_len_src, max_seq_len_tgt=cfg.max_seq_len_tgt,
                num_heads=cfg.num_heads, dim_model=cfg.dim_model, dim_feedforward=cfg.dim_feedforward,
                enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, embedding_dropout=0.0, dropout=0.0,
                acous_dim=cfg.acous_dim, acous_hidden_size=cfg.acous_hidden_size, mode=mode,
                load_mode='null')
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    for mod in m.modules():                     # the hidden attention dropout p=0.1 (layers.py:207)
        if type(mod).__name__ == 'ScaledDotProductAttention':
            mod.dropout.p = 0.0
    return m.to(device)


def train_step(model, batch, device, mode='ST'):
    """Mirror of Trainer_ST._train_batch's inner body (trainer_st.py:253-288), one minibatch, coeff 1."""
    from modules.loss import NLLLoss
    src = batch['src'].to(device)
    tgt = batch['tgt'].to(device)
    feats = batch['acous_feats'].to(device)
    lens = [torch.tensor([n]) for n in batch['acous_lens']]
    out = model.forward_train(src, tgt=tgt, acous_feats=feats, acous_lens=lens, mode=mode,
                              use_gpu=(device != 'cpu'))
    key = 'logps_st' if 'ST' in mode else 'logps_mt'
    logps = out[key][:, :-1, :]
    loss = NLLLoss()
    loss.reset()
    mask = tgt.data.ne(0)
    loss.eval_batch_with_mask(logps.reshape(-1, logps.size(-1)), tgt[:, 1:].reshape(-1),
                              mask[:, 1:].reshape(-1))
    loss.norm_term = 1.0 * torch.sum(mask[:, 1:])
    loss.normalise()
    return loss, out
