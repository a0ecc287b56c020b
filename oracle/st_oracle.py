"""CPU oracle for the joint speech-translation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32/fp64, CPU) restatement of the reference's algorithm for the
forward pass of the joint ST model and its masked-NLL loss.  It is *not* product code: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it.  The product path (`speech-translation-joint-embedding-passing_b200/`) never does.

Parity status: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, generated in the build container by
`oracle/make_golden.py` (imports /root/reference unmodified) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` checks this file against every one of those fixtures.

All third-party arithmetic in the reference is PyTorch (torch.nn.LSTM, Linear, LayerNorm, softmax,
log_softmax, topk), so the oracle calls the same torch primitives at the same call
sites, which also makes its CPU timing representative of the reference's CPU path.  The LSTM is
additionally restated from first principles (`lstm_layer_loops`) for small cases.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Parameters are addressed by the reference's own state_dict names, e.g.
`las.encoder.acous_enc_l1.weight_ih_l0`, `enc_src.enc_layers.0.slf_attn.w_qs.weight`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

PAD, UNK, BOS, EOS, SPC = 0, 1, 2, 3, 4  # utils/config.py:7

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class STConfig:
    """Hyper-parameters of models/Seq2seq.py:30-61 that shape the hot path."""
    enc_vocab_size: int = 10000
    dec_vocab_size: int = 10000
    enc_embedding_size: int = 200
    dec_embedding_size: int = 200
    max_seq_len_src: int = 32
    max_seq_len_tgt: int = 50
    num_heads: int = 8
    dim_model: int = 512
    dim_feedforward: int = 1024
    enc_layers: int = 6
    dec_layers: int = 6
    acous_dim: int = 80
    acous_hidden_size: int = 256
    num_unilstm_dec: int = 3      # Seq2seq.py:153

    @property
    def d_k(self) -> int:
        return int(self.dim_model / self.num_heads)   # TFEnc.py:44-45

    @property
    def has_dec_emb_proj(self) -> bool:
        return self.dec_embedding_size != self.dim_model  # Seq2seq.py:128


def padded_len(n):
    """The reference's pad rule `n + 8 - n % 8` (Enc.py:142, Dec.py:175, trainer_st.py:252).
    Adds a full 8 when n is already a multiple of 8."""
    return n + 8 - n % 8


# --------------------------------------------------------------------------------------------
# parameter initialisation with the reference's parameter names / shapes
# --------------------------------------------------------------------------------------------
def param_shapes(cfg: STConfig, with_templates: bool = True) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every parameter `Seq2seq(mode='ST')` registers (Seq2seq.py:98-180, Las.py:50-79,
    Enc.py:50-66, Dec.py:74-118, TFEnc.py:51-61, TFDec.py:48-58, layers.py:131-139,238-241)."""
    D, FF, H, E_s, E_t = (cfg.dim_model, cfg.dim_feedforward, cfg.acous_hidden_size,
                          cfg.enc_embedding_size, cfg.dec_embedding_size)
    s: Dict[str, Tuple[int, ...]] = {}
    s['enc_embedder.weight'] = (cfg.enc_vocab_size, E_s)
    s['dec_embedder.weight'] = (cfg.dec_vocab_size, E_t)
    s['enc_emb_proj.weight'] = (D, E_s + D)
    if cfg.has_dec_emb_proj:
        s['dec_emb_proj.weight'] = (D, E_t)
    for layer in range(1, 5):
        inp = cfg.acous_dim if layer == 1 else 4 * H
        for sfx in ('', '_reverse'):
            base = f'las.encoder.acous_enc_l{layer}.'
            s[base + 'weight_ih_l0' + sfx] = (4 * H, inp)
            s[base + 'weight_hh_l0' + sfx] = (4 * H, H)
            s[base + 'bias_ih_l0' + sfx] = (4 * H,)
            s[base + 'bias_hh_l0' + sfx] = (4 * H,)
    s['las.decoder.embedder.weight'] = (cfg.enc_vocab_size, E_s)
    s['las.decoder.acous_att.linear_att_w.weight'] = (D, 2 * H)
    s['las.decoder.acous_ffn.weight'] = (D, 2 * H + D)
    s['las.decoder.acous_out.weight'] = (cfg.enc_vocab_size, D)
    s['las.decoder.acous_out.bias'] = (cfg.enc_vocab_size,)
    for i in range(cfg.num_unilstm_dec):
        inp = E_s + D if i == 0 else D
        base = f'las.decoder.dec.l{i}.'
        s[base + 'weight_ih_l0'] = (4 * D, inp)
        s[base + 'weight_hh_l0'] = (4 * D, D)
        s[base + 'bias_ih_l0'] = (4 * D,)
        s[base + 'bias_hh_l0'] = (4 * D,)

    def mha(prefix):
        s[prefix + 'w_qs.weight'] = (D, D)
        s[prefix + 'w_ks.weight'] = (D, D)
        s[prefix + 'w_vs.weight'] = (D, D)
        s[prefix + 'fc.weight'] = (D, D)
        s[prefix + 'layer_norm.weight'] = (D,)
        s[prefix + 'layer_norm.bias'] = (D,)

    def ffn(prefix):
        s[prefix + 'w_1.weight'] = (FF, D)
        s[prefix + 'w_1.bias'] = (FF,)
        s[prefix + 'w_2.weight'] = (D, FF)
        s[prefix + 'w_2.bias'] = (D,)
        s[prefix + 'layer_norm.weight'] = (D,)
        s[prefix + 'layer_norm.bias'] = (D,)

    enc_prefixes = [f'enc_src.enc_layers.{i}.' for i in range(cfg.enc_layers)]
    dec_prefixes = [f'dec_tgt.dec_layers.{i}.' for i in range(cfg.dec_layers)]
    if with_templates:   # template layers are registered but never run (TFEnc.py:51-58)
        enc_prefixes = ['enc_src.enc.'] + enc_prefixes
        dec_prefixes = ['dec_tgt.dec.'] + dec_prefixes
    for p in enc_prefixes:
        mha(p + 'slf_attn.')
        ffn(p + 'pos_ffn.')
    s['enc_src.norm.weight'] = (D,)
    s['enc_src.norm.bias'] = (D,)
    for p in dec_prefixes:
        mha(p + 'decslf_attn.')
        mha(p + 'encdec_attn.')
        ffn(p + 'pos_ffn.')
    s['dec_tgt.norm.weight'] = (D,)
    s['dec_tgt.norm.bias'] = (D,)
    s['out_tgt.weight'] = (cfg.dec_vocab_size, D)
    return s


def init_params(cfg: STConfig, seed: int = 333, dtype=torch.float32, scale: float = 1.0) -> Params:
    """Deterministic synthetic weights with PyTorch-like fan-in scaling.  Used when no reference
    state_dict is at hand (GPU box).  Not bit-identical to the reference's own init — parity tests
    always load the *same* dict into both sides, so only the distribution matters."""
    g = torch.Generator().manual_seed(seed)
    out: Params = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith('layer_norm.weight') or name.endswith('norm.weight'):
            t = torch.ones(shape) + 0.05 * torch.randn(shape, generator=g)
        elif name.endswith('layer_norm.bias') or name.endswith('norm.bias'):
            t = 0.05 * torch.randn(shape, generator=g)
        elif 'embedder' in name:
            t = torch.randn(shape, generator=g)
            t[PAD].zero_()                                   # padding_idx=PAD (Seq2seq.py:106-107)
        elif '.acous_enc_l' in name or '.dec.l' in name:     # nn.LSTM: U(-1/sqrt(H), 1/sqrt(H))
            hid = cfg.acous_hidden_size if '.acous_enc_l' in name else cfg.dim_model
            k = 1.0 / math.sqrt(hid)
            t = (torch.rand(shape, generator=g) * 2 - 1) * k
        else:                                                # nn.Linear default
            fan_in = shape[-1] if len(shape) > 1 else shape[0]
            k = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * k
        out[name] = (t * scale).to(dtype)
    return out


# --------------------------------------------------------------------------------------------
# LSTM (reference: torch.nn.LSTM at Enc.py:50-66,153-209 and Dec.py:104-118,395-415)
# --------------------------------------------------------------------------------------------
# Dropout hook (training-mode nn.Dropout call sites of the reference).  None = every dropout at p = 0, which is what
# the golden fixtures and the parity contract use.  Tests that exercise dropout install
# DROP(x, tag) -> x * mask / (1 - p) with the mask the CUDA path drew for the site called `tag`, so that the comparison
# stays exact; the tags name the reference call sites (see each _drop call).
DROP = None


def _drop(x, tag: str):
    return x if DROP is None else DROP(x, tag)


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """One LSTM step, PyTorch gate order (i, f, g, o).  This is what torch.nn.LSTM computes for a
    length-1 sequence (Dec.py:395-415)."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=-1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def lstm_layer_loops(x, lens, w_ih, w_hh, b_ih, b_hh, reverse: bool):
    """First-principles restatement of one direction of a packed-sequence LSTM layer
    (pack_padded_sequence -> nn.LSTM -> pad_packed_sequence, Enc.py:150-157): each sequence runs over
    its own `lens[b]` steps from a zero state, the reverse direction starts at the sequence's own
    last frame, and outputs beyond the length are zero.  Python loops: small cases only."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    out = x.new_zeros(B, T, H)
    for b in range(B):
        h = x.new_zeros(1, H)
        c = x.new_zeros(1, H)
        steps = range(int(lens[b]) - 1, -1, -1) if reverse else range(int(lens[b]))
        rows = {}
        for t in steps:
            h, c = lstm_cell(x[b:b + 1, t], h, c, w_ih, w_hh, b_ih, b_hh)
            rows[t] = h
        if rows:
            out[b, :int(lens[b])] = torch.cat([rows[t] for t in range(int(lens[b]))], 0)
    return out


def blstm_layer(P: Params, prefix: str, x, lens, loops: bool = False):
    """Bidirectional packed LSTM layer as called at Enc.py:150-157.  `lens` is an int64 CPU tensor."""
    names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
    fw = [P[prefix + n] for n in names]
    bw = [P[prefix + n + '_reverse'] for n in names]
    if loops:
        return torch.cat([lstm_layer_loops(x, lens, *fw, reverse=False),
                          lstm_layer_loops(x, lens, *bw, reverse=True)], dim=-1)
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
    hid = fw[1].shape[1]
    n_seq = x.shape[0]
    zeros = x.new_zeros(2, n_seq, hid)
    if packed.sorted_indices is not None:
        pass  # zero state: permutation is irrelevant
    out_data, _, _ = torch._VF.lstm(packed.data, packed.batch_sizes, (zeros, zeros), fw + bw,
                                    True, 1, 0.0, False, True)
    out_packed = torch.nn.utils.rnn.PackedSequence(out_data, packed.batch_sizes,
                                                   packed.sorted_indices, packed.unsorted_indices)
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out_packed, batch_first=True)
    return out


def las_encoder(P: Params, cfg: STConfig, acous_feats, acous_lens, loops: bool = False):
    """Pyramidal BLSTM (Enc.forward, Enc.py:120-223) in ST mode: no SpecAug (is_training=False,
    Seq2seq.py:485-487), dropout 0, batch_norm off (Seq2seq.py:157).
    acous_feats [B, T, F] with T a multiple of 8; acous_lens: sequence of raw lengths or None."""
    B, T, _ = acous_feats.shape
    if acous_lens is None:
        lens = torch.full((B,), T, dtype=torch.int64)                       # Enc.py:139-140
    else:
        lens = torch.tensor([int(padded_len(int(n))) for n in acous_lens])  # Enc.py:142
    x = acous_feats
    for layer in range(1, 5):
        t_l = T // (2 ** (layer - 1))
        out = blstm_layer(P, f'las.encoder.acous_enc_l{layer}.', x, lens, loops=loops)
        # pad_packed_sequence returns max(lens) frames; the reference then reshapes to the full
        # length (Enc.py:159-160), which requires max(lens) == t_l.
        assert out.shape[1] == t_l, 'max padded length must equal the feature length'
        out = _drop(out, f'las.enc.l{layer}')                               # Enc.py:159,178,195,212
        if layer < 4:
            x = out.reshape(B, t_l // 2, 2 * out.shape[-1])                 # Enc.py:166-167
            lens = lens // 2                                                # Enc.py:170
        else:
            x = out
    return x


# --------------------------------------------------------------------------------------------
# LAS attention decoder (Dec.py:130-233, 320-438; attention.py:190-193,203-289 bilinear mode)
# --------------------------------------------------------------------------------------------
def las_decoder(P: Params, cfg: STConfig, acous_outputs, acous_lens, tgt=None,
                teacher_forcing: bool = False, hoist_keys: bool = False):
    """Returns (sequence_embs [B,S,D], sequence_logps [B,S,V], symbols [B,S,1] int64, lengths list[int]).

    Free-running (argmax feedback) when teacher_forcing is False, which is what ST mode uses
    (Seq2seq.py:485-487 -> Dec.py:196 with ratio 0.0).  `hoist_keys=True` computes the step-invariant
    bilinear key projection once (same numbers) — used to keep the CPU baseline from
    being charged for the reference's redundant work only when explicitly asked."""
    B = acous_outputs.shape[0]
    D = cfg.dim_model
    pre = 'las.decoder.'
    if tgt is None:
        tgt = torch.full((B, cfg.max_seq_len_src), BOS, dtype=torch.int64)  # Dec.py:158-160
    S_full = tgt.shape[1]
    lengths = [S_full] * B                                                  # Dec.py:163
    emb_tgt = F.embedding(tgt, P[pre + 'embedder.weight'], padding_idx=PAD)  # Dec.py:166
    emb_tgt = _drop(emb_tgt, 'las.dec.emb')                                 # embedding_dropout, Dec.py:166

    t_k = acous_outputs.shape[1]
    if acous_lens is not None:                                              # Dec.py:173-181
        lens8 = torch.tensor([padded_len(int(n)) / 8 for n in acous_lens])
        key_mask = torch.arange(t_k).expand(B, t_k) >= lens8.unsqueeze(1)
    else:
        key_mask = None

    w_att = P[pre + 'acous_att.linear_att_w.weight']
    wk_hoisted = F.linear(acous_outputs, w_att) if hoist_keys else None

    tgt_chunk = emb_tgt[:, 0]                                               # Dec.py:199
    cell_value = acous_outputs.new_zeros(B, D)                              # Dec.py:200-201
    hc: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * cfg.num_unilstm_dec
    embs, logps, syms = [], [], []
    for idx in range(S_full - 1):                                           # Dec.py:205
        # --- forward_step (Dec.py:344-438)
        x = torch.cat([tgt_chunk, cell_value], dim=-1)                      # Dec.py:383
        new_hc = []
        for i in range(cfg.num_unilstm_dec):                                # Dec.py:393-419
            lp = f'{pre}dec.l{i}.'
            h0, c0 = hc[i] if hc[i] is not None else (x.new_zeros(B, D), x.new_zeros(B, D))
            h1, c1 = lstm_cell(x, h0, c0, P[lp + 'weight_ih_l0'], P[lp + 'weight_hh_l0'],
                               P[lp + 'bias_ih_l0'], P[lp + 'bias_hh_l0'])
            new_hc.append((h1, c1))
            out = h1
            if 0 < i < cfg.num_unilstm_dec - 1:                             # Dec.py:417-418
                out = out + x
            x = _drop(out, f'las.dec.s{idx}.l{i}')                          # Dec.py:403,419
        hc = new_hc
        dec_out = x
        # bilinear attention (attention.py:190-193): score = q . (W k_j)
        wk = wk_hoisted if hoist_keys else F.linear(acous_outputs, w_att)
        scores = torch.bmm(dec_out.unsqueeze(1), wk.transpose(1, 2))        # [B,1,Tk]
        if key_mask is not None:
            scores = scores.masked_fill(key_mask.unsqueeze(1), -1e12)       # attention.py:250-252
        probs = F.softmax(scores, dim=2)                                    # attention.py:268
        context = torch.bmm(probs, acous_outputs).squeeze(1)                # attention.py:273
        context = _drop(context, f'las.dec.s{idx}.att')                     # Dec.py:429
        ff_in = torch.cat([context, dec_out], dim=-1)                       # Dec.py:431
        cell_value = F.linear(ff_in, P[pre + 'acous_ffn.weight'])           # Dec.py:433
        logits = F.linear(cell_value, P[pre + 'acous_out.weight'], P[pre + 'acous_out.bias'])
        logp = F.log_softmax(logits, dim=1)                                 # Dec.py:436
        # --- decode (Dec.py:320-341)
        symbols = logp.topk(1)[1]                                           # [B,1]
        logps.append(logp)
        syms.append(symbols)
        ended = (symbols.view(-1) == EOS) | (symbols.view(-1) == PAD)
        for b in range(B):
            if lengths[b] > idx and bool(ended[b]):
                lengths[b] = len(syms)
        if teacher_forcing:
            tgt_chunk = emb_tgt[:, idx + 1]                                 # Dec.py:221
        else:
            tgt_chunk = F.embedding(symbols.view(-1), P[pre + 'embedder.weight'], padding_idx=PAD)
        embs.append(cell_value)                                             # Dec.py:224
    return (torch.stack(embs, 1), torch.stack(logps, 1), torch.stack(syms, 1), lengths)


# --------------------------------------------------------------------------------------------
# Transformer blocks (modules/layers.py)
# --------------------------------------------------------------------------------------------
def position_signal(max_len: int, d_model: int):
    """Sinusoid table (layers.py:292-309): sin on even columns, cos on odd columns."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def multihead_attention(P: Params, prefix: str, cfg: STConfig, q_in, k_in, v_in, mask):
    """MultiheadAttention.forward (layers.py:142-197) + ScaledDotProductAttention (layers.py:213-229)
    with both dropouts at p=0.  LayerNorm (eps 1e-6) is applied to the *query input only*; keys and
    values are projected from the raw input (layers.py:153-160).  mask: [B, 1|Lq, Lk], nonzero=keep."""
    h, dk = cfg.num_heads, cfg.d_k
    B, Lq, D = q_in.shape
    Lk = k_in.shape[1]
    residual = q_in
    qn = F.layer_norm(q_in, (D,), P[prefix + 'layer_norm.weight'], P[prefix + 'layer_norm.bias'], 1e-6)
    q = F.linear(qn, P[prefix + 'w_qs.weight']).view(B, Lq, h, dk).transpose(1, 2)
    k = F.linear(k_in, P[prefix + 'w_ks.weight']).view(B, Lk, h, dk).transpose(1, 2)
    v = F.linear(v_in, P[prefix + 'w_vs.weight']).view(B, Lk, h, dk).transpose(1, 2)
    attn = torch.matmul(q / (dk ** 0.5), k.transpose(2, 3))                 # layers.py:216
    if mask is not None:
        attn = attn.masked_fill(mask.unsqueeze(1) == 0, -1e9)               # layers.py:224
    attn = _drop(F.softmax(attn, dim=-1), prefix + 'attn')                  # layers.py:226 (hard-wired p = 0.1)
    o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, Lq, -1)
    out = _drop(F.linear(o, P[prefix + 'fc.weight']), prefix + 'fc') + residual   # layers.py:194-195
    return out, attn


def positionwise_ffn(P: Params, prefix: str, x):
    """PositionwiseFeedForward.forward (layers.py:243-252)."""
    D = x.shape[-1]
    y = F.layer_norm(x, (D,), P[prefix + 'layer_norm.weight'], P[prefix + 'layer_norm.bias'], 1e-6)
    y = F.linear(F.relu(F.linear(y, P[prefix + 'w_1.weight'], P[prefix + 'w_1.bias'])),
                 P[prefix + 'w_2.weight'], P[prefix + 'w_2.bias'])
    return _drop(y, prefix + 'ffn') + x                                     # layers.py:249-250


def tf_encoder(P: Params, cfg: STConfig, src_emb, src_mask):
    """Encoder.forward, 'standard' type (TFEnc.py:69-95): add the time signal once, N layers, LN 1e-6."""
    L, D = src_emb.shape[1], src_emb.shape[2]
    x = src_emb + position_signal(max(500, L), D)[:, :L].to(src_emb.dtype)
    att = None
    for i in range(cfg.enc_layers):
        p = f'enc_src.enc_layers.{i}.'
        x, att = multihead_attention(P, p + 'slf_attn.', cfg, x, x, x, src_mask)   # layers.py:59-61
        x = positionwise_ffn(P, p + 'pos_ffn.', x)
    x = F.layer_norm(x, (D,), P['enc_src.norm.weight'], P['enc_src.norm.bias'], 1e-6)
    return x, att


def tf_decoder(P: Params, cfg: STConfig, tgt_emb, memory, tgt_mask, src_mask):
    """Decoder.forward, 'standard' type (TFDec.py:66-136): final LayerNorm uses eps 1e-5 (TFDec.py:58)."""
    L, D = tgt_emb.shape[1], tgt_emb.shape[2]
    x = tgt_emb + position_signal(max(500, L), D)[:, :L].to(tgt_emb.dtype)
    for i in range(cfg.dec_layers):
        p = f'dec_tgt.dec_layers.{i}.'
        x, _ = multihead_attention(P, p + 'decslf_attn.', cfg, x, x, x, tgt_mask)  # layers.py:102-104
        x, _ = multihead_attention(P, p + 'encdec_attn.', cfg, x, memory, memory, src_mask)
        x = positionwise_ffn(P, p + 'pos_ffn.', x)
    x = F.layer_norm(x, (D,), P['dec_tgt.norm.weight'], P['dec_tgt.norm.bias'], 1e-5)
    return x


# --------------------------------------------------------------------------------------------
# Seq2seq glue (models/Seq2seq.py)
# --------------------------------------------------------------------------------------------
def pad_mask(seq):
    return (seq != PAD).unsqueeze(-2)                                       # layers.py:269-275


def subsequent_mask(n: int):
    return (1 - torch.triu(torch.ones((1, n, n)), diagonal=1)).bool()       # layers.py:278-289


def mix_embeddings(P: Params, src_ids, emb_dyn):
    """The embedding-passing mix, Seq2seq._get_src_emb (Seq2seq.py:183-199):
    enc_emb_proj(cat(E_static[src], e_dyn)), Linear(E+D -> D, no bias), dropout 0."""
    static = F.embedding(src_ids, P['enc_embedder.weight'], padding_idx=PAD)
    return F.linear(_drop(torch.cat([static, emb_dyn], dim=2), 'mix'), P['enc_emb_proj.weight'])   # Seq2seq.py:195


def target_embeddings(P: Params, cfg: STConfig, tgt):
    """Seq2seq._get_tgt_emb (Seq2seq.py:202-211)."""
    mask = pad_mask(tgt).to(torch.uint8) & subsequent_mask(tgt.size(-1)).to(torch.uint8)
    emb = _drop(F.embedding(tgt, P['dec_embedder.weight'], padding_idx=PAD), 'tgt_emb')   # Seq2seq.py:207-209
    if cfg.has_dec_emb_proj:
        emb = F.linear(emb, P['dec_emb_proj.weight'])
    return mask, emb


def length_mask(lengths: Sequence[int], max_len: int):
    """src_mask_input from the LAS lengths (Seq2seq.py:494-497)."""
    ln = torch.tensor(list(lengths), dtype=torch.int64)
    return (torch.arange(max_len).expand(len(ln), max_len) < ln.unsqueeze(1)).unsqueeze(1)


def translation_decoder(P: Params, cfg: STConfig, emb_tgt, enc_out, tgt_mask, src_mask, beam_width=1):
    """Seq2seq._decoder_de (Seq2seq.py:249-257)."""
    dec_out = tf_decoder(P, cfg, emb_tgt, enc_out, tgt_mask, src_mask)
    logits = F.linear(dec_out, P['out_tgt.weight'])
    logps = torch.log_softmax(logits, dim=2)
    scores, preds = logps.detach().topk(beam_width)
    return dec_out, logits, logps, preds, scores


def forward_train_st(P: Params, cfg: STConfig, src, tgt, acous_feats, acous_lens,
                     hoist_keys: bool = False, loops: bool = False, las_tgt=None):
    """Seq2seq.forward_train(mode='ST') (Seq2seq.py:468-507).  `las_tgt` [B, S+1] (BOS first): feed these tokens to the
    LAS decoder instead of its own arg-max (Dec.py:196-221 teacher forcing) -- with las_tgt = BOS | the free-running
    symbols this is the same computation, which is how tests pin the symbol path when comparing reduced precision."""
    tgt_mask, emb_tgt = target_embeddings(P, cfg, tgt)
    enc_ac = las_encoder(P, cfg, acous_feats, acous_lens, loops=loops)
    emb_dyn, logps_src, preds_src, lengths = las_decoder(P, cfg, enc_ac, acous_lens, tgt=las_tgt,
                                                         teacher_forcing=las_tgt is not None,
                                                         hoist_keys=hoist_keys)
    src_trim = src[:, 1:]                                                   # Seq2seq.py:214-219
    emb_src = mix_embeddings(P, src_trim, emb_dyn)
    src_mask_input = length_mask(lengths, emb_src.size(1))
    enc_out, _ = tf_encoder(P, cfg, emb_src, src_mask_input)
    _, _, logps_tgt, preds_tgt, _ = translation_decoder(P, cfg, emb_tgt, enc_out, tgt_mask,
                                                        src_mask_input)
    return {'emb_st': emb_src, 'preds_st': preds_tgt, 'logps_st': logps_tgt,
            'emb_dyn': emb_dyn, 'preds_asr': preds_src, 'lengths_asr': lengths,
            'logps_asr': logps_src, 'enc_out': enc_out}


def forward_train_asr(P: Params, cfg: STConfig, src, acous_feats, acous_lens):
    """Seq2seq.forward_train(mode='ASR') (Seq2seq.py:422-436): teacher-forced LAS.
    SpecAug (Enc.py:87-117) mutates the input with python `random`; callers that want parity with the
    reference pass already-augmented features, so it is not repeated here."""
    enc_ac = las_encoder(P, cfg, acous_feats, acous_lens)
    emb, logps, preds, lengths = las_decoder(P, cfg, enc_ac, acous_lens, tgt=src, teacher_forcing=True)
    return {'emb_asr': emb, 'logps_asr': logps, 'preds_asr': preds, 'lengths_asr': lengths}


def forward_train_mt(P: Params, cfg: STConfig, src, tgt, emb_dyn_ave):
    """Seq2seq.forward_train(mode='MT') (Seq2seq.py:438-466): static embedding + constant average
    dynamic embedding; src key-pad mask from the token ids."""
    tgt_mask, emb_tgt = target_embeddings(P, cfg, tgt)
    src_trim = src[:, 1:]
    dyn = emb_dyn_ave.to(acous_dtype(P)).repeat(src_trim.size(0), src_trim.size(1), 1)
    emb_src = mix_embeddings(P, src_trim, dyn)
    src_mask_input = pad_mask(src_trim)
    enc_out, _ = tf_encoder(P, cfg, emb_src, src_mask_input)
    _, _, logps_tgt, preds_tgt, _ = translation_decoder(P, cfg, emb_tgt, enc_out, tgt_mask,
                                                        src_mask_input)
    return {'emb_mt': emb_src, 'preds_mt': preds_tgt, 'logps_mt': logps_tgt}


def acous_dtype(P: Params):
    return P['enc_emb_proj.weight'].dtype


def masked_nll(logps, tgt, coeff: float = 1.0, n_minibatch: int = 1):
    """Loss of Trainer_ST._train_batch (trainer_st.py:268-288) with eval_with_mask and normalise_loss:
    sum over non-PAD targets of -logp[target], / #non-PAD, * coeff / n_minibatch.
    logps [B, L, V] (all positions; the last one is dropped here), tgt [B, L]."""
    lp = logps[:, :-1, :]
    target = tgt[:, 1:].reshape(-1)
    mask = tgt[:, 1:].ne(PAD).reshape(-1)
    per_tok = F.nll_loss(lp.reshape(-1, lp.size(-1)), target, reduction='none')   # loss.py:116-118
    total = per_tok.masked_select(mask).sum()                                     # loss.py:130-132
    return total / (1.0 * mask.sum()) * coeff / n_minibatch


def train_step_st(P: Params, cfg: STConfig, src, tgt, acous_feats, acous_lens, **kw):
    """forward_train('ST') + masked NLL; returns (loss, out_dict).  Call loss.backward() for grads."""
    out = forward_train_st(P, cfg, src, tgt, acous_feats, acous_lens, **kw)
    return masked_nll(out['logps_st'], tgt), out


# --------------------------------------------------------------------------------------------
# greedy / beam inference (Seq2seq.py:260-393, 512-638, 641-796)
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def forward_eval_st(P: Params, cfg: STConfig, acous_feats, acous_lens):
    """Seq2seq.forward_eval(mode='ST') (Seq2seq.py:593-636) — greedy, no KV cache, early stop when
    every row has produced EOS.  Returns preds [B, max_seq_len_tgt] int64 (BOS first, PAD tail)."""
    B = acous_feats.size(0)
    enc_ac = las_encoder(P, cfg, acous_feats, acous_lens)
    emb_dyn, _, preds_src, lengths = las_decoder(P, cfg, enc_ac, acous_lens)
    emb_src = mix_embeddings(P, preds_src.squeeze(2), emb_dyn)              # Seq2seq.py:608
    src_mask_input = length_mask(lengths, emb_src.size(1))
    enc_out, _ = tf_encoder(P, cfg, emb_src, src_mask_input)
    L = cfg.max_seq_len_tgt
    eos = torch.zeros(B, dtype=torch.bool)
    preds = torch.full((B, 1), BOS, dtype=torch.int64)
    for i in range(1, L):                                                   # Seq2seq.py:622
        tgt_mask, emb_tgt = target_embeddings(P, cfg, preds)
        _, _, _, pred, _ = translation_decoder(P, cfg, emb_tgt, enc_out, tgt_mask, src_mask_input)
        eos = eos | (pred[:, i - 1].squeeze(1) == EOS)                      # Seq2seq.py:284-285
        preds = torch.cat((preds, pred[:, i - 1]), dim=1)
        if int(eos.sum()) == B:                                             # Seq2seq.py:297-302
            if preds.size(1) != L:
                preds = torch.cat((preds, torch.full((B, L - preds.size(1)), PAD, dtype=torch.int64)), 1)
            break
    return {'preds_st': preds, 'preds_asr': preds_src, 'lengths_asr': lengths}


@torch.no_grad()
def forward_translate_st(P: Params, cfg: STConfig, acous_feats, acous_lens, beam_width: int = 1,
                         penalty_factor: float = 1.0, max_seq_len: int = 32):
    """Seq2seq.forward_translate(mode='ST') (Seq2seq.py:697-739) with the beam bookkeeping of
    _prep_translate/_step_translate (Seq2seq.py:307-393), including the final slice quirk
    `preds_expand.reshape(batch, -1)[:, :max_seq_len]` (Seq2seq.py:738)."""
    B = acous_feats.size(0)
    k = beam_width
    enc_ac = las_encoder(P, cfg, acous_feats, acous_lens)
    emb_dyn, _, preds_src, lengths = las_decoder(P, cfg, enc_ac, acous_lens)
    emb_src = mix_embeddings(P, preds_src.squeeze(2), emb_dyn)
    S = emb_src.size(1)
    src_mask_input = length_mask(lengths, S)
    enc_out, _ = tf_encoder(P, cfg, emb_src, src_mask_input)
    # _prep_translate (Seq2seq.py:307-334): a b c -> aaa bbb ccc
    eos_mask = torch.zeros(B * k, dtype=torch.bool)
    len_map = torch.ones(B * k)
    enc_exp = enc_out.repeat(1, k, 1).view(-1, S, cfg.dim_model)
    preds_exp = torch.full((B * k, 1), BOS, dtype=torch.int64)
    scores_exp = torch.zeros(B * k)
    mask_exp = src_mask_input.repeat(1, k, 1).view(-1, 1, S)
    for i in range(1, max_seq_len):                                         # Seq2seq.py:720
        tgt_mask, emb_tgt = target_embeddings(P, cfg, preds_exp)
        _, _, _, pred_all, score_all = translation_decoder(P, cfg, emb_tgt, enc_exp, tgt_mask,
                                                           mask_exp, beam_width=k)
        pred = pred_all[:, i - 1]                                           # [(B k), k]
        score = score_all[:, i - 1]
        if i == 1:                                                          # Seq2seq.py:349-356
            scores_exp = scores_exp + score.reshape(B, -1)[:, :k].contiguous().view(-1)
            pred_select = pred.reshape(B, -1)[:, :k].contiguous().view(-1)
            preds_exp = torch.cat((preds_exp, pred_select.unsqueeze(-1)), dim=1)
        else:                                                               # Seq2seq.py:358-380
            eos_exp = eos_mask.reshape(-1, 1).repeat(1, k)
            eos_exp[:, 0] = False
            score_temp = scores_exp.reshape(-1, 1) + score.masked_fill(
                eos_mask.reshape(-1, 1), 0).masked_fill(eos_exp, -1e9)
            score_temp = score_temp / (len_map.reshape(-1, 1) ** penalty_factor)
            score_select, pos = score_temp.reshape(B, -1).topk(k)
            scores_exp = score_select.view(-1) * (len_map.reshape(-1, 1) ** penalty_factor).view(-1)
            pos = (pos.float() + torch.arange(0, B * k * k, k * k).float().reshape(B, 1)).long()
            r_idx, c_idx = pos // k, pos % k
            pred_select = pred[r_idx, c_idx].view(-1)
            preds_exp[:, :i] = preds_exp[r_idx.view(-1), :i]
            preds_exp = torch.cat((preds_exp, pred_select.unsqueeze(-1)), dim=1)
        eos_mask = (pred_select == EOS) | eos_mask                          # Seq2seq.py:384-385
        len_map = len_map + torch.ones(B * k).masked_fill(eos_mask, 0)      # Seq2seq.py:386-387
        if int(eos_mask.sum()) == eos_mask.size(0):
            break
    return preds_exp.reshape(B, -1)[:, :max_seq_len].contiguous()           # Seq2seq.py:738


# --------------------------------------------------------------------------------------------
# synthetic inputs — shared by tests and bench so both arms see the same batch
# --------------------------------------------------------------------------------------------
def synthetic_batch(cfg: STConfig, batch: int, frames: int, seed: int = 333, ragged: bool = False,
                    tgt_len: Optional[int] = None):
    """fbank = randn(B, T_pad, F), T_pad = frames + 8 - frames % 8; token rows: BOS, ids in [5,V), EOS at a
    random position >= 8 (or the last slot), PAD after."""
    g = torch.Generator().manual_seed(seed)
    t_pad = padded_len(frames)
    if ragged:
        lens = [int(v) for v in torch.randint(max(1, frames // 2), frames + 1, (batch,), generator=g)]
        lens[0] = frames
    else:
        lens = [frames] * batch
    feats = torch.randn(batch, t_pad, cfg.acous_dim, generator=g)
    for b, n in enumerate(lens):                       # frames past the raw length are zero padding
        feats[b, n:] = 0.0
    def tokens(length, vocab):
        ids = torch.randint(5, vocab, (batch, length), generator=g)
        ids[:, 0] = BOS
        lo = min(8, length - 1)
        pos = torch.randint(lo, length, (batch,), generator=g)
        for b in range(batch):
            ids[b, pos[b]] = EOS
            ids[b, pos[b] + 1:] = PAD
        return ids
    src = tokens(cfg.max_seq_len_src, cfg.enc_vocab_size)
    tgt = tokens(tgt_len or cfg.max_seq_len_tgt, cfg.dec_vocab_size)
    return {'src': src, 'tgt': tgt, 'acous_feats': feats, 'acous_lens': lens}


# --------------------------------------------------------------------------------------------
# Input stage (utils/dataset.py:121-184), SURVEY.md 8 f-3
# --------------------------------------------------------------------------------------------
def load_acous_from_flis(flis, spkids=None, norm_path=None):
    """Dataset.load_mu_std + Dataset.load_acous_from_flis: `.npy` fbank matrices, optional per-speaker (x - mu) / std
    (statistics truncated to the feature width, dataset.py:168-171), zero padding to max_len + 8 - max_len % 8 frames
    through the reference's own trick of padding against a dummy of that length (dataset.py:178-182)."""
    import os
    import numpy as np
    feat_lis, max_len, cache = [], 0, {}
    for i, f in enumerate(flis):
        featarr = np.load(f)
        acous_dim = featarr.shape[1]
        if spkids is not None and norm_path is not None:
            spk = spkids[i]
            if spk not in cache:                                              # dataset.py:141-151
                cache[spk] = [np.load(os.path.join(norm_path, spk + '.mu.npy')),
                              np.load(os.path.join(norm_path, spk + '.std.npy'))]
            mu, std = cache[spk]
            if mu.shape[0] != acous_dim:                                      # dataset.py:168-171
                mu, std = mu[:acous_dim], std[:acous_dim]
            featarr = 1. * (featarr - mu) / std                               # dataset.py:173
        feat = torch.FloatTensor(featarr)
        max_len = max(max_len, feat.size(0))
        feat_lis.append(feat)
    divisible_eight = max_len + 8 - max_len % 8                               # dataset.py:179
    feat_lis.append(torch.ones(divisible_eight, acous_dim))
    return torch.nn.utils.rnn.pad_sequence(feat_lis, batch_first=True)[:-1]
