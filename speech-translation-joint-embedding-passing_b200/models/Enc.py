"""Pyramidal BLSTM acoustic encoder on the b200st persistent recurrence kernels.

Mirror of the reference's models/Enc.py (class Enc): same constructor, same `forward(acous_feats,
acous_lens, is_training, hidden, use_gpu)`, same parameter names (the four torch.nn.LSTM modules are kept
as parameter containers: `acous_enc_l{1..4}.weight_ih_l0[_reverse]` ...), same semantics:
  * lengths are padded with the `n + 8 - n % 8` rule (Enc.py:142),
  * each layer is a packed bidirectional LSTM: zero state, outputs beyond the length are zero and the
    reverse direction starts at each sequence's own end (Enc.py:150-157),
  * consecutive frame pairs are concatenated between layers (Enc.py:166-167) — here folded into the
    recurrence kernel's store, so no reshape/copy happens,
  * SpecAug (Enc.py:87-117) draws from python `random` and zeroes the INPUT tensor in place.
Internally activations are time-major [T, B, *] so that each time step's rows are contiguous.
"""
import random

import torch
import torch.nn as nn

from b200st import functional as BF
from b200st import runtime as rt
from b200st.kernels import K
from b200st.hostutil import check_device


def padded_lengths(acous_lens, batch_size, acous_len, device):
    """Enc.py:139-142 / Dec.py:175.  Returns (int32 device tensor of padded lengths, host list or None)."""
    if acous_lens is None:
        host = [int(acous_len)] * batch_size
    elif torch.is_tensor(acous_lens) and acous_lens.is_cuda:
        ln = acous_lens.reshape(-1).to(torch.int32)                # device-resident raw lengths (graph mode)
        return ln + 8 - ln % 8, None
    else:
        host = [int(e) + 8 - int(e) % 8 for e in acous_lens]
    return torch.tensor(host, dtype=torch.int32).to(device, non_blocking=True), host


class Enc(nn.Module):

    def __init__(self, acous_dim=26, acous_hidden_size=256, acous_norm=False, spec_aug=False,
                 batch_norm=False, enc_mode='pyramid', dropout=0.0, batch_first=True):
        super().__init__()
        if enc_mode != 'pyramid' or batch_norm or not batch_first:
            raise NotImplementedError("b200st Enc implements enc_mode='pyramid', batch_norm=False, "
                                      "batch_first=True (what Seq2seq constructs, Seq2seq.py:155-158)")
        self.acous_dim = acous_dim
        self.acous_hidden_size = acous_hidden_size
        self.acous_norm = acous_norm
        self.spec_aug = spec_aug
        self.batch_norm = batch_norm
        self.enc_mode = enc_mode
        self.dropout = nn.Dropout(dropout)
        h = acous_hidden_size
        mk = lambda i: torch.nn.LSTM(i, h, num_layers=1, batch_first=batch_first, bias=True,
                                     dropout=dropout, bidirectional=True)
        self.acous_enc_l1 = mk(acous_dim)       # Enc.py:50-66
        self.acous_enc_l2 = mk(h * 4)
        self.acous_enc_l3 = mk(h * 4)
        self.acous_enc_l4 = mk(h * 4)

    def check_var(self, var_name, var_val_set=None):
        if not hasattr(self, var_name):
            setattr(self, var_name, var_val_set if var_val_set is not None else None)

    def pre_process_acous(self, acous_feats):
        """SpecAug, Enc.py:87-117 (same draw order: t, f, t0, f0, twice; in place)."""
        self.check_var('spec_aug', False)
        if not self.spec_aug:
            return acous_feats
        max_time, max_channel = acous_feats.size(1), acous_feats.size(2)
        const_t = int(min(40, 0.2 * max_time))
        for _ in range(2):
            t = random.randint(0, const_t)
            f = random.randint(0, 7)
            t0 = random.randint(0, max_time - t - 1)
            f0 = random.randint(0, max_channel - f - 1)
            acous_feats[:, t0:t0 + t, :] = 0
            acous_feats[:, :, f0:f0 + f] = 0
        return acous_feats

    @staticmethod
    def _dir_weights(lstm, reverse):
        sfx = '_reverse' if reverse else ''
        return tuple(getattr(lstm, n + sfx) for n in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0'))

    def forward(self, acous_feats, acous_lens=None, is_training=False, hidden=None, use_gpu=False,
                lens_dev=None):
        batch_size, acous_len = acous_feats.size(0), acous_feats.size(1)
        assert acous_len % 8 == 0, 'feature length must be a multiple of 8 (trainer_st.py:252 pads it)'
        if is_training:
            acous_feats = self.pre_process_acous(acous_feats)
        if lens_dev is None:
            lens_dev, host = padded_lengths(acous_lens, batch_size, acous_len, acous_feats.device)
            if host is not None:   # Enc.py:159-160 reshapes to the full length: max(lens) must equal it
                assert max(host) == acous_len, 'padded max length must equal the feature length'
        # [B, T, F] -> time-major [T, B, F] in the compute dtype (one pass)
        x = K().transpose01(acous_feats.contiguous(), out_dtype=rt.compute_dtype())
        lens = lens_dev
        layers = (self.acous_enc_l1, self.acous_enc_l2, self.acous_enc_l3, self.acous_enc_l4)
        for li, lstm in enumerate(layers):
            last = li == len(layers) - 1
            x = BF.blstm_layer(x, lens, self._dir_weights(lstm, False), self._dir_weights(lstm, True),
                               pair=1 if last else 2, batch_first_out=last)
            # Enc.py:159,178,195,212: dropout on the layer output (element-wise, so it commutes with the frame-pair
            # concat that the recurrence kernel has already folded into its store addresses)
            x = BF.dropout(x, float(self.dropout.p), self.training, tag=f'las.enc.l{li + 1}')
            if not last:
                lens = lens // 2                                   # Enc.py:170,187,204
        return x                                                   # [B, T/8, 2H]
