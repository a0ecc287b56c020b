"""Input stage for the acoustic features: what `Dataset.load_file` / `load_mu_std` /
`load_acous_from_flis` do per batch (utils/dataset.py:121-184) — read `<utt>.npy` fbank matrices, apply the
per-speaker mean / std normalisation (`<spk>.mu.npy`, `<spk>.std.npy`), pad the batch with zeros to
`max_len + 8 - max_len % 8` frames — split between host and device the B200 way:

  host  : `load_fbank_batch()` reads the files and lays the utterances BACK TO BACK in one pinned buffer (no padding
          crosses PCIe) together with offsets, lengths and one row of speaker statistics per utterance;
  device: `fbank_to_device()` issues the H2D copies (asynchronous, on the caller's stream) and ONE kernel
          (`b200st_fbank_norm_pad`) that normalises and writes the zero-padded [B, T_pad, F] batch the model consumes.

`FbankPrefetcher` runs the host half for batch i+1 on a worker thread and its H2D + kernel on a side stream while the
model trains on batch i (the reference's DataLoader has num_workers=0, dataset.py:605-607).
"""
from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .kernels import K


def padded_frames(max_len: int) -> int:
    return max_len + 8 - max_len % 8                                   # dataset.py:179 (adds 8 when already a multiple)


def load_fbank_batch(flis: Sequence[str], spkids: Optional[Sequence[str]] = None, norm_path: Optional[str] = None,
                     pin: bool = True) -> Dict:
    """Host half.  flis: one `.npy` [len, F] per utterance; spkids + norm_path: speaker statistics (dataset.py:134-153),
    omitted when acous_norm is off.  Statistics longer than F are truncated ("get rid of training energy term",
    dataset.py:168-171)."""
    arrs = [np.load(f) for f in flis]
    F = arrs[0].shape[1]
    lens = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    offsets = np.zeros(len(arrs), dtype=np.int64)
    offsets[1:] = np.cumsum(lens[:-1], dtype=np.int64)
    total = int(lens.sum())
    mk = (lambda *s, dtype: torch.empty(s, dtype=dtype).pin_memory()) if (pin and torch.cuda.is_available()) else \
         (lambda *s, dtype: torch.empty(s, dtype=dtype))
    packed = mk(total, F, dtype=torch.float32)
    pk = packed.numpy()
    for a, o, n in zip(arrs, offsets, lens):
        pk[o:o + n] = a                                                 # (casts to fp32 like torch.FloatTensor(featarr))
    out = {'packed': packed, 'offsets': torch.from_numpy(offsets), 'lens': torch.from_numpy(lens), 'mu': None, 'sd': None,
           'T_pad': padded_frames(int(lens.max())), 'F': F, 'acous_lens': [int(n) for n in lens]}
    if spkids is not None and norm_path is not None:
        cache = {}
        mu = mk(len(arrs), F, dtype=torch.float32)
        sd = mk(len(arrs), F, dtype=torch.float32)
        for i, spk in enumerate(spkids):
            if spk not in cache:
                cache[spk] = (np.load(os.path.join(norm_path, spk + '.mu.npy'))[:F],
                              np.load(os.path.join(norm_path, spk + '.std.npy'))[:F])
            mu.numpy()[i], sd.numpy()[i] = cache[spk]
        out['mu'], out['sd'] = mu, sd
    return out


def fbank_to_device(batch: Dict, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device half: H2D of the packed utterances (+ statistics) and the normalise-and-pad kernel, all asynchronous on the
    current stream.  Returns acous_feats fp32 [B, T_pad, F]."""
    dev = lambda t: None if t is None else t.to(device, non_blocking=True)
    return K().fbank_norm_pad(dev(batch['packed']), dev(batch['offsets']), dev(batch['lens']), dev(batch['mu']),
                              dev(batch['sd']), batch['T_pad'], out=out)


class FbankPrefetcher:
    """Iterates over batches of (flis, spkids): the file reads of batch i+1 run on a worker thread and its transfer +
    normalisation on a side stream while the caller consumes batch i.  Yields (acous_feats [B, T_pad, F] on the device,
    acous_lens list) — what `forward_train(acous_feats=..., acous_lens=...)` takes."""

    def __init__(self, batches: Sequence, device, norm_path: Optional[str] = None):
        self.batches, self.device, self.norm_path = list(batches), torch.device(device), norm_path
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == 'cuda' else None

    def _host(self, i, slot):
        flis, spk = self.batches[i]
        slot['host'] = load_fbank_batch(flis, spk, self.norm_path if spk is not None else None)

    def _device(self, slot):
        if self.stream is None:
            slot['feats'] = fbank_to_device(slot['host'], self.device)
            return
        with torch.cuda.stream(self.stream):
            slot['feats'] = fbank_to_device(slot['host'], self.device)
            slot['ready'] = torch.cuda.Event()
            slot['ready'].record(self.stream)

    def __iter__(self):
        n = len(self.batches)
        if n == 0:
            return
        nxt = {}
        self._host(0, nxt)
        self._device(nxt)
        for i in range(n):
            cur, nxt = nxt, {}
            worker = None
            if i + 1 < n:
                worker = threading.Thread(target=self._host, args=(i + 1, nxt))
                worker.start()
            if 'ready' in cur:
                torch.cuda.current_stream().wait_event(cur['ready'])
                cur['feats'].record_stream(torch.cuda.current_stream())
            yield cur['feats'], cur['host']['acous_lens']
            if worker is not None:
                worker.join()
                self._device(nxt)
