"""Input stage: `.npy` fbank batches -> per-speaker normalised, zero-padded [B, T_pad, F].

Chain of evidence: the REAL reference method (`Dataset.load_acous_from_flis` + `load_mu_std`, utils/dataset.py:134-184,
imported from /root/reference when it is mounted — build container only) == the oracle restatement == the host loader +
device kernel (bit-exact with float32 statistics)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from b200st import kernels
from b200st.data import FbankPrefetcher, fbank_to_device, load_fbank_batch, padded_frames
from fake_kernels import FakeKernels
from oracle import st_oracle as O

REF = os.environ.get('ST_REFERENCE', '/root/reference')


def _write(tmp, lens, F=12, stat_dim=None, stat_dtype=np.float32, seed=0):
    rng = np.random.default_rng(seed)
    flis, spkids = [], []
    norm = os.path.join(tmp, 'norm')
    os.makedirs(norm, exist_ok=True)
    spk_names = ['spkA', 'spkB', 'spkC']
    for s in spk_names:
        d = stat_dim or F
        np.save(os.path.join(norm, s + '.mu.npy'), rng.normal(size=d).astype(stat_dtype))
        np.save(os.path.join(norm, s + '.std.npy'), (0.5 + rng.random(d)).astype(stat_dtype))
    for i, n in enumerate(lens):
        f = os.path.join(tmp, f'utt{i}.npy')
        np.save(f, (rng.normal(size=(n, F)) * 3 + 1).astype(np.float32))
        flis.append(f)
        spkids.append(spk_names[i % 3])
    return flis, spkids, norm


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'utils')), reason='reference tree not mounted (GPU box)')
def test_oracle_equals_reference_dataset_methods(tmp_path):
    """Pin the oracle restatement with the unmodified reference code (import shims of SURVEY.md 8c only)."""
    code = r'''
import sys, types, json, numpy as np, torch
sys.dont_write_bytecode = True
sys.path.insert(0, %r)
for name in ['bpemb', 'matplotlib', 'matplotlib.pyplot', 'torchtext']:
    sys.modules[name] = types.ModuleType(name)
sys.modules['bpemb'].BPEmb = object
sys.modules['matplotlib'].use = lambda *a, **k: None
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
from utils.dataset import IterDataset
spec = json.load(open(sys.argv[1]))
stub = types.SimpleNamespace(batches=[{'acous_flis': spec['flis'], 'acous_spkids': spec['spkids']}],
                             acous_norm_path=spec['norm'], acous_norm=spec['use_norm'])
stub.load_mu_std = lambda i: IterDataset.load_mu_std(stub, i)
stub.load_acous_from_flis = lambda i, norm_param=None: IterDataset.load_acous_from_flis(stub, i, norm_param=norm_param)
out = IterDataset.load_file(stub, 0)
np.save(sys.argv[2], out.numpy())
''' % REF
    import json
    import subprocess
    for use_norm, stat_dim, stat_dtype in ((True, None, np.float32), (True, 15, np.float64), (False, None, np.float32)):
        tmp = str(tmp_path / f'c{int(use_norm)}{stat_dim}')
        os.makedirs(tmp)
        flis, spkids, norm = _write(tmp, [37, 64, 5, 50], stat_dim=stat_dim, stat_dtype=stat_dtype)
        spec = os.path.join(tmp, 'spec.json')
        json.dump({'flis': flis, 'spkids': spkids, 'norm': norm, 'use_norm': use_norm}, open(spec, 'w'))
        script = os.path.join(tmp, 'run_ref.py')
        open(script, 'w').write(code)
        res = subprocess.run([sys.executable, script, spec, os.path.join(tmp, 'ref.npy')], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        ref = torch.from_numpy(np.load(os.path.join(tmp, 'ref.npy')))
        got = O.load_acous_from_flis(flis, spkids if use_norm else None, norm if use_norm else None)
        assert got.shape == ref.shape == (4, 72, 12)
        assert torch.equal(got, ref)


@pytest.mark.parametrize('lens', [[37, 64, 5, 50], [8], [16, 16], [1, 200, 199]])
@pytest.mark.parametrize('use_norm', [True, False])
def test_host_loader_and_fake_kernel_equal_oracle(tmp_path, lens, use_norm):
    flis, spkids, norm = _write(str(tmp_path), lens, stat_dim=13)          # statistics carry an extra (energy) term
    old = kernels.set_backend(FakeKernels())
    try:
        b = load_fbank_batch(flis, spkids if use_norm else None, norm if use_norm else None, pin=False)
        assert b['T_pad'] == padded_frames(max(lens)) and b['T_pad'] % 8 == 0 and b['T_pad'] > max(lens)
        assert b['packed'].shape == (sum(lens), 12) and b['offsets'].tolist() == np.concatenate([[0], np.cumsum(lens)[:-1]]).tolist()
        got = fbank_to_device(b, 'cpu')
        ref = O.load_acous_from_flis(flis, spkids if use_norm else None, norm if use_norm else None)
        assert torch.equal(got, ref)
    finally:
        kernels.set_backend(old)


def test_prefetcher_yields_every_batch_in_order_cpu(tmp_path):
    old = kernels.set_backend(FakeKernels())
    try:
        batches = []
        for j, lens in enumerate(([20, 31], [8, 9, 10], [64])):
            d = str(tmp_path / f'b{j}')
            os.makedirs(d)
            flis, spk, norm = _write(d, lens, seed=j)
            batches.append((flis, spk))
        # every batch directory has its own statistics with identical names: use the last one for all (any fixed set works)
        out = list(FbankPrefetcher(batches, 'cpu', norm_path=norm))
        assert len(out) == 3
        for (feats, lens), (flis, spk) in zip(out, batches):
            assert torch.equal(feats, O.load_acous_from_flis(flis, spk, norm))
            assert lens == [np.load(f).shape[0] for f in flis]
    finally:
        kernels.set_backend(old)


@pytest.mark.gpu
@pytest.mark.parametrize('lens', [[37, 64, 5, 50], [1000] * 16 + [517, 3], [8]])
@pytest.mark.parametrize('use_norm', [True, False])
def test_device_stage_equals_oracle_bit_exact(tmp_path, lens, use_norm):
    flis, spkids, norm = _write(str(tmp_path), lens, F=80, stat_dim=81)
    b = load_fbank_batch(flis, spkids if use_norm else None, norm if use_norm else None)
    got = fbank_to_device(b, torch.device('cuda'))
    ref = O.load_acous_from_flis(flis, spkids if use_norm else None, norm if use_norm else None)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), ref)


@pytest.mark.gpu
def test_prefetcher_gpu(tmp_path):
    batches = []
    for j, lens in enumerate(([200, 311], [80, 90, 100], [640], [33, 34])):
        d = str(tmp_path / f'b{j}')
        os.makedirs(d)
        flis, spk, norm = _write(d, lens, F=40, seed=j)
        batches.append((flis, spk))
    for (feats, lens), (flis, spk) in zip(FbankPrefetcher(batches, 'cuda', norm_path=norm), batches):
        assert feats.is_cuda and torch.equal(feats.cpu(), O.load_acous_from_flis(flis, spk, norm))
