"""pytest configuration: markers, sys.path for the oracle and the product package, golden loader."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['st_tiny_ragged', 'st_tiny_aligned', 'st_small']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz fixture written by oracle/make_golden.py (outputs of the real reference)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))

    def group(self, prefix, as_torch=True):
        out = {}
        for k in self.z.files:
            if k.startswith(prefix + '/'):
                v = self.z[k]
                out[k[len(prefix) + 1:]] = torch.from_numpy(np.array(v)) if as_torch else v
        return out

    def __getitem__(self, k):
        return torch.from_numpy(np.array(self.z[k]))

    @property
    def cfg(self):
        from oracle.st_oracle import STConfig
        c = {k: int(v) for k, v in self.group('cfg', as_torch=False).items()}
        return STConfig(enc_vocab_size=c['V'], dec_vocab_size=c['V'], enc_embedding_size=c['E'],
                        dec_embedding_size=c['E'], max_seq_len_src=c['S'], max_seq_len_tgt=c['L'],
                        num_heads=c['heads'], dim_model=c['dim_model'], dim_feedforward=c['FF'],
                        enc_layers=c['layers'], dec_layers=c['layers'], acous_dim=c['F'],
                        acous_hidden_size=c['H'])

    def params(self, requires_grad=False, dtype=torch.float32):
        p = {k: v.to(dtype).clone() for k, v in self.group('param').items()}
        if requires_grad:
            for v in p.values():
                v.requires_grad_(True)
        return p

    def inputs(self):
        override = os.environ.get('B200ST_TEST_INPUTS')
        if override and os.path.exists(override):
            return torch.load(override)
        i = self.group('in')
        i['acous_lens'] = [int(v) for v in i['acous_lens']]
        return i


class SeededGolden(Golden):
    """A fixture whose weights are not stored: they are `oracle.st_oracle.init_params(cfg, seed, scale)` (seeded CPU
    generator; make_golden.case_st_seeded loaded exactly those into the real reference).  Gradients are stored as a norm
    plus a strided sample of <= 4096 entries per parameter."""

    def params(self, requires_grad=False, dtype=torch.float32):
        from oracle.st_oracle import init_params
        P = init_params(self.cfg, seed=int(self.z['seed']), scale=float(self.z['wscale']), dtype=dtype)
        for k, v in self.group('param_abssum').items():       # the seeded generator must reproduce the fixture's weights
            got = float(P[k].double().abs().sum())
            assert abs(got - float(v)) <= 1e-9 * max(1.0, float(v)), f'seeded init drifted for {k}: regenerate the fixture'
        if requires_grad:
            for v in P.values():
                v.requires_grad_(True)
        return P

    def grad_check(self, named_grads, tol, per_param_factor=1.0):
        """Every stored parameter: |norm - ref| and the error on the stored sample, both relative to the parameter's
        reference norm floored at 1e-3 of the global norm (the metric of test_gpu_parity._grad_check), must stay within
        per_param_factor * tol; the error over ALL stored samples together within tol of their norm.  Returns
        (worst per-parameter error / tol, global sample error)."""
        norms = self.group('st_gradnorm')
        gn = sum(float(v) ** 2 for v in norms.values()) ** 0.5
        worst = num = den = 0.0
        for name, ref_norm in norms.items():
            g = named_grads[name]
            assert g is not None, name
            g = g.detach().double().cpu().reshape(-1)
            sample = self['st_gradsample/' + name].double()
            stride = max(1, g.numel() // 4096)
            got = g[::stride][:4096]
            frac = (sample.numel() / g.numel()) ** 0.5          # a sample carries ~sqrt(fraction) of the norm
            bound = tol * max(float(ref_norm), 1e-3 * gn)
            e1 = abs(float(g.norm()) - float(ref_norm)) / bound
            e2 = float((got - sample).norm()) / max(bound * frac, 1e-30)
            num += float((got - sample).norm()) ** 2
            den += float(sample.norm()) ** 2
            worst = max(worst, e1, e2 / 2)                       # sampling noise: allow 2x on the sample
            assert e1 <= per_param_factor and e2 <= 2 * per_param_factor, (name, e1, e2, float(ref_norm), gn)
        glob = (num / max(den, 1e-300)) ** 0.5
        assert glob <= tol, glob
        return worst, glob


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


def rel_err(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
