"""Data-parallel path on CPU: world_size 2 over gloo (host logic only; kernels replaced by the torch fake).

Oracle: N ranks x per-rank loss normalisation + MEAN all-reduce of gradients must equal the
reference's serial accumulation `batch_size = N*b, minibatch_partition = N` (trainer_st.py:225-290), which is
reproduced here by running Trainer_ST with minibatch_partition=2 in one process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import Golden


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _items(I, sl=None):
    sl = sl or slice(None)
    return {'srcid': [I['src'][sl]], 'tgtid': [I['tgt'][sl]], 'acous_feat': [I['acous_feats'][sl]],
            'acouslen': I['acous_lens'][sl], 'srclen': None, 'tgtlen': None}


def _worker(rank, world, port, name, out, part=1):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), 'speech-translation-joint-embedding-passing_b200'), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from b200st import kernels
    from fake_kernels import FakeKernels
    from b200st.dp import GradAllReducer
    from helpers import build_model
    from b200st.train_step import Trainer_ST
    kernels.set_backend(FakeKernels())
    g = Golden(name)
    m = build_model(g.cfg, g.params())
    m.train()
    I = g.inputs()
    B = I['src'].size(0)
    half = B // 2
    sl = slice(0, half) if rank == 0 else slice(half, 2 * half)
    red = GradAllReducer(m, bucket_bytes=64 << 10)           # small buckets: several async all-reduces in flight
    tr = Trainer_ST(use_gpu=False, batch_size=half, minibatch_partition=part, reducer=red, fused_loss=(part == 1))
    n_hooked = [0]
    orig_flush = red._flush
    def counting_flush():
        n_hooked[0] += len(red._pending)
        orig_flush()
    red._flush = counting_flush
    # both halves must be padded to the same feature length the single-process run used
    items = _items(I, sl)
    tr._train_batch(m, items)
    if rank == 0:
        n_grads = sum(1 for p in m.parameters() if p.grad is not None)
        assert n_hooked[0] == n_grads, (n_hooked[0], n_grads)      # every gradient reduced exactly once
        torch.save({n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['st_small'])
def test_dp2_equals_minibatch_partition(tmp_path, name):
    from b200st import kernels
    from fake_kernels import FakeKernels
    from helpers import build_model
    from b200st.train_step import Trainer_ST
    g = Golden(name)
    I = g.inputs()
    B = (I['src'].size(0) // 2) * 2
    # make every minibatch share one padded feature length (as the DP ranks do)
    I = dict(I, acous_lens=I['acous_lens'][:B], src=I['src'][:B], tgt=I['tgt'][:B], acous_feats=I['acous_feats'][:B])
    maxlen = max(I['acous_lens'])
    halves = [I['acous_lens'][:B // 2], I['acous_lens'][B // 2:]]
    if any(max(h) + 8 - max(h) % 8 != maxlen + 8 - maxlen % 8 for h in halves):
        I['acous_lens'] = list(I['acous_lens'])
        I['acous_lens'][0] = maxlen
        I['acous_lens'][B // 2] = maxlen
    out = str(tmp_path / 'dp_grads.pt')
    port = _free_port()
    # serial reference semantics in this process
    old = kernels.set_backend(FakeKernels())
    try:
        m = build_model(g.cfg, g.params())
        m.train()
        tr = Trainer_ST(use_gpu=False, batch_size=B, minibatch_partition=2)
        tr._train_batch(m, _items(I))
        ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    finally:
        kernels.set_backend(old)
    torch.save(I, str(tmp_path / 'inputs.pt'))
    # the workers re-load the golden; patch its inputs through an env-provided file
    os.environ['B200ST_TEST_INPUTS'] = str(tmp_path / 'inputs.pt')
    try:
        mp.spawn(_worker, args=(2, port, name, out), nprocs=2, join=True)
    finally:
        os.environ.pop('B200ST_TEST_INPUTS', None)
    got = torch.load(out)
    assert set(got) == set(ref)
    for n in ref:
        denom = float(ref[n].norm()) + 1e-12
        assert float((got[n] - ref[n]).norm()) / denom < 1e-5, n


def test_dp2_with_minibatch_partition2_equals_partition4(tmp_path):
    """Each rank accumulates over TWO slices (minibatch_partition=2) and the reducer is armed for the last slice only, so the
    accumulated gradient is reduced once: 2 ranks x 2 slices == the reference's batch 4, minibatch_partition=4."""
    from b200st import kernels
    from fake_kernels import FakeKernels
    from helpers import build_model
    from b200st.train_step import Trainer_ST
    name = 'st_small'
    g = Golden(name)
    I = g.inputs()
    I = dict(I, acous_lens=I['acous_lens'][:4], src=I['src'][:4], tgt=I['tgt'][:4], acous_feats=I['acous_feats'][:4])
    T = I['acous_feats'].size(1)
    I['acous_lens'] = [T - 8 if T % 8 == 0 else T - T % 8] * 4      # every single-utterance slice pads to the same T
    I['acous_lens'] = [min(n, T - 1) for n in I['acous_lens']]
    assert all(n + 8 - n % 8 == T for n in I['acous_lens'])
    out = str(tmp_path / 'dp_grads.pt')
    old = kernels.set_backend(FakeKernels())
    try:
        m = build_model(g.cfg, g.params())
        m.train()
        Trainer_ST(use_gpu=False, batch_size=4, minibatch_partition=4, fused_loss=False)._train_batch(m, _items(I))
        ref = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    finally:
        kernels.set_backend(old)
    torch.save(I, str(tmp_path / 'inputs.pt'))
    os.environ['B200ST_TEST_INPUTS'] = str(tmp_path / 'inputs.pt')
    try:
        mp.spawn(_worker, args=(2, _free_port(), name, out, 2), nprocs=2, join=True)
    finally:
        os.environ.pop('B200ST_TEST_INPUTS', None)
    got = torch.load(out)
    assert set(got) == set(ref)
    for n in ref:
        assert float((got[n] - ref[n]).norm()) / (float(ref[n].norm()) + 1e-12) < 1e-5, n
