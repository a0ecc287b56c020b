"""Fused clip + Adam against the calls it replaces: torch.nn.utils.clip_grad_norm_ followed by
torch.optim.Adam.step() behind Optimizer.step() (reference modules/optim.py:31-36, trainer_base.py:422-426)."""
import copy

import pytest
import torch

from b200st import kernels
from conftest import rel_err
from fake_kernels import FakeKernels

SHAPES = [(300, 64), (64,), (1000, 200), (7,), (33, 5, 3), (8192,), (8193,), (1,), (513, 1024)]


def _make(device, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.3).to(device)) for s in SHAPES]


def _grads(params, it, scale):
    g = torch.Generator().manual_seed(100 + it)
    for i, p in enumerate(params):
        if i == 3:            # a parameter the loss never reaches (template layers, acous_out in ST mode): skipped
            continue
        p.grad = (torch.randn(p.shape, generator=g) * scale).to(p.device)


def _run_pair(device, max_norm, scale, wd, steps=4):
    from modules.optim import Optimizer, lr_scheduler
    ours, ref = _make(device), _make(device)
    ours[3].requires_grad_(True)
    opt = Optimizer(torch.optim.Adam(ours, lr=1e-3, weight_decay=wd), max_grad_norm=max_norm)
    adam_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
    for it in range(steps):
        lr_scheduler(opt.optimizer, it + 1, init_lr=1e-4, peak_lr=2e-3, warmup_steps=2)
        lr_scheduler(adam_ref, it + 1, init_lr=1e-4, peak_lr=2e-3, warmup_steps=2)
        _grads(ours, it, scale)
        _grads(ref, it, scale)
        opt.step()
        if max_norm > 0:
            torch.nn.utils.clip_grad_norm_([p for p in ref if p.grad is not None], max_norm)
        adam_ref.step()
    return ours, ref, opt, adam_ref


def _check(ours, ref, opt, adam_ref, tol):
    for i, (a, b) in enumerate(zip(ours, ref)):
        assert rel_err(a, b) < tol, (i, rel_err(a, b))
        sa, sb = opt.optimizer.state.get(a), adam_ref.state.get(b)
        assert bool(sa) == bool(sb)
        if sa:
            for k in ('exp_avg', 'exp_avg_sq'):
                assert rel_err(sa[k], sb[k]) < tol, (i, k, rel_err(sa[k], sb[k]))


@pytest.mark.parametrize('max_norm,scale,wd', [(1.0, 1.0, 0.0), (1.0, 1e-4, 0.0), (0, 1.0, 0.0), (5.0, 0.5, 0.01)])
def test_host_logic_cpu(max_norm, scale, wd):
    """Host side (state layout, pointer-table bookkeeping, lr plumbing) with the torch stand-in for the kernels."""
    old = kernels.set_backend(FakeKernels())
    try:
        ours, ref, opt, adam_ref = _run_pair('cpu', max_norm, scale, wd)
        _check(ours, ref, opt, adam_ref, 1e-5)
    finally:
        kernels.set_backend(old)


def test_state_dict_layout_cpu():
    old = kernels.set_backend(FakeKernels())
    try:
        ours, ref, opt, adam_ref = _run_pair('cpu', 1.0, 1.0, 0.0, steps=2)
        sd, sd_ref = opt.optimizer.state_dict(), adam_ref.state_dict()
        assert sd['param_groups'][0].keys() == sd_ref['param_groups'][0].keys()
        assert set(sd['state'][0].keys()) == set(sd_ref['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}
        assert float(sd['state'][0]['step']) == float(sd_ref['state'][0]['step']) == 2.0
        # a fresh torch Adam accepts the state (checkpoint compatibility, checkpoint.py:76 / trainer_base.py:197-200)
        fresh = torch.optim.Adam(_make('cpu'), lr=1e-3)
        fresh.load_state_dict(copy.deepcopy(sd))
    finally:
        kernels.set_backend(old)


@pytest.mark.gpu
@pytest.mark.parametrize('max_norm,scale,wd', [(1.0, 1.0, 0.0), (1.0, 1e-4, 0.0), (0, 1.0, 0.0), (5.0, 0.5, 0.01)])
def test_fused_clip_adam_matches_torch(max_norm, scale, wd):
    ours, ref, opt, adam_ref = _run_pair('cuda', max_norm, scale, wd, steps=5)
    _check(ours, ref, opt, adam_ref, 1e-5)
    if max_norm > 0:
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ours if p.grad is not None)).item()
        assert abs(float(opt._fused.grad_norm) - gn) < 1e-5 * gn


@pytest.mark.gpu
def test_fused_clip_adam_in_cuda_graph():
    """Captured once, replayed: the device step counter and the device learning rate drive the bias corrections."""
    from modules.optim import Optimizer
    ours, ref = _make('cuda'), _make('cuda')
    opt = Optimizer(torch.optim.Adam(ours, lr=1e-3), max_grad_norm=1.0)
    adam_ref = torch.optim.Adam(ref, lr=1e-3)
    gbuf = [torch.zeros_like(p) for p in ours]
    for p, g in zip(ours, gbuf):
        p.grad = g
    opt._engine().prepare()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph):
            opt.step()
    torch.cuda.current_stream().wait_stream(s)
    for it in range(4):
        lr = 1e-3 * (it + 1)
        gen = torch.Generator().manual_seed(7 + it)
        for p, g, q in zip(ours, gbuf, ref):
            new = torch.randn(p.shape, generator=gen).cuda()
            g.copy_(new)
            q.grad = new.clone()
        opt.set_lr(lr)
        graph.replay()
        for grp in adam_ref.param_groups:
            grp['lr'] = lr
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        adam_ref.step()
    torch.cuda.synchronize()
    for a, b in zip(ours, ref):
        assert rel_err(a, b) < 1e-5
    assert float(opt.optimizer.state[ours[0]]['step']) == 4.0
