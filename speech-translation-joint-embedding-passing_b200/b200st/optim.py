"""Fused gradient-norm clip + Adam behind the reference's `Optimizer.step()`.

Reference: `Optimizer(torch.optim.Adam(model.parameters(), lr=...), max_grad_norm=...)` (trainer_base.py:422-426)
whose `step()` runs `clip_grad_norm_` over every parameter and then `torch.optim.Adam.step()` (modules/optim.py:31-36)
— ~270 tensors, several launches each.  Here the whole step is three launches over a device pointer table
(csrc/optim.cu), with no host synchronisation, so it can sit inside the whole-step CUDA graph right behind backward
(and behind the gradient all-reduce when data-parallel: the clip acts on the reduced gradient).

State stays where torch keeps it: `adam.state[p] = {'step', 'exp_avg', 'exp_avg_sq'}` with torch's shapes and dtypes,
so `optimizer.state_dict()` / `load_state_dict()` round-trip with checkpoints written by the reference.  The only
difference is that every `state[p]['step']` is the SAME 0-dim fp32 CUDA tensor, advanced by the kernel.
Difference from `clip_grad_norm_`: the clip coefficient is applied to the gradient on the fly; `p.grad` itself is
left unscaled (the reference zeroes it right after the step, trainer_st.py:291-292).
"""
from __future__ import annotations

from typing import List

import torch

from . import runtime as rt
from .kernels import K


def _capturing(dev: torch.device) -> bool:
    return dev.type == 'cuda' and torch.cuda.is_current_stream_capturing()


def _pinned(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.pin_memory() if dev.type == 'cuda' else t


class FusedClipAdam:
    """Drives csrc/optim.cu for one `torch.optim.Adam` instance (one hyper-parameter set for all groups).
    Limitation: ONE step counter is shared by every parameter (seeded from the largest existing per-parameter step), so the
    bias correction of a parameter that receives its first gradient later than the others follows the shared count --
    identical to torch.optim.Adam whenever all trained parameters get a gradient from the first step on, which is the case
    for every trainer mode of the reference (parameters without a gradient are never touched)."""

    def __init__(self, adam: torch.optim.Adam, max_grad_norm: float = 0.0):
        if type(adam) is not torch.optim.Adam:
            raise NotImplementedError(f'b200st fused optimizer step implements torch.optim.Adam (what the reference '
                                      f'constructs, trainer_base.py:423), got {type(adam).__name__}')
        g0 = adam.param_groups[0]
        for g in adam.param_groups:
            if g.get('amsgrad') or g.get('maximize'):
                raise NotImplementedError('amsgrad / maximize are not implemented by the fused Adam kernel')
            if any(g[k] != g0[k] for k in ('lr', 'betas', 'eps', 'weight_decay')):
                raise NotImplementedError('parameter groups with different lr/betas/eps/weight_decay (the reference builds '
                                          'ONE group, trainer_base.py:423; the fused kernel applies one set to every tensor)')
        self.adam = adam
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self._sig = None            # (param ids, grad pointers) the device table was built for
        self._bufs = None
        self._step = None           # shared fp32 device scalar = state[p]['step'] of every parameter
        self._lr = None
        self._lr_host = None

    # -- state in torch's own layout ----------------------------------------------------------------
    def _ensure_state(self, params: List[torch.Tensor]):
        dev = params[0].device
        if self._step is None:
            start = 0.0
            for p in params:
                st = self.adam.state.get(p)
                if st and 'step' in st:
                    start = max(start, float(st['step']))
            self._step = torch.full((), start, dtype=torch.float32, device=dev)
            self._lr = torch.zeros(1, dtype=torch.float32, device=dev)
        for p in params:
            st = self.adam.state[p]
            if 'exp_avg' not in st:
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st['step'] = self._step

    def _build(self, params: List[torch.Tensor]):
        dev = params[0].device
        chunk = K().opt_chunk()
        ncol = K().opt_table_cols()
        rows, bmap = [], []
        keys_done, copies = {}, {}
        index = rt.copies_index()
        for i, p in enumerate(params):
            g, st = p.grad, self.adam.state[p]
            for t, what in ((p, 'parameter'), (g, 'gradient'), (st['exp_avg'], 'exp_avg'), (st['exp_avg_sq'], 'exp_avg_sq')):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                    raise RuntimeError(f'fused Adam needs dense fp32 CUDA tensors on one device; {what} #{i} is '
                                       f'{t.dtype} {tuple(t.shape)} strides {t.stride()} on {t.device}')
            # bf16 operand copies of this parameter (b200st.runtime cache): the kernel rewrites up to two of them with
            # the updated value, so no cast pass follows the step; further copies (rare) are marked stale instead
            dests = [(key, t) for key, t in index.get(id(p), []) if t.is_contiguous() and t.numel() == p.numel()][:2]
            for key, _ in dests:
                keys_done.setdefault(key, 0)
                keys_done[key] += 1
            d = [t.data_ptr() for _, t in dests] + [0, 0]
            copies[p] = [t for _, t in dests]
            rows.append([p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                         p.numel(), d[0], d[1]])
            bmap += [[i, c] for c in range((p.numel() + chunk - 1) // chunk)]
        b = self._bufs
        if b is None or b['table'].size(0) != len(rows) or b['blockmap'].size(0) != len(bmap):
            if _capturing(dev):
                raise RuntimeError('the set of parameters with gradients changed inside a CUDA graph capture; run '
                                   'FusedClipAdam.prepare() after a warm-up backward, before capturing')
            blockmap = _pinned(torch.tensor(bmap, dtype=torch.int32), dev).to(dev, non_blocking=True)
            b = self._bufs = {'blockmap': blockmap, 'partials': torch.empty(len(bmap), dtype=torch.float32, device=dev),
                              'scal': torch.zeros(4, dtype=torch.float32, device=dev),
                              'table_host': _pinned(torch.empty((len(rows), ncol), dtype=torch.int64), dev),
                              'table': torch.empty((len(rows), ncol), dtype=torch.int64, device=dev)}
        # a cached copy is "maintained" when every member parameter writes its block of it (a concatenated K|V copy needs
        # both of its members in the table)
        maintained = []
        for key, n in keys_done.items():
            need = len(key) - 1 if isinstance(key, tuple) else 1
            if n >= need:
                maintained.append(key)
        rt.set_maintained(maintained)
        self._copies = copies
        if hasattr(K(), 'adam_copies'):
            K().adam_copies = copies
        host = torch.tensor(rows, dtype=torch.int64)
        if _capturing(dev):
            # a capture records a copy NODE that re-reads its source on every replay: stage through the persistent
            # pinned buffer, which is written here and never again while the graph lives
            b['table_host'].copy_(host)
            b['table'].copy_(b['table_host'], non_blocking=True)
        else:
            # eager: the GPU may still be a step behind, so never overwrite a staging buffer a pending copy reads --
            # a fresh pinned tensor per rebuild (the caching host allocator recycles it once the copy has run)
            b['table'].copy_(_pinned(host, dev), non_blocking=True)

    def prepare(self):
        """Allocate state and pointer tables for the parameters that currently hold a gradient, without stepping.
        Call after a warm-up backward and before capturing `step()` into a CUDA graph."""
        params = [p for g in self.adam.param_groups for p in g['params'] if p.grad is not None]
        if params:
            self._ensure_state(params)
            self._build(params)
            self._sig = (tuple(id(p) for p in params), tuple(p.grad.data_ptr() for p in params), rt.cache_len())
            self.set_lr(self.adam.param_groups[0]['lr'])

    # -- the step ---------------------------------------------------------------------------------------
    def set_lr(self, lr: float):
        """Writes the learning rate into its device scalar (a fill launch, no sync).  `step()` does this itself from
        `param_groups[0]['lr']`; call it explicitly before replaying a CUDA graph that captured `step()`."""
        if self._lr is not None and lr != self._lr_host:
            self._lr.fill_(float(lr))
            self._lr_host = float(lr)

    def step(self):
        params = [p for g in self.adam.param_groups for p in g['params'] if p.grad is not None]
        if not params:
            return
        sig = (tuple(id(p) for p in params), tuple(p.grad.data_ptr() for p in params), rt.cache_len())
        if sig != self._sig:
            self._ensure_state(params)
            self._build(params)
            self._sig = sig
        g0 = self.adam.param_groups[0]
        if not _capturing(params[0].device):
            self.set_lr(g0['lr'])
        elif self._lr_host is None:
            raise RuntimeError('call set_lr() once before capturing the optimizer step into a CUDA graph')
        b = self._bufs
        K().clip_adam_step((params, [p.grad for p in params], [self.adam.state[p]['exp_avg'] for p in params],
                            [self.adam.state[p]['exp_avg_sq'] for p in params]),
                           b['table'], b['blockmap'], b['partials'], b['scal'], self._step, self._lr,
                           max_grad_norm=self.max_grad_norm, beta1=g0['betas'][0], beta2=g0['betas'][1],
                           eps=g0['eps'], weight_decay=g0['weight_decay'])
        # the kernels wrote through raw pointers: parameter version counters did not move.  Operand copies the kernel
        # rewrote itself stay valid; any other cached copy is marked stale (re-cast in place on next use)
        rt.after_raw_update()

    @property
    def grad_norm(self) -> torch.Tensor:
        """||g||_2 of the last step before clipping (device scalar; reading it synchronises)."""
        return self._bufs['scal'][3]
