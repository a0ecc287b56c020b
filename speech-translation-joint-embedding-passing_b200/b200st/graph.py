"""Whole-step CUDA graph: forward_train('ST') + masked NLL + backward captured once and replayed.

The training step launches ~1500 small kernels (31 LAS decoder steps x their GEMMs/cells, 12 Transformer
layers, 4 BLSTM layers ...).  Eagerly, Python + launch overhead is comparable to the GPU time, so the step is
captured into one CUDA graph (static input buffers, graph-private memory pool) and replayed: the reference's
per-step host syncs are already gone (lengths and masks live on the device), which is what makes the step
capturable at all.  bf16 weight shadows are re-cast INSIDE the graph, so replays stay correct after an
optimizer step updates the fp32 parameters in place.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import runtime as rt


class GraphedTrainStep:
    """step = Trainer_ST._train_batch_device(model, items) with items held in static device buffers.

    usage:  g = GraphedTrainStep(model, trainer, example_items);  loss = g(new_items)   # device scalar
    Gradients are left in `p.grad` (static buffers, overwritten by every replay)."""

    def __init__(self, model, trainer, items: Dict, warmup: int = 3, with_optimizer: bool = False):
        """with_optimizer: also capture `trainer.optimizer.step()` (fused clip + Adam, modules/optim.py) behind the
        backward / gradient all-reduce, i.e. the whole of Trainer_ST._train_batch (trainer_st.py:211-299) is one
        graph.  Warm-up passes do NOT step the optimizer.  Set the learning rate with `trainer.optimizer.set_lr()`
        before a replay; the step count lives on the device."""
        self.model, self.trainer = model, trainer
        if with_optimizer and trainer.optimizer is None:
            raise ValueError('with_optimizer=True needs trainer.optimizer')
        dev = next(model.parameters()).device
        self.static = {
            'srcid': [items['srcid'][0].to(dev).clone()],
            'tgtid': [items['tgtid'][0].to(dev).clone()],
            'acous_feat': [items['acous_feat'][0].to(dev).clone()],
            # raw lengths as a DEVICE tensor: the padded-length rule is then applied on the device
            'acouslen': torch.as_tensor([int(n) for n in items['acouslen']], dtype=torch.int32).to(dev),
        }
        self._max_len = int(max(int(n) for n in items['acouslen']))
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model.zero_grad(set_to_none=True)
                trainer._train_batch_device(model, self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if with_optimizer:
            trainer.optimizer._engine().prepare()     # state + pointer tables for the parameters that got a gradient
        model.zero_grad(set_to_none=True)
        self.with_optimizer = with_optimizer
        if with_optimizer:
            # the captured Adam kernel rewrites the bf16 operand copies itself (no cast pass in the graph); every other
            # cached copy (e.g. fp32 concatenations) is marked stale so that its refresh IS captured
            rt.after_raw_update()
        else:
            rt.clear_cache()                  # weight copies get re-cast inside the captured region on every replay
        rt.reset_deferred()
        rt.unit_grad_probes.clear()
        self.graph = torch.cuda.CUDAGraph()
        # captured on a high-priority stream: the kernel nodes of the dependent chain inherit it, the deferred
        # weight-gradient branches (runtime.side_streams pool 'dw', default priority) fill the SMs the chain leaves idle
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=rt.CHAIN_PRIORITY)):
            self.loss = trainer._train_batch_device(model, self.static)
            if with_optimizer:
                trainer.optimizer.step()
        self._epoch = rt.cache_epoch()
        self._probes_pending = bool(rt.unit_grad_probes)

    def _after_replay(self):
        if self._probes_pending:          # first replay only: one host sync
            self._probes_pending = False
            rt.check_unit_grad_probes()

    def _check_epoch(self):
        if rt.cache_epoch() != self._epoch:
            raise RuntimeError('b200st: a cached low-precision weight copy was re-allocated after this step was captured '
                               '(compute dtype switched, model.to(), load_state_dict(assign=True) or a parameter .data swap): '
                               'the graph holds stale pointers; build a new GraphedTrainStep')

    def load(self, items: Dict, non_blocking: bool = True):
        """Copy a new batch (host pinned or device tensors of the captured shapes) into the static buffers."""
        for k in ('srcid', 'tgtid', 'acous_feat'):
            self.static[k][0].copy_(items[k][0], non_blocking=non_blocking)
        lens = items['acouslen']
        if torch.is_tensor(lens):
            self.static['acouslen'].copy_(lens.reshape(-1).to(torch.int32), non_blocking=non_blocking)
        else:
            assert max(int(n) for n in lens) + 8 - max(int(n) for n in lens) % 8 == self.static['acous_feat'][0].size(1)
            self.static['acouslen'].copy_(torch.as_tensor([int(n) for n in lens], dtype=torch.int32), non_blocking=non_blocking)

    # -- input pipelining: the next batch's host->device copy runs on a copy stream while the current step computes ----
    def prefetch(self, items: Dict):
        """Start copying `items` (pinned host tensors of the captured shapes) into a device staging set on a side stream;
        returns immediately.  `step_prefetched()` swaps the staged batch in (three small device-to-device copies) and
        replays the graph, so the PCIe transfer of batch i+1 is hidden under step i."""
        if not hasattr(self, '_stage'):
            self._stage = [{k: torch.empty_like(self.static[k][0]) for k in ('srcid', 'tgtid', 'acous_feat')} |
                           {'acouslen': torch.empty_like(self.static['acouslen'])} for _ in range(2)]
            self._copy_stream = torch.cuda.Stream()
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            self._slot = 0
            self._pending = None
        slot = self._slot
        self._slot ^= 1
        st = self._stage[slot]
        lens = items['acouslen']
        if not torch.is_tensor(lens):
            lens = torch.as_tensor([int(n) for n in lens], dtype=torch.int32)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._free[slot])        # the step that last read this slot has swapped it in
            for k in ('srcid', 'tgtid', 'acous_feat'):
                st[k].copy_(items[k][0], non_blocking=True)
            st['acouslen'].copy_(lens.reshape(-1).to(torch.int32), non_blocking=True)
            self._ready[slot].record(self._copy_stream)
        self._pending = slot

    def step_prefetched(self):
        """Replay the step on the batch handed to the last `prefetch()`."""
        slot = self._pending
        assert slot is not None, 'call prefetch() first'
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[slot])
        st = self._stage[slot]
        for k in ('srcid', 'tgtid', 'acous_feat'):
            self.static[k][0].copy_(st[k], non_blocking=True)
        self.static['acouslen'].copy_(st['acouslen'], non_blocking=True)
        self._free[slot].record(cur)
        self._pending = None
        self._check_epoch()
        self.graph.replay()
        self._after_replay()
        if self.with_optimizer:
            rt.after_raw_update()             # copies the captured kernel does not maintain are stale now
        return self.loss

    def __call__(self, items: Optional[Dict] = None):
        if items is not None:
            self.load(items)
        self._check_epoch()
        self.graph.replay()
        self._after_replay()
        if self.with_optimizer:
            rt.after_raw_update()             # copies the captured kernel does not maintain are stale now
        return self.loss
