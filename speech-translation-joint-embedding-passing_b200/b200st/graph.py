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
        rt.clear_cache()                      # weight shadows get (re)built inside the captured region
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = trainer._train_batch_device(model, self.static)
            if with_optimizer:
                trainer.optimizer.step()

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

    def __call__(self, items: Optional[Dict] = None):
        if items is not None:
            self.load(items)
        self.graph.replay()
        return self.loss
