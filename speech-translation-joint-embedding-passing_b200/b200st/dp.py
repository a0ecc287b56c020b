"""Data-parallel gradient exchange: bucketed NCCL all-reduce launched from autograd hooks so that it
overlaps with the rest of backward (the BLSTM backward is the long tail that hides the other buckets).

The reference has no multi-GPU path; its batch-splitting mechanism is serial gradient accumulation over
`minibatch_partition` slices (trainer_base.py:83-85, trainer_st.py:225-290), each slice's loss normalised
by its own #non-PAD tokens and divided by n_minibatch.  N ranks x per-rank normalisation + MEAN all-reduce
is arithmetically the same thing, so `batch_size=N*b, minibatch_partition=N` on the reference is the
oracle for N ranks of batch b here.  No token-count all-reduce is needed.

Parameters that never receive a gradient (template layers enc_src.enc.*, dec_tgt.dec.*, and
las.decoder.acous_out.* in ST-only mode) simply never fire their hook; buckets flush what is ready at
`finish()`, so nothing waits on them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20,
                 process_group: Optional[dist.ProcessGroup] = None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_bytes = bucket_bytes
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._pending: List[torch.Tensor] = []      # grads ready but not yet sent
        self._pending_bytes = 0
        self._inflight = []                          # (work, flat, [grads])
        self._hooks = []
        self._stream = None
        self._armed = True
        self.trace = None              # set to [] to record (start event, end event, bytes, n tensors) per bucket (diagnostics)
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # -- hook path --------------------------------------------------------------------------------
    def arm(self, on: bool = True):
        """Gradient accumulation over `minibatch_partition` slices (trainer_st.py:225-290): the trainers arm the hooks
        for the LAST slice's backward only, so that every parameter's ACCUMULATED gradient is queued and reduced exactly
        once.  (Reducing in place while a later slice's AccumulateGrad adds into the same tensor on the compute stream
        would race, and would multiply the traffic by n_minibatch.)"""
        self._armed = bool(on)

    def _on_grad(self, p: torch.Tensor):
        g = p.grad
        if g is None or not self._armed:
            return
        self._pending.append(g)
        self._pending_bytes += g.numel() * g.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self._pending:
            return
        grads, self._pending, self._pending_bytes = self._pending, [], 0
        if grads[0].is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            # the bucket becomes ready on the compute stream; ship it from a side stream.  NCCL: the bucket's gradient
            # tensors are reduced IN PLACE by one grouped (coalesced) launch with ReduceOp.AVG -- no flatten copy, no
            # divide pass, no unflatten copies at the end of backward (those used to be the serial tail of the step).
            self._stream.wait_stream(torch.cuda.current_stream())
            from . import runtime as rt
            for st in rt.deferred_streams():       # weight-gradient GEMMs still running on side streams (graph capture)
                self._stream.wait_stream(st)
            with torch.cuda.stream(self._stream):
                if dist.get_backend(self.group) == 'nccl':
                    ev = None
                    if self.trace is not None:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        ev[0].record(self._stream)
                    with dist._coalescing_manager(self.group, device=grads[0].device, async_ops=True) as cm:
                        for g in grads:
                            dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                    if ev is not None:
                        cm.wait()                         # stream-level wait (the NCCL kernels run on the group's own stream)
                        ev[1].record(self._stream)
                        self.trace.append((ev[0], ev[1], sum(g.numel() * g.element_size() for g in grads), len(grads)))
                    self._inflight.append((cm, None, grads))
                    return
                flat = torch._utils._flatten_dense_tensors(grads)
                flat.div_(self.world)
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            flat = torch._utils._flatten_dense_tensors(grads)
            flat.div_(self.world)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((work, flat, grads))

    def finish(self):
        """Call after backward(): sends the last partial bucket, waits, writes the means back."""
        if self.world <= 1:
            return
        self._flush()
        for work, flat, grads in self._inflight:
            work.wait()
            if flat is None:                 # reduced in place
                continue
            if flat.is_cuda:
                with torch.cuda.stream(self._stream):
                    for g, s in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                        g.copy_(s)
            else:
                for g, s in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                    g.copy_(s)
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._inflight = []

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
