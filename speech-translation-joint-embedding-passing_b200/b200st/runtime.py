"""Process-wide numerical mode and cached low-precision operands.

Compute dtype: 'fp32' (exact CUDA-core GEMMs; the 1e-4 parity mode) or 'bf16' (tcgen05 tensor-core GEMMs,
fp32 accumulation; the 2e-2 mode the benchmark runs).  nn.Parameters always stay fp32 and keep their
reference names; in bf16 mode kernels read a bf16 shadow copy that is refreshed whenever the parameter's
version counter changes (optimizer step, load_state_dict), so no re-packed weight ever escapes state_dict.
"""
from __future__ import annotations

import os
import weakref

import torch

_DTYPES = {'fp32': torch.float32, 'float32': torch.float32, 'bf16': torch.bfloat16,
           'bfloat16': torch.bfloat16}
_compute_dtype = _DTYPES[os.environ.get('B200ST_DTYPE', 'fp32').lower()]
_cache = {}
_epoch = [0]               # bumped whenever a cached operand copy is (re)created or dropped: captured CUDA graphs that
                           # baked an operand pointer compare it to know when they are stale


_las_persistent = [os.environ.get('B200ST_LAS_PERSISTENT', '1') != '0']


def las_persistent(on=None) -> bool:
    """The persistent one-launch LAS decoder loop (csrc/las_decoder.cu) is used where its shape constraints hold (bf16,
    decoder width 512, 3 layers, no dropout); `las_persistent(False)` forces the step-by-step kernels (test hook).
    Returns the previous / current setting."""
    old = _las_persistent[0]
    if on is not None:
        _las_persistent[0] = bool(on)
    return old


def cache_epoch() -> int:
    return _epoch[0]


def cache_len() -> int:
    return len(_cache) + (_epoch[0] << 20)


def set_compute_dtype(d):
    global _compute_dtype
    _compute_dtype = _DTYPES[d.lower()] if isinstance(d, str) else d
    _cache.clear()
    _epoch[0] += 1


def compute_dtype() -> torch.dtype:
    return _compute_dtype


def operand(param: torch.Tensor) -> torch.Tensor:
    """The tensor a GEMM should read for `param` in the current compute dtype.  A cached copy is refreshed IN PLACE when
    the parameter changed (same storage, so pointers baked into captured CUDA graphs stay valid; the epoch tells their
    owners that the contents were stale until this refresh)."""
    p = param.detach()
    if p.dtype == _compute_dtype:
        return p
    key = id(param)
    hit = _cache.get(key)
    ver = param._version
    if hit is not None and hit[0]() is param and hit[2] == p.data_ptr():
        if hit[1] == ver:
            return hit[3]
        if hit[3].shape == p.shape and hit[3].dtype == _compute_dtype:
            from .kernels import K
            K().cast(p.contiguous(), _compute_dtype, out=hit[3])
            _cache[key] = (hit[0], ver, hit[2], hit[3])
            return hit[3]
    from .kernels import K
    shadow = K().cast(p.contiguous(), _compute_dtype)
    if hit is not None:
        _epoch[0] += 1         # a NEW buffer replaces one that captured graphs may have baked in (model.to, p.data swap)
    _cache[key] = (weakref.ref(param), ver, p.data_ptr(), shadow)
    return shadow


def operand_cat(*params: torch.Tensor) -> torch.Tensor:
    """Row-wise concatenation [sum(out_i), in] of several weight matrices in the compute dtype (e.g. w_ks | w_vs so
    that K and V are projected by ONE GEMM).  Cached like `operand`; refreshed in place when any member changed."""
    key = ('cat',) + tuple(id(p) for p in params)
    vers = tuple(p._version for p in params)
    ptrs = tuple(p.data_ptr() for p in params)
    hit = _cache.get(key)
    same = hit is not None and all(r() is p for r, p in zip(hit[0], params)) and hit[2] == ptrs
    if same and hit[1] == vers:
        return hit[3]
    from .kernels import K
    rows = [p.size(0) for p in params]
    buf = hit[3] if same else torch.empty((sum(rows), params[0].size(1)), dtype=_compute_dtype, device=params[0].device)
    if hit is not None and not same:
        _epoch[0] += 1         # re-allocated: stale pointers in captured graphs
    r0 = 0
    for p, n in zip(params, rows):
        src = p.detach().contiguous()
        if src.dtype == _compute_dtype:
            buf[r0:r0 + n].copy_(src)
        else:
            K().cast(src, _compute_dtype, out=buf[r0:r0 + n])
        r0 += n
    _cache[key] = (tuple(weakref.ref(p) for p in params), vers, ptrs, buf)
    return buf


_maintained = set()        # cache keys whose copies the fused optimizer kernel rewrites itself (b200st/optim.py)


def copies_index():
    """{id(param): [(cache key, bf16 tensor view)]} over every cached operand copy (a parameter's own copy and its row
    block inside concatenated copies) -- the destinations the fused Adam kernel can keep up to date."""
    out = {}
    for key, hit in _cache.items():
        if hit[3].dtype != torch.bfloat16:
            continue
        if isinstance(key, tuple):
            members = [r() for r in hit[0]]
            if any(m is None for m in members):
                continue
            r0 = 0
            for m in members:
                out.setdefault(id(m), []).append((key, hit[3][r0:r0 + m.size(0)]))
                r0 += m.size(0)
        else:
            p = hit[0]()
            if p is not None:
                out.setdefault(id(p), []).append((key, hit[3]))
    return out


def set_maintained(keys):
    _maintained.clear()
    _maintained.update(keys)


def after_raw_update(updated=None):
    """Parameters were written through raw pointers (fused optimizer step): every cached copy that the kernel did not
    rewrite itself is marked stale.  A concatenated copy counts as maintained only if it was registered as such."""
    for key, hit in list(_cache.items()):
        if key in _maintained:
            continue
        vers = -1 if not isinstance(hit[1], tuple) else tuple(-1 for _ in hit[1])
        _cache[key] = (hit[0], vers, hit[2], hit[3])


def refresh_all():
    """Re-cast (in place) every cached operand copy whose parameter changed since it was made.  Owners of captured
    graphs that read the copies WITHOUT re-casting them (the inference graphs) call this before a replay; it is a host
    loop over the cache and launches nothing when the weights did not change."""
    for key, hit in list(_cache.items()):
        if isinstance(hit[1], tuple):
            params = [r() for r in hit[0]]
            if all(p is not None for p in params) and tuple(p._version for p in params) != hit[1]:
                operand_cat(*params)
        else:
            p = hit[0]()
            if p is not None and p._version != hit[1]:
                operand(p)


def clear_cache():
    """Marks every cached operand copy stale (parameters were written through raw pointers, e.g. by the fused optimizer
    step): the next `operand()` re-casts into the SAME buffer.  Captured graphs that read those buffers without
    re-casting (inference graphs) call `refresh_all()` before a replay."""
    for key, hit in list(_cache.items()):
        vers = -1 if not isinstance(hit[1], tuple) else tuple(-1 for _ in hit[1])
        _cache[key] = (hit[0], vers, hit[2], hit[3])


# ------------------------------------------------------------------------------------------------
# side streams: independent kernels of one step are enqueued on forked streams; inside a CUDA graph
# capture they become parallel branches, in eager mode they overlap through the hardware queues.
# ------------------------------------------------------------------------------------------------
_side = {}
_rot = [0]
DW_STREAMS = 4            # streams the deferred weight-gradient work rotates over
CHAIN_PRIORITY = -1       # CUDA stream priority of the dependent chain (capture stream + its forked branches); deferred work: 0


def side_streams(device, n, pool='main'):
    """`n` persistent side streams for `device` (from the named pool).  Returns None entries (fork/join become no-ops, everything runs
    inline) on CPU and whenever no CUDA graph is being captured: eagerly, the event traffic of forking costs more
    host time than the overlap wins, while inside a capture the forks are free and become parallel graph branches."""
    if device.type != 'cuda' or not torch.cuda.is_current_stream_capturing():
        return [None] * n
    key = (device.index if device.index is not None else torch.cuda.current_device(), pool)
    streams = _side.setdefault(key, [])
    if pool == 'dw' and DW_STREAMS > n:
        # Deferred weight-gradient work (joined only at the end of backward): the callers' requests rotate over DW_STREAMS
        # default-priority streams.  With two, the ~250 deferred GEMMs / column sums of a step queued up behind each other
        # and the optimizer ended up waiting for that backlog (4 streams: -0.10 ms; 8: no further change).
        while len(streams) < DW_STREAMS:
            streams.append(torch.cuda.Stream(device=device, priority=0))
        off = _rot[0] % DW_STREAMS
        _rot[0] += n
        return [streams[(off + i) % DW_STREAMS] for i in range(n)]
    while len(streams) < n:
        # branches of the step's dependent chain share the capture stream's HIGH priority (b200st/graph.py): a chain
        # kernel that is ready takes free SMs before deferred work does (-0.10 ms once the backlog above is gone; with
        # the backlog it made the step slower)
        streams.append(torch.cuda.Stream(device=device, priority=0 if pool == 'dw' else CHAIN_PRIORITY))
    return streams[:n]


class fork:
    """`with fork(stream): launch(...)` — the body runs on `stream`, ordered after everything enqueued so far on the
    current stream.  `stream is None` (CPU) runs the body inline."""

    def __init__(self, stream):
        self.stream = stream
        self.ctx = None

    def __enter__(self):
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join(stream):
    """Make the current stream wait for everything enqueued so far on `stream`."""
    if stream is not None:
        torch.cuda.current_stream().wait_stream(stream)


# Deferred joins: weight-gradient GEMMs of a backward node are forked onto side streams and NOT joined before the node
# returns -- nothing downstream in backward reads a weight gradient, so they overlap the rest of the backward pass (the
# next layer's latency-bound recurrence uses 64 of the 148 SMs).  Whoever consumes gradients (the trainer after
# loss.backward(), the data-parallel reducer, the optimizer) calls join_deferred() / waits on deferred_streams() first.
# The tensors those launches read are kept alive until then: a block freed at the end of the node could otherwise be
# handed to a main-stream kernel that has no ordering against the side stream.
_deferred = []
_keepalive = []
_expect = []               # (weakref(param), data_ptr of the gradient tensor returned for it) -- see defer()


def can_defer(*params) -> bool:
    """Deferral is only safe when autograd will ADOPT the returned gradient tensors as `param.grad` (first accumulation
    of the step): an existing `.grad` makes AccumulateGrad add in place on the main stream right away."""
    return all(p is None or p.grad is None for p in params)


def defer(stream, tensors=(), grads=()):
    """`tensors`: inputs the side-stream launches read (kept alive until the join).  `grads`: (param, gradient tensor)
    pairs; only the gradient's ADDRESS is remembered (holding the tensor would stop autograd from adopting it)."""
    if stream is None:
        return
    if stream not in _deferred:
        _deferred.append(stream)
    _keepalive.extend(t for t in tensors if t is not None)      # (lists of tensors are kept as they are)
    for prm, g in grads:
        if prm is not None and g is not None:
            _expect.append((weakref.ref(prm), g.data_ptr()))


def deferred_streams():
    return list(_deferred)


def reset_deferred():
    """Forget bookkeeping left behind by a backward pass that did not reach join_deferred() (an exception inside a
    capture); called before a new capture starts."""
    _deferred.clear()
    _keepalive.clear()
    _expect.clear()


def join_deferred():
    for s in _deferred:
        join(s)
    _deferred.clear()
    _keepalive.clear()
    bad = []
    for ref, ptr in _expect:
        prm = ref()
        if prm is not None and (prm.grad is None or prm.grad.data_ptr() != ptr):
            bad.append(tuple(prm.shape))
    _expect.clear()
    if bad:
        raise RuntimeError(f'b200st: autograd copied {len(bad)} weight-gradient tensor(s) instead of adopting them while '
                           f'their GEMMs were still running on a side stream (shapes {bad[:4]}...); the copies are '
                           f'undefined.  This is a bug in the deferred-join bookkeeping, not in the caller.')



# ------------------------------------------------------------------------------------------------
# Pre-norm hand-over: a sub-layer whose last GEMM ran with the fused residual + LayerNorm epilogue
# (b200st_gemm_ln) has already produced LayerNorm(out) for the sub-layer that consumes `out` next.  The
# producer offers it here; the very next LayerNorm on the path takes it if it is about to normalise
# exactly that tensor with exactly those parameters, and launches nothing.  One slot, strictly
# producer -> next consumer: any other LayerNorm call in between discards the offer.
# ------------------------------------------------------------------------------------------------
_prenorm = [None]

# Upstream gradients of the fused loss that a capture ASSUMED to be exactly 1 (functional._FusedSoftmaxNLL.backward skips a
# full pass over [rows, V] then): device copies taken inside the capture, checked on the host after the first replay.
unit_grad_probes = []


def check_unit_grad_probes():
    """Raise if a captured graph skipped the `dlogits * g` pass although its upstream gradient is not 1."""
    probes, bad = list(unit_grad_probes), []
    unit_grad_probes.clear()
    for t in probes:
        v = float(t)
        if v != 1.0:
            bad.append(v)
    if bad:
        raise RuntimeError(f'b200st: a CUDA graph was captured assuming the fused loss receives an upstream gradient of 1 '
                           f'(as its eager warm-up did), but the replay saw {bad}: capture with the same loss scaling as the '
                           f'warm-up passes')


def offer_prenorm(out, ln_w, ln_b, eps, yn, mean, rstd):
    _prenorm[0] = (out.data_ptr(), out.numel(), out.dtype, ln_w.data_ptr(), ln_b.data_ptr(), float(eps), yn, mean, rstd)


def take_prenorm(x, ln_w, ln_b, eps):
    """(LayerNorm(x), mean, rstd) computed by the producer of `x`, or None."""
    e, _prenorm[0] = _prenorm[0], None
    if e is None:
        return None
    if (x.data_ptr(), x.numel(), x.dtype, ln_w.data_ptr(), ln_b.data_ptr(), float(eps)) != e[:6]:
        return None
    return e[6], e[7], e[8]


# ------------------------------------------------------------------------------------------------
# dropout randomness: counter-based (csrc/philox.cuh).  A mask is a pure function of
# (seed, step, site, element index); seed and step live in a 2 x int64 DEVICE array so that a captured
# CUDA graph draws fresh masks on every replay (`begin_step` is a captured launch that adds 1 to step).
# `site` numbers the dropout calls inside one forward pass (host counter, deterministic call order);
# backward passes the site it saved, so no mask tensor is ever stored.
# ------------------------------------------------------------------------------------------------
_rng = {}
_seed = [None]             # set by manual_seed(); devices that create their state later start from it too
_site = [0]
site_log = None            # TEST HOOK: set to a dict to record {tag: (site, p)} of every dropout call


def rng_state(device) -> torch.Tensor:
    key = (device.type, device.index)
    st = _rng.get(key)
    if st is None:
        seed = _seed[0] if _seed[0] is not None else torch.initial_seed() & 0x7fffffffffffffff
        st = torch.tensor([seed, 0], dtype=torch.int64).to(device)
        _rng[key] = st
    return st


def manual_seed(seed: int):
    """Re-seed the dropout generator (all devices, including those whose state is created later) and restart its step
    counter."""
    _seed[0] = int(seed) & 0x7fffffffffffffff
    for st in _rng.values():
        st.copy_(torch.tensor([int(seed) & 0x7fffffffffffffff, 0], dtype=torch.int64))
    _rng_step.clear()
    new_step()


_pending = [True]


def new_step():
    """Marks the start of a forward pass (Seq2seq.forward_train / LAS.forward in training mode): site numbering
    restarts and the first dropout call of the pass advances the device step counter (fresh masks).  Costs nothing
    when no dropout is active."""
    _site[0] = 0
    _pending[0] = True
    _prenorm[0] = None        # an offer nobody took (a forward pass that ended early) must not outlive its pass


def next_site(tag: str = '', p: float = 0.0) -> int:
    _site[0] += 1
    if site_log is not None:
        site_log[tag] = (_site[0], p)
    return _site[0]


_rng_step = {}


def current_rng(device) -> torch.Tensor:
    """The [seed, step] pair of THIS forward pass (a snapshot taken at its first dropout call): the forward launch and
    the backward launch of a dropout site both read it, whatever happens to the live counter in between."""
    key = (device.type, device.index)
    if _pending[0] or key not in _rng_step:
        from .kernels import K
        st = rng_state(device)
        K().rng_advance(st)
        _rng_step[key] = st.clone()
        _pending[0] = False
    return _rng_step[key]
