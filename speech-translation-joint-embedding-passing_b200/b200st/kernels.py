"""Tensor-level wrappers over the C ABI (include/b200st.h).

`CudaKernels` is the only implementation shipped: every method launches hand-written sm_100a kernels from
libb200st.so on the current CUDA stream.  PyTorch is used for device memory (torch.empty) and the stream
handle only.  Tests may install a different object with the same methods via `set_backend()` to exercise
the host-side orchestration on a machine without a GPU; nothing in the product does.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

from . import lib as _lib

F32, BF16 = 0, 1


def _dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if torch.is_tensor(t_or_dtype) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise TypeError(f'b200st kernels take float32 or bfloat16 activations, got {d}')


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor whose last dim is dense; returned as-is (row stride may exceed the width)."""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.size(1) == 1), (t.shape, t.stride())
    return t


class CudaKernels:
    name = 'cuda'

    def __init__(self):
        self.lib = _lib.load()
        mode = os.environ.get('B200ST_BLSTM_BACKEND')      # profiling hook: see b200st_set_blstm_backend in include/b200st.h
        if mode is not None:
            self.lib.b200st_set_blstm_backend(int(mode))

    # -- plumbing ---------------------------------------------------------------------------------
    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _need_cuda(*ts):
        for t in ts:
            if t is not None and not t.is_cuda:
                raise RuntimeError('b200st kernels are CUDA-only (sm_100a); got a CPU tensor. '
                                   'There is no CPU fallback.')

    def launch_count(self) -> int:
        return int(self.lib.b200st_launch_count())

    def set_gemm_backend(self, mode: int) -> int:
        """0 auto, 1 CUDA cores only, 2 tensor cores required.  Returns the previous mode (test hook)."""
        return int(self.lib.b200st_set_gemm_backend(int(mode)))

    def set_gemm_persistent(self, on: int) -> int:
        """1 = persistent tcgen05 GEMM for the many-tile shapes (default), 0 = one tile per CTA.  Returns the old value."""
        return int(self.lib.b200st_set_gemm_persistent(int(on)))

    def debug_stamp(self, slot):
        """slot: 1-element int64 CUDA tensor (view); receives the device nanosecond timer at this point of the stream."""
        _lib.check(self.lib.b200st_debug_stamp(_p(slot), self._stream()), 'debug_stamp')

    def set_gemm_sm_budget(self, n: int) -> int:
        return int(self.lib.b200st_set_gemm_sm_budget(int(n)))

    def set_blstm_backend(self, mode: int) -> int:
        """0 auto (tcgen05 recurrence for bf16/H=256), 1 CUDA cores only.  Returns the previous mode."""
        return int(self.lib.b200st_set_blstm_backend(int(mode)))

    def set_mha_backend(self, mode: int) -> int:
        """Bit mask: 1 = Transformer attention core on CUDA-core tiles only (default tcgen05 for bf16, d=64, L<=64);
        2 = LAS attention step with one CTA per sequence (default: 4-CTA key-split cluster).  Returns the previous mask."""
        return int(self.lib.b200st_set_mha_backend(int(mode)))

    # -- GEMM -------------------------------------------------------------------------------------
    def gemm(self, a, b, *, trans_a=False, trans_b=False, out=None, out_dtype=None, bias=None,
             relu=False, residual=None, alpha=1.0, relu_gate=None):
        """out = relu?(alpha * op(a) @ op(b) + bias) + residual.  a, b: 2-D (row-strided) or 3-D batched
        (batch stride arbitrary).  op(a) is [M,K], op(b) is [K,N]."""
        if relu_gate is not None:       # out = (relu_gate > 0) ? a @ b : 0  -- rides the residual slot with relu mode 2
            assert residual is None and not relu
            residual, relu = relu_gate, 2
        self._need_cuda(a, b, out, bias, residual)
        batched = a.dim() == 3
        if batched:
            assert b.dim() == 3 and a.size(0) == b.size(0)
            nb = a.size(0)
            a2, b2 = a[0], b[0]
            sa, sb = a.stride(0), b.stride(0)
        else:
            nb, a2, b2, sa, sb = 1, a, b, 0, 0
        _rows(a2), _rows(b2)
        M, K = (a2.size(1), a2.size(0)) if trans_a else (a2.size(0), a2.size(1))
        Kb, N = (b2.size(1), b2.size(0)) if trans_b else (b2.size(0), b2.size(1))
        assert K == Kb, f'gemm inner dims differ: {K} vs {Kb}'
        assert a.dtype == b.dtype
        if out is None:
            od = out_dtype or (residual.dtype if residual is not None else a.dtype)
            out = torch.empty((nb, M, N) if batched else (M, N), dtype=od, device=a.device)
        o2 = out[0] if batched else out
        _rows(o2)
        assert o2.size(0) == M and o2.size(1) == N, (o2.shape, M, N)
        so = out.stride(0) if batched else 0
        if residual is not None:
            r2 = residual[0] if batched else residual
            _rows(r2)
            assert residual.dtype == out.dtype and r2.shape == o2.shape
            ldr, sr = r2.stride(0), (residual.stride(0) if batched else 0)
        else:
            ldr, sr = 0, 0
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        _lib.check(self.lib.b200st_gemm(
            _dt(a), _dt(out), int(trans_a), int(trans_b), M, N, K, float(alpha),
            _p(a), a2.stride(0), sa, _p(b), b2.stride(0), sb, _p(out), o2.stride(0), so,
            _p(residual), ldr, sr, _p(bias), int(relu), nb, self._stream()), 'gemm')
        return out

    def gemm2(self, a, b, a2, b2, *, trans_b=False, out=None, out_dtype=None, bias=None, residual=None, alpha=1.0):
        """out = alpha * (a @ op(b) + a2 @ op(b2)) + bias + residual, 2-D row-strided operands: one launch with a
        two-segment K loop (tensor-core path) instead of a GEMM plus an accumulate-into GEMM."""
        self._need_cuda(a, b, a2, b2, out, bias, residual)
        for t in (a, b, a2, b2):
            _rows(t)
        M, K = a.shape
        K2 = a2.size(1)
        N = b.size(0) if trans_b else b.size(1)
        assert a2.size(0) == M and (b.size(1) if trans_b else b.size(0)) == K
        assert (b2.size(0) if trans_b else b2.size(1)) == N and (b2.size(1) if trans_b else b2.size(0)) == K2
        assert a.dtype == b.dtype == a2.dtype == b2.dtype
        if out is None:
            out = torch.empty((M, N), dtype=out_dtype or (residual.dtype if residual is not None else a.dtype),
                              device=a.device)
        _rows(out)
        ldr = 0
        if residual is not None:
            _rows(residual)
            assert residual.dtype == out.dtype and residual.shape == out.shape
            ldr = residual.stride(0)
        _lib.check(self.lib.b200st_gemm2(_dt(a), _dt(out), 0, int(trans_b), M, N, K, K2, float(alpha), _p(a), a.stride(0),
                                         _p(b), b.stride(0), _p(a2), a2.stride(0), _p(b2), b2.stride(0), _p(out),
                                         out.stride(0), _p(residual), ldr, _p(bias), 0, self._stream()), 'gemm2')
        return out

    def gemm_ln_ok(self, a, w, residual=None, bias=None) -> bool:
        """Can `gemm_ln` serve  a @ w^T (+ bias) + residual  followed by a LayerNorm over the 512 output columns?"""
        if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or a.dim() != 2 or w.dim() != 2 or not a.is_cuda:
            return False
        if w.size(0) != 512 or a.size(1) != w.size(1) or a.size(1) % 8 or a.size(1) < 64:
            return False
        ts = [a, w] + ([residual] if residual is not None else [])
        if any(t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16 for t in ts):
            return False
        return residual is None or (residual.dtype == torch.bfloat16 and tuple(residual.shape) == (a.size(0), 512))

    def gemm_ln(self, a, w, bias, residual, gamma, beta, eps, dropout=None):
        """y = dropout?(a @ w^T (+ bias)) + residual;  yn, mean, rstd = LayerNorm(y; gamma, beta, eps)  in ONE launch
        (csrc/gemm_ln.cu).  a [M, K], w [512, K] bf16; dropout = None or (p, rng_state, site).  Returns (y, yn, mean, rstd)."""
        self._need_cuda(a, w, bias, residual, gamma, beta)
        dp, rng, site = dropout if dropout is not None and dropout[0] > 0 else (0.0, None, 0)
        M, K = a.shape
        N = w.size(0)
        y = torch.empty((M, N), dtype=a.dtype, device=a.device)
        yn = torch.empty((M, N), dtype=a.dtype, device=a.device)
        mean = torch.empty(M, dtype=torch.float32, device=a.device)
        rstd = torch.empty(M, dtype=torch.float32, device=a.device)
        _lib.check(self.lib.b200st_gemm_ln(_dt(a), M, N, K, _p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(residual),
                                           residual.stride(0) if residual is not None else 0, _p(y), N, _p(gamma), _p(beta),
                                           float(eps), _p(yn), N, _p(mean), _p(rstd), float(dp), _p(rng), int(site),
                                           self._stream()), 'gemm_ln')
        return y, yn, mean, rstd

    def gemm_lnbwd_ok(self, a, w, x, add=None) -> bool:
        """Can `gemm_lnbwd` serve  LayerNorm-backward(a @ w; x, ...) + add  (w [K, 512], x / add dense [M, 512] bf16)?"""
        if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or a.dim() != 2 or w.dim() != 2 or not a.is_cuda:
            return False
        if w.size(1) != 512 or a.size(1) != w.size(0) or a.size(1) % 8 or a.size(1) < 64:
            return False
        if any(t.stride(1) != 1 or t.stride(0) % 8 or t.data_ptr() % 16 for t in (a, w)):
            return False
        for t in (x, add):
            if t is not None and not (t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() == a.size(0) * 512
                                      and t.data_ptr() % 16 == 0):
                return False
        return True

    def gemm_lnbwd(self, a, w, x, gamma, mean, rstd, add=None):
        """dx = LayerNorm-backward(dy = a @ w; x, gamma, mean, rstd) [+ add] in ONE launch (csrc/gemm_ln.cu).
        Returns (dx [M, 512], partials fp32 [row blocks, 1024] = dgamma | dbeta contributions to be column-summed)."""
        self._need_cuda(a, w, x, gamma, mean, rstd, add)
        M, K = a.shape
        N = w.size(1)
        dx = torch.empty((M, N), dtype=a.dtype, device=a.device)
        partials = torch.empty((int(self.lib.b200st_gemm_lnbwd_blocks(M)), 2 * N), dtype=torch.float32, device=a.device)
        _lib.check(self.lib.b200st_gemm_lnbwd(_dt(a), M, N, K, _p(a), a.stride(0), _p(w), w.stride(0), _p(x), _p(gamma),
                                              _p(mean), _p(rstd), _p(add), _p(dx), _p(partials), self._stream()), 'gemm_lnbwd')
        return dx, partials

    # -- LayerNorm --------------------------------------------------------------------------------
    def layernorm_fwd(self, x, gamma, beta, eps, save_stats=True):
        self._need_cuda(x, gamma, beta)
        assert x.is_contiguous()
        rows, cols = x.numel() // x.size(-1), x.size(-1)
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save_stats else None
        _lib.check(self.lib.b200st_layernorm_fwd(_dt(x), _p(x), _p(gamma), _p(beta), _p(y), _p(mean),
                                                 _p(rstd), rows, cols, float(eps), self._stream()),
                   'layernorm_fwd')
        return y, mean, rstd

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, dgamma, dbeta, add=None):
        """dx = [add +] LayerNorm gradient; dgamma / dbeta (fp32, pre-zeroed or running sums) are accumulated into."""
        assert dy.is_contiguous() and x.is_contiguous() and (add is None or (add.is_contiguous() and add.dtype == x.dtype))
        rows, cols = x.numel() // x.size(-1), x.size(-1)
        dx = torch.empty_like(x)
        _lib.check(self.lib.b200st_layernorm_bwd_add(_dt(x), _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(add),
                                                     _p(dx), _p(dgamma), _p(dbeta), rows, cols,
                                                     self._stream()), 'layernorm_bwd')
        return dx

    def layernorm_bwd_partial(self, dy, x, gamma, mean, rstd, add=None):
        """(dx, partials) with partials fp32 [n_blocks, 2 * cols]: per-CTA dgamma | dbeta contributions, to be column-summed
        by the caller (off the critical path); None for partials' shape support -> (None, None)."""
        assert dy.is_contiguous() and x.is_contiguous() and (add is None or (add.is_contiguous() and add.dtype == x.dtype))
        rows, cols = x.numel() // x.size(-1), x.size(-1)
        nb = int(self.lib.b200st_layernorm_bwd_partial_blocks(rows, cols))
        if nb == 0:
            return None, None
        dx = torch.empty_like(x)
        partials = torch.empty((nb, 2 * cols), dtype=torch.float32, device=x.device)
        _lib.check(self.lib.b200st_layernorm_bwd_partial(_dt(x), _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(add),
                                                         _p(dx), _p(partials), rows, cols, self._stream()),
                   'layernorm_bwd_partial')
        return dx, partials

    # -- multi-head attention core ----------------------------------------------------------------
    @staticmethod
    def _bld(t):
        """[B, L, HD] view with dense last dim and stride(0) == L * stride(1)."""
        assert t.dim() == 3 and t.stride(2) == 1 and (t.size(0) == 1 or t.stride(0) == t.size(1) * t.stride(1)), \
            (t.shape, t.stride())
        return t.stride(1)

    def mha_fwd(self, q, k, v, mask, n_head, temperature, want_probs=True, dropout=None):
        """dropout: None or (p, rng_state, site) -- attention dropout on the probabilities (layers.py:226)."""
        self._need_cuda(q, k, v, mask)
        B, Lq, HD = q.shape
        Lk = k.size(1)
        d = HD // n_head
        o = torch.empty((B, Lq, HD), dtype=q.dtype, device=q.device)
        p = torch.empty((B, n_head, Lq, Lk), dtype=q.dtype, device=q.device) if want_probs else None
        if mask is not None:
            assert mask.dtype in (torch.uint8, torch.bool) and mask.dim() == 3 and mask.stride(2) == 1
            assert mask.size(0) == B and mask.size(2) == Lk and mask.size(1) in (1, Lq)
            msb, msq = mask.stride(0), (mask.stride(1) if mask.size(1) == Lq and Lq > 1 else 0)
            if mask.size(1) == 1:
                msq = 0
        else:
            msb = msq = 0
        if dropout is not None and dropout[0] > 0:
            dp, rng, site = dropout
            _lib.check(self.lib.b200st_mha_fwd_dropout(_dt(q), _p(q), self._bld(q), _p(k), self._bld(k), _p(v),
                                                       self._bld(v), _p(mask), msb, msq, _p(o), HD, _p(p), B,
                                                       n_head, Lq, Lk, d, float(temperature), float(dp), _p(rng),
                                                       int(site), self._stream()), 'mha_fwd_dropout')
            return o, p
        _lib.check(self.lib.b200st_mha_fwd(_dt(q), _p(q), self._bld(q), _p(k), self._bld(k), _p(v),
                                           self._bld(v), _p(mask), msb, msq, _p(o), HD, _p(p), B,
                                           n_head, Lq, Lk, d, float(temperature), self._stream()),
                   'mha_fwd')
        return o, p

    def mha_bwd(self, dout, q, k, v, p, n_head, temperature, dq=None, dk=None, dv=None, dropout=None):
        """dq/dk/dv may be caller-provided [B, L, HD] views with a dense last dim (e.g. column slices of one fused
        [B*L, 2*HD] K|V gradient buffer)."""
        B, Lq, HD = q.shape
        Lk = k.size(1)
        d = HD // n_head
        assert dout.is_contiguous()
        ds = torch.empty_like(p)
        dq = torch.empty((B, Lq, HD), dtype=q.dtype, device=q.device) if dq is None else dq
        dk = torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device) if dk is None else dk
        dv = torch.empty((B, Lk, HD), dtype=q.dtype, device=q.device) if dv is None else dv
        if dropout is not None and dropout[0] > 0:
            dp, rng, site = dropout
            _lib.check(self.lib.b200st_mha_bwd_dropout(_dt(q), _p(dout), HD, _p(q), self._bld(q), _p(k),
                                                       self._bld(k), _p(v), self._bld(v), _p(p), _p(ds), _p(dq),
                                                       self._bld(dq), _p(dk), self._bld(dk), _p(dv), self._bld(dv), B,
                                                       n_head, Lq, Lk, d, float(temperature), float(dp), _p(rng),
                                                       int(site), self._stream()), 'mha_bwd_dropout')
            return dq, dk, dv
        _lib.check(self.lib.b200st_mha_bwd(_dt(q), _p(dout), HD, _p(q), self._bld(q), _p(k),
                                           self._bld(k), _p(v), self._bld(v), _p(p), _p(ds), _p(dq), self._bld(dq),
                                           _p(dk), self._bld(dk), _p(dv), self._bld(dv), B, n_head, Lq, Lk, d,
                                           float(temperature), self._stream()), 'mha_bwd')
        return dq, dk, dv

    def mha_decode(self, q, k_cache, v_cache, Lk, n_head, temperature, anc=None, bdiv=1, mask=None, mask_bdiv=1):
        """Attention of ONE new query per hypothesis over cached keys/values.  q [n_hyp, H*d]; k_cache / v_cache
        [slots, Lmax, H*d] views (dense last dim, arbitrary slot / position strides, e.g. the two column halves of one
        K|V cache); the first Lk positions are attended.  anc: int32 [>=Lk, n_hyp] ancestry table (slot holding
        position t of hypothesis b) or None -> slot = b // bdiv.  mask: uint8 [rows, 1, >=Lk] or [rows, >=Lk], row
        b // mask_bdiv."""
        self._need_cuda(q, k_cache, v_cache, anc, mask)
        n_hyp, HD = q.shape
        d = HD // n_head
        assert q.stride(1) == 1 and k_cache.dim() == 3 and k_cache.stride(2) == 1 and v_cache.stride(2) == 1
        assert k_cache.stride(0) == v_cache.stride(0) and k_cache.stride(1) == v_cache.stride(1)
        assert k_cache.dtype == q.dtype == v_cache.dtype and k_cache.size(1) >= Lk
        if anc is not None:
            assert anc.dtype == torch.int32 and anc.is_contiguous() and anc.size(1) == n_hyp and anc.size(0) >= Lk
        msb = 0
        if mask is not None:
            assert mask.dtype in (torch.uint8, torch.bool) and mask.stride(-1) == 1 and mask.size(-1) >= Lk
            msb = mask.stride(0)
        o = torch.empty((n_hyp, HD), dtype=q.dtype, device=q.device)
        _lib.check(self.lib.b200st_mha_decode(_dt(q), _p(q), q.stride(0), _p(k_cache), _p(v_cache), k_cache.stride(0),
                                              k_cache.stride(1), _p(anc), n_hyp, int(bdiv), _p(mask), msb,
                                              int(mask_bdiv), _p(o), HD, n_head, int(Lk), d, float(temperature),
                                              self._stream()), 'mha_decode')
        return o

    # -- LSTM -------------------------------------------------------------------------------------
    def lstm_cell_fwd(self, gates, c_prev, residual=None, save_acts=True, h_out=None, c_out=None,
                      acts_out=None, res_out=None, gates_b=None, gates_c=None):
        self._need_cuda(gates, c_prev, residual, gates_b, gates_c)
        for g in (gates_b, gates_c):
            assert g is None or (g.is_contiguous() and g.shape == gates.shape and g.dtype == gates.dtype)
        B, H4 = gates.shape
        H = H4 // 4
        assert gates.is_contiguous()
        h = h_out if h_out is not None else torch.empty((B, H), dtype=gates.dtype, device=gates.device)
        c = c_out if c_out is not None else torch.empty((B, H), dtype=torch.float32, device=gates.device)
        assert h.is_contiguous() and c.is_contiguous()
        acts = acts_out if acts_out is not None else (
            torch.empty((B, H4), dtype=torch.float32, device=gates.device) if save_acts else None)
        out_res = None
        if residual is not None:
            out_res = res_out if res_out is not None else torch.empty_like(h)
            assert residual.is_contiguous() and out_res.is_contiguous()
        _lib.check(self.lib.b200st_lstm_cell_fwd(_dt(gates), _p(gates), _p(gates_b), _p(gates_c), _p(c_prev), _p(h), _p(c),
                                                 _p(acts), _p(residual), _p(out_res), B, H,
                                                 self._stream()), 'lstm_cell_fwd')
        return h, c, acts, out_res

    def lstm_cell_bwd(self, dhs: Sequence[Optional[torch.Tensor]], dc_next, acts, c_prev, c, dtype,
                      dgates_out=None):
        dhs = [d for d in dhs if d is not None]
        assert 1 <= len(dhs) <= 3 and all(d.is_contiguous() for d in dhs)
        dhs = dhs + [None] * (3 - len(dhs))
        B, H = c.shape
        dgates = dgates_out if dgates_out is not None else torch.empty((B, 4 * H), dtype=dtype, device=c.device)
        assert dgates.is_contiguous()
        dc_prev = torch.empty((B, H), dtype=torch.float32, device=c.device)
        _lib.check(self.lib.b200st_lstm_cell_bwd(_dt(dtype), _p(dhs[0]), _p(dhs[1]), _p(dhs[2]),
                                                 _p(dc_next), _p(acts), _p(c_prev), _p(c), _p(dgates),
                                                 _p(dc_prev), B, H, self._stream()), 'lstm_cell_bwd')
        return dgates, dc_prev

    def blstm_fwd(self, xproj, w_hh_f, w_hh_r, lens, out, out_ld_t, out_ld_b, pair, save=True):
        self._need_cuda(xproj, w_hh_f, w_hh_r, lens, out)
        _, T, B, H4 = xproj.shape
        H = H4 // 4
        assert xproj.is_contiguous() and w_hh_f.is_contiguous() and w_hh_r.is_contiguous()
        assert lens.dtype == torch.int32 and out.dtype == xproj.dtype
        dev = xproj.device
        hs = torch.empty((2, T + 1, B, H), dtype=xproj.dtype, device=dev) if save else None
        # saved gates / cell states: opaque to the caller; the register-resident kernels use a blocked layout sized for
        # B rounded up to whole groups of 16 sequences (b200st_blstm_saved_layout)
        Bs = -(-B // 16) * 16 if self.blstm_saved_blocked(xproj.dtype, H) else B
        acts = torch.empty((2, T, Bs, H4), dtype=torch.float32, device=dev) if save else None
        cs = torch.empty((2, T, Bs, H), dtype=torch.float32, device=dev) if save else None
        _lib.check(self.lib.b200st_blstm_fwd(_dt(xproj), _p(xproj), _p(w_hh_f), _p(w_hh_r), _p(lens),
                                             _p(out), out_ld_t, out_ld_b, pair, _p(hs), _p(acts),
                                             _p(cs), T, B, H, self._stream()), 'blstm_fwd')
        return hs, acts, cs

    def blstm_saved_blocked(self, dtype, H) -> bool:
        return bool(self.lib.b200st_blstm_saved_layout(_dt(dtype), int(H)))

    @staticmethod
    def blstm_unblock(acts, cs, B):
        """Blocked saved state (b200st_blstm_saved_layout == 1) -> plain acts [2, T, B, 4H], cs [2, T, B, H] (tests)."""
        two, T, Bs, H4 = acts.shape
        G = Bs // 16
        # dims: dir, T, grp, rank, nt, ub, r, c, gate, seq
        a = acts.view(2, T, G, 8, 2, 4, 8, 4, 4, 2).permute(0, 1, 2, 4, 7, 9, 8, 3, 5, 6).reshape(2, T, Bs, H4)
        c = cs.view(2, T, G, 8, 2, 4, 8, 4, 2).permute(0, 1, 2, 4, 7, 8, 3, 5, 6).reshape(2, T, Bs, H4 // 4)
        return a[:, :, :B].contiguous(), c[:, :, :B].contiguous()

    @staticmethod
    def blstm_block(acts, cs):
        """Plain saved state -> the blocked layout (inverse of blstm_unblock; pads B to a multiple of 16 with zeros)."""
        two, T, B, H4 = acts.shape
        Bs = -(-B // 16) * 16
        G = Bs // 16
        a = torch.zeros((2, T, Bs, H4), dtype=acts.dtype, device=acts.device)
        c = torch.zeros((2, T, Bs, H4 // 4), dtype=cs.dtype, device=cs.device)
        a[:, :, :B] = acts
        c[:, :, :B] = cs
        # plain dims: dir, T, (grp, nt, c, seq), (gate, rank, ub, r)
        a = a.view(2, T, G, 2, 4, 2, 4, 8, 4, 8).permute(0, 1, 2, 7, 3, 8, 9, 4, 6, 5).reshape(2, T, Bs, H4)
        c = c.view(2, T, G, 2, 4, 2, 8, 4, 8).permute(0, 1, 2, 6, 3, 7, 8, 4, 5).reshape(2, T, Bs, H4 // 4)
        return a.contiguous(), c.contiguous()

    def blstm_bwd(self, dout, out_ld_t, out_ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens, dtype):
        _, T, _, H = cs.shape
        B = lens.numel()
        assert dout.is_contiguous() and acts.is_contiguous() and cs.is_contiguous()
        assert cs.shape[2] == (-(-B // 16) * 16 if self.blstm_saved_blocked(dtype, H) else B), 'saved-state layout mismatch'
        dgates = torch.empty((2, T, B, 4 * H), dtype=dtype, device=cs.device)
        _lib.check(self.lib.b200st_blstm_bwd(_dt(dtype), _p(dout), out_ld_t, out_ld_b, pair, _p(acts),
                                             _p(cs), _p(w_hh_f), _p(w_hh_r), _p(lens), _p(dgates), T, B,
                                             H, self._stream()), 'blstm_bwd')
        return dgates

    # -- LAS attention / decode helpers -------------------------------------------------------------
    def las_stack_grad(self, w, x):
        """out[b, t, :] = sum_s w[s, b, t] * x[s, b, :]  (w fp32 [S, B, Tk], x [S, B, D]) -> [B, Tk, D] in x.dtype."""
        self._need_cuda(w, x)
        assert w.dtype == torch.float32 and w.is_contiguous() and x.is_contiguous() and w.shape[:2] == x.shape[:2]
        S, B, Tk = w.shape
        D = x.size(2)
        out = torch.empty((B, Tk, D), dtype=x.dtype, device=x.device)
        _lib.check(self.lib.b200st_las_stack_grad(_dt(x), _p(w), _p(x), _p(out), S, B, Tk, D, self._stream()), 'las_stack_grad')
        return out

    def las_attn_fwd(self, q, wk, vals, klens, ctx_out=None, probs_out=None):
        self._need_cuda(q, wk, vals, klens)
        B, Tk, D = wk.shape
        Dv = vals.size(2)
        assert q.is_contiguous() and wk.is_contiguous() and vals.is_contiguous()
        ctx = ctx_out if ctx_out is not None else torch.empty((B, Dv), dtype=q.dtype, device=q.device)
        probs = probs_out if probs_out is not None else torch.empty((B, Tk), dtype=torch.float32, device=q.device)
        assert ctx.is_contiguous() and probs.is_contiguous()
        _lib.check(self.lib.b200st_las_attn_fwd(_dt(q), _p(q), _p(wk), _p(vals), _p(klens), _p(ctx),
                                                _p(probs), B, Tk, D, Dv, self._stream()),
                   'las_attn_fwd')
        return ctx, probs

    def las_attn_bwd(self, dctx, wk, vals, probs, dscore_out=None):
        B, Tk, D = wk.shape
        Dv = vals.size(2)
        assert dctx.is_contiguous() and probs.is_contiguous()
        dscore = dscore_out if dscore_out is not None else torch.empty((B, Tk), dtype=torch.float32, device=wk.device)
        assert dscore.is_contiguous()
        dq = torch.empty((B, D), dtype=wk.dtype, device=wk.device)
        _lib.check(self.lib.b200st_las_attn_bwd(_dt(wk), _p(dctx), _p(wk), _p(vals), _p(probs),
                                                _p(dscore), _p(dq), B, Tk, D, Dv, self._stream()),
                   'las_attn_bwd')
        return dscore, dq

    def las_decoder_fwd(self, wk, enc, klens, gx0, wx, whh, bias, wffn, wout, bout, CV, H, C, ACT, RES1, CTX, PROBS, LOGITS,
                        SYM, lengths, teacher):
        """The whole S-step LAS decoder loop in ONE persistent launch (csrc/las_decoder.cu; slot list in include/b200st.h).
        bf16, decoder width 512, 3 layers.  wx / whh / bias / H / C / ACT are 3-element lists; SYM int64 [S, B] contiguous."""
        tensors = [wk, enc, klens, gx0] + [t for i in range(3) for t in (wx[i], whh[i], bias[i])] + \
                  [wffn, wout, bout, CV] + [t for i in range(3) for t in (H[i], C[i], ACT[i])] + [RES1, CTX, PROBS, LOGITS, SYM, lengths]
        self._need_cuda(*tensors)
        B, Tk, D = wk.shape
        S = SYM.size(0)
        V = wout.size(0)
        bf = torch.bfloat16
        assert D == 512 and enc.shape == (B, Tk, 512) and wk.dtype == bf and enc.dtype == bf and wk.is_contiguous() and enc.is_contiguous()
        assert all(w.dtype == bf and w.stride(1) == 1 and w.shape == (2048, 512) for w in wx + whh) and all(w.is_contiguous() for w in whh)
        assert wffn.shape == (512, 1024) and wffn.is_contiguous() and wout.shape == (V, 512) and wout.is_contiguous()
        assert gx0.dtype == bf and gx0.is_contiguous() and gx0.size(-1) == 2048
        assert CV.shape == (S + 1, B, 512) and all(h.shape == (S + 1, B, 512) and h.is_contiguous() for h in H + C)
        assert SYM.dtype == torch.int64 and SYM.is_contiguous() and SYM.shape == (S, B) and lengths.dtype == torch.int32
        scratch = torch.empty(2 * B + 2, dtype=torch.int64, device=wk.device)      # best[2][B] | barrier
        ptr = lambda t: 0 if t is None else t.data_ptr()
        vals = [ptr(wk), ptr(enc), ptr(klens), ptr(gx0)]
        for i in range(3):
            vals += [ptr(wx[i]), wx[i].stride(0), ptr(whh[i]), ptr(bias[i])]
        vals += [ptr(wffn), ptr(wout), ptr(bout), ptr(CV)]
        for i in range(3):
            vals += [ptr(H[i]), ptr(C[i]), ptr(ACT[i])]
        vals += [ptr(RES1), ptr(CTX), ptr(PROBS), ptr(LOGITS), ptr(SYM), ptr(lengths), scratch.data_ptr(),
                 scratch.data_ptr() + 16 * B, B, Tk, S, V, int(bool(teacher))]
        arr = (ctypes.c_int64 * len(vals))(*vals)
        _lib.check(self.lib.b200st_las_decoder_fwd(arr, len(vals), self._stream()), 'las_decoder_fwd')
        return scratch           # keep alive until the launch is enqueued (stream-ordered allocator: safe to drop after)

    def argmax_rows(self, x, idx_out, lengths=None, step=0, embed=None, embed2=None):
        """x [rows, cols] (row-strided); idx_out: int64 1-D view (any stride) of length rows.  With `lengths`
        (int32 [rows]) the LAS decode-length rule (Dec.py:334-340) is applied in the same launch; with
        `embed = (table fp32 [cols, dim], out [rows, dim])` the chosen token's embedding row is written too, and with
        `embed2 = (table2 [cols, dim2] in x's dtype, out2 [rows, dim2])` its row of a second table."""
        self._need_cuda(x, idx_out, lengths)
        _rows(x)
        assert idx_out.dtype == torch.int64 and idx_out.dim() == 1 and idx_out.numel() == x.size(0)
        if lengths is not None:
            assert lengths.dtype == torch.int32 and lengths.is_contiguous() and lengths.numel() == x.size(0)
        table = emb = None
        ld_emb = dim = 0
        if embed is not None:
            table, emb = embed
            self._need_cuda(table, emb)
            assert table.dtype == torch.float32 and table.is_contiguous() and table.size(0) == x.size(1)
            assert emb.dtype == x.dtype and emb.dim() == 2 and emb.stride(1) == 1 and emb.shape == (x.size(0), table.size(1))
            ld_emb, dim = emb.stride(0), table.size(1)
        table2 = out2 = None
        ld2 = dim2 = 0
        if embed2 is not None:
            table2, out2 = embed2
            self._need_cuda(table2, out2)
            assert table2.dtype == x.dtype == out2.dtype and table2.is_contiguous() and table2.size(0) == x.size(1)
            assert out2.dim() == 2 and out2.stride(1) == 1 and out2.shape == (x.size(0), table2.size(1))
            ld2, dim2 = out2.stride(0), table2.size(1)
        _lib.check(self.lib.b200st_argmax_rows_embed2(
            _dt(x), _p(x), x.stride(0), x.size(0), x.size(1), _p(idx_out),
            idx_out.stride(0) if idx_out.numel() > 1 else 1, _p(lengths), int(step), _p(table), _p(emb), ld_emb, dim,
            _p(table2), _p(out2), ld2, dim2, self._stream()), 'argmax_rows')
        return idx_out

    def topk_logsoftmax(self, x, k):
        """(score fp32 [rows, k], pred int64 [rows, k]): the k largest log_softmax(x, -1) values per row and their indices
        (Seq2seq.py:254-257: log_softmax + topk in one pass); k <= 8."""
        self._need_cuda(x)
        _rows(x)
        rows, cols = x.shape
        score = torch.empty((rows, k), dtype=torch.float32, device=x.device)
        pred = torch.empty((rows, k), dtype=torch.int64, device=x.device)
        _lib.check(self.lib.b200st_topk_logsoftmax(_dt(x), _p(x), x.stride(0), rows, cols, int(k), _p(score), _p(pred),
                                                   self._stream()), 'topk_logsoftmax')
        return score, pred

    def beam_select(self, scores, cand_score, cand_pred, eos, len_map, penalty, pos, first, preds, anc, tokmask, k, done_u,
                    ticket, n_done):
        """One decode position of the beam bookkeeping, in place (Seq2seq.py:337-393; include/b200st.h)."""
        self._need_cuda(scores, cand_score, cand_pred, eos, len_map, preds, anc, tokmask, done_u, ticket, n_done)
        n_hyp = scores.numel()
        assert n_hyp % k == 0 and cand_score.shape == (n_hyp, k) and cand_pred.shape == (n_hyp, k)
        assert scores.dtype == torch.float32 and cand_score.dtype == torch.float32 and cand_pred.dtype == torch.int64
        assert cand_score.is_contiguous() and cand_pred.is_contiguous() and len_map.dtype == torch.float32
        assert eos.dtype in (torch.bool, torch.uint8) and eos.numel() == n_hyp and n_done.dtype == torch.int64
        assert preds.dtype == torch.int64 and preds.stride(1) == 1 and tokmask.dtype == torch.uint8 and tokmask.stride(1) == 1
        assert anc is None or (anc.dtype == torch.int32 and anc.is_contiguous() and anc.size(1) == n_hyp)
        assert done_u.dtype == torch.int32 and done_u.numel() >= n_hyp // k and ticket.dtype == torch.int32
        _lib.check(self.lib.b200st_beam_select(_p(scores), _p(cand_score), _p(cand_pred), _p(eos), _p(len_map), float(penalty),
                                               int(pos), int(bool(first)), _p(preds), preds.stride(0), _p(anc), _p(tokmask),
                                               tokmask.stride(0), int(k), n_hyp // k, _p(done_u), _p(ticket), _p(n_done),
                                               self._stream()), 'beam_select')

    def las_update_lengths(self, sym, lengths, step):
        assert sym.dtype == torch.int64 and sym.dim() == 1 and lengths.dtype == torch.int32
        _lib.check(self.lib.b200st_las_update_lengths(_p(sym), sym.stride(0) if sym.numel() > 1 else 1,
                                                      _p(lengths), int(step), sym.numel(),
                                                      self._stream()), 'las_update_lengths')

    # -- embeddings / mix ---------------------------------------------------------------------------
    def embedding_fwd(self, ids, table, dtype, out=None):
        self._need_cuda(ids, table, out)
        assert ids.dtype == torch.int64 and ids.is_contiguous() and table.is_contiguous()
        n, dim = ids.numel(), table.size(1)
        if out is None:
            out = torch.empty(tuple(ids.shape) + (dim,), dtype=dtype, device=table.device)
            ld = dim
        else:
            assert out.dim() == 2 and out.stride(1) == 1 and out.size(0) == n
            ld = out.stride(0)
        _lib.check(self.lib.b200st_embedding_fwd(_dt(out), _p(ids), _p(table), _p(out), ld, n, dim,
                                                 table.size(0), self._stream()), 'embedding_fwd')
        return out

    def embedding_bwd(self, ids, dout, dtable, padding_idx):
        """dtable (fp32, pre-zeroed or running) += scatter of dout rows; dout 2-D row-strided."""
        assert ids.is_contiguous() and dtable.dtype == torch.float32 and dtable.is_contiguous()
        _rows(dout)
        n, dim = ids.numel(), dtable.size(1)
        assert dout.size(0) == n and dout.size(1) == dim
        _lib.check(self.lib.b200st_embedding_bwd(_dt(dout), _p(ids), _p(dout), dout.stride(0),
                                                 _p(dtable), n, dim, dtable.size(0),
                                                 -1 if padding_idx is None else int(padding_idx),
                                                 self._stream()), 'embedding_bwd')
        return dtable

    def mix_gather_concat(self, ids, table, dyn):
        """cat[i] = [table[ids[i]], dyn[i]];  ids [n], dyn [n, D] row-strided -> [n, E + D]."""
        self._need_cuda(ids, table, dyn)
        _rows(dyn)
        n, E, D = ids.numel(), table.size(1), dyn.size(1)
        cat = torch.empty((n, E + D), dtype=dyn.dtype, device=dyn.device)
        _lib.check(self.lib.b200st_mix_gather_concat(_dt(dyn), _p(ids), _p(table), _p(dyn),
                                                     dyn.stride(0), _p(cat), n, E, D, table.size(0),
                                                     self._stream()), 'mix_gather_concat')
        return cat

    # -- softmax / loss -----------------------------------------------------------------------------
    def log_softmax_fwd(self, x, want_argmax=False):
        self._need_cuda(x)
        assert x.dim() == 2 and x.is_contiguous()
        y = torch.empty_like(x)
        am = torch.empty(x.size(0), dtype=torch.int64, device=x.device) if want_argmax else None
        _lib.check(self.lib.b200st_log_softmax_fwd(_dt(x), _p(x), _p(y), x.size(0), x.size(1), _p(am),
                                                   self._stream()), 'log_softmax_fwd')
        return y, am

    def log_softmax_bwd(self, dy, y):
        assert dy.is_contiguous() and y.is_contiguous() and dy.dim() == 2
        dx = torch.empty_like(y)
        _lib.check(self.lib.b200st_log_softmax_bwd(_dt(y), _p(dy), _p(y), _p(dx), y.size(0), y.size(1),
                                                   self._stream()), 'log_softmax_bwd')
        return dx

    def masked_nll_fwd(self, logp, target, mask):
        self._need_cuda(logp, target, mask)
        _rows(logp)
        assert target.dtype == torch.int64 and target.is_contiguous()
        if mask is not None:
            assert mask.dtype in (torch.uint8, torch.bool) and mask.is_contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=logp.device)
        _lib.check(self.lib.b200st_masked_nll_fwd(_dt(logp), _p(logp), logp.stride(0), _p(target),
                                                  _p(mask), _p(loss), logp.size(0), logp.size(1),
                                                  self._stream()), 'masked_nll_fwd')
        return loss

    def masked_nll_bwd(self, gscale, target, mask, rows, cols, dtype):
        assert gscale.dtype == torch.float32
        d = torch.empty((rows, cols), dtype=dtype, device=target.device)
        _lib.check(self.lib.b200st_masked_nll_bwd(_dt(dtype), _p(gscale), _p(target), _p(mask), _p(d),
                                                  cols, rows, cols, self._stream()), 'masked_nll_bwd')
        return d

    def softmax_nll_fused(self, logits, target, mask, scale, eps=0.0, inplace=False):
        """Returns (loss_sum[1], dlogits) with dlogits = (softmax - onehot) * mask * scale[0]."""
        self._need_cuda(logits, target, mask, scale)
        _rows(logits)
        assert scale.dtype == torch.float32 and target.is_contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
        d = logits if inplace else torch.empty((logits.size(0), logits.size(1)), dtype=logits.dtype,
                                               device=logits.device)
        _lib.check(self.lib.b200st_softmax_nll_fused(_dt(logits), _p(logits), logits.stride(0),
                                                     _p(target), _p(mask), _p(scale), float(eps),
                                                     _p(loss), _p(d), d.stride(0), logits.size(0),
                                                     logits.size(1), self._stream()),
                   'softmax_nll_fused')
        return loss, d

    # -- glue ---------------------------------------------------------------------------------------
    def add(self, a, b, out=None):
        self._need_cuda(a, b)
        assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape and a.dtype == b.dtype
        out = torch.empty_like(a) if out is None else out
        _lib.check(self.lib.b200st_add(_dt(a), _p(a), _p(b), _p(out), a.numel(), self._stream()), 'add')
        return out

    def add_posenc(self, x, pe):
        """x [B, L, D] + pe[:L] (fp32 table [>=L, D])."""
        self._need_cuda(x, pe)
        B, L, D = x.shape
        assert x.is_contiguous() and pe.is_contiguous() and pe.size(0) >= L and pe.size(1) == D
        out = torch.empty_like(x)
        _lib.check(self.lib.b200st_add_posenc(_dt(x), _p(x), _p(pe), _p(out), B, L, D, self._stream()),
                   'add_posenc')
        return out

    def transpose01(self, x, out_dtype=None):
        self._need_cuda(x)
        assert x.dim() == 3 and x.is_contiguous()
        A, Bd, C = x.shape
        out = torch.empty((Bd, A, C), dtype=out_dtype or x.dtype, device=x.device)
        _lib.check(self.lib.b200st_transpose01(_dt(x), _dt(out), _p(x), _p(out), A, Bd, C,
                                               self._stream()), 'transpose01')
        return out

    def cast(self, x, dtype, out=None):
        self._need_cuda(x)
        if x.dtype == dtype and out is None:
            return x
        assert x.is_contiguous()
        if out is None:
            out = torch.empty(x.shape, dtype=dtype, device=x.device)
        assert out.is_contiguous() and out.numel() == x.numel() and out.dtype == dtype
        _lib.check(self.lib.b200st_cast(_dt(x), _dt(out), _p(x), _p(out), x.numel(), self._stream()),
                   'cast')
        return out

    def colsum(self, x, out=None, accumulate=False):
        self._need_cuda(x)
        _rows(x)
        if out is None:
            out = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
            accumulate = False
        _lib.check(self.lib.b200st_colsum(_dt(x), _p(x), x.stride(0), _p(out), x.size(0), x.size(1),
                                          int(accumulate), self._stream()), 'colsum')
        return out

    def relu_bwd(self, dy, y):
        assert dy.is_contiguous() and y.is_contiguous()
        dx = torch.empty_like(dy)
        _lib.check(self.lib.b200st_relu_bwd(_dt(dy), _p(dy), _p(y), _p(dx), dy.numel(), self._stream()),
                   'relu_bwd')
        return dx

    def token_mask(self, ids, pad, causal):
        self._need_cuda(ids)
        assert ids.dtype == torch.int64 and ids.is_contiguous() and ids.dim() == 2
        B, L = ids.shape
        mask = torch.empty((B, L if causal else 1, L), dtype=torch.uint8, device=ids.device)
        _lib.check(self.lib.b200st_token_mask(_p(ids), _p(mask), B, L, int(pad), int(causal),
                                              self._stream()), 'token_mask')
        return mask

    def length_mask(self, lengths, L):
        self._need_cuda(lengths)
        assert lengths.dtype == torch.int32 and lengths.is_contiguous()
        B = lengths.numel()
        mask = torch.empty((B, 1, L), dtype=torch.uint8, device=lengths.device)
        _lib.check(self.lib.b200st_length_mask(_p(lengths), _p(mask), B, L, self._stream()),
                   'length_mask')
        return mask


    # -- dropout ------------------------------------------------------------------------------------
    def dropout(self, x, p, rng, site, residual=None, out=None, ld_mask=None, col_off=0):
        """out = x * keep / (1 - p) (+ residual) over a 2-D row-strided view (any leading dims are flattened when x
        is contiguous).  keep depends on (rng = device [seed, step], site, row * ld_mask + col_off + col) only; the
        backward pass is the same call on the gradient.  ld_mask / col_off let a column slice of a wider tensor
        reuse the mask of the whole (e.g. the two halves of the embedding-passing concat)."""
        self._need_cuda(x, rng, residual, out)
        x2 = x if x.dim() == 2 else x.reshape(-1, x.size(-1))
        _rows(x2)
        rows, cols = x2.shape
        if out is None:
            out = torch.empty((rows, cols), dtype=x.dtype, device=x.device)
            ret = out.view(x.shape) if x.dim() != 2 else out
        else:
            ret = out
            out = out if out.dim() == 2 else out.view(-1, out.size(-1))
        _rows(out)
        assert out.shape == x2.shape and out.dtype == x.dtype
        r2, ldr = None, 0
        if residual is not None:
            r2 = residual if residual.dim() == 2 else residual.reshape(-1, residual.size(-1))
            _rows(r2)
            assert r2.shape == x2.shape and r2.dtype == x.dtype
            ldr = r2.stride(0)
        _lib.check(self.lib.b200st_dropout(_dt(x), _p(x2), x2.stride(0), _p(r2), ldr, _p(out), out.stride(0), rows, cols,
                                           int(ld_mask if ld_mask is not None else cols), int(col_off), float(p),
                                           _p(rng), int(site), self._stream()), 'dropout')
        return ret

    def rng_advance(self, rng):
        self._need_cuda(rng)
        assert rng.dtype == torch.int64 and rng.numel() == 2
        _lib.check(self.lib.b200st_rng_advance(_p(rng), self._stream()), 'rng_advance')

    # -- input stage ------------------------------------------------------------------------------------
    def fbank_norm_pad(self, packed, offsets, lens, mu, sd, T_pad, out=None):
        """packed fp32 [N, F] (utterances back to back), offsets int64 [B], lens int32 [B], mu / sd fp32 [B, F] or None
        -> normalised, zero-padded fp32 [B, T_pad, F] (utils/dataset.py:155-184)."""
        self._need_cuda(packed, offsets, lens, mu, sd, out)
        B, F = lens.numel(), packed.size(1)
        assert packed.dtype == torch.float32 and packed.is_contiguous() and offsets.dtype == torch.int64
        assert lens.dtype == torch.int32 and (mu is None) == (sd is None)
        if mu is not None:
            assert mu.shape == (B, F) and sd.shape == (B, F) and mu.is_contiguous() and sd.is_contiguous()
            assert mu.dtype == torch.float32 and sd.dtype == torch.float32
        if out is None:
            out = torch.empty((B, T_pad, F), dtype=torch.float32, device=packed.device)
        assert out.shape == (B, T_pad, F) and out.is_contiguous() and out.dtype == torch.float32
        _lib.check(self.lib.b200st_fbank_norm_pad(_p(packed), _p(offsets), _p(lens), _p(mu), _p(sd), _p(out), B, T_pad, F,
                                                  self._stream()), 'fbank_norm_pad')
        return out

    # -- fused clip + Adam (modules/optim.py:31-36) -------------------------------------------------
    def opt_chunk(self) -> int:
        return int(self.lib.b200st_opt_chunk())

    def opt_table_cols(self) -> int:
        return int(self.lib.b200st_opt_table_cols())

    def clip_adam_step(self, tensors, table, blockmap, partials, scal, step, lr, *, max_grad_norm, beta1, beta2,
                       eps, weight_decay):
        """One optimizer step over every tensor in `table` (device int64 [n][6], see include/b200st.h); `tensors`
        (lists of params / grads / exp_avg / exp_avg_sq) is what the table points at and is only used by test
        stand-ins.  step, lr: fp32 device scalars; scal: fp32 [4] device scratch that receives
        {clip coefficient, lr / bias_correction1, sqrt(bias_correction2), grad norm}."""
        self._need_cuda(table, blockmap, partials, scal, step, lr)
        nb = blockmap.size(0)
        clip = max_grad_norm is not None and max_grad_norm > 0
        if clip:
            _lib.check(self.lib.b200st_multi_sqnorm(_p(table), _p(blockmap), nb, _p(partials), self._stream()),
                       'multi_sqnorm')
        _lib.check(self.lib.b200st_adam_prepare(_p(partials) if clip else None, nb, float(max_grad_norm or 0.0),
                                                _p(lr), float(beta1), float(beta2), _p(step), _p(scal),
                                                self._stream()), 'adam_prepare')
        _lib.check(self.lib.b200st_multi_adam(_p(table), _p(blockmap), nb, _p(scal), float(beta1), float(beta2),
                                              float(eps), float(weight_decay), self._stream()), 'multi_adam')


_backend = None


def K():
    """The kernel backend.  Created on first use; raises if libb200st.so was not built."""
    global _backend
    if _backend is None:
        _backend = CudaKernels()
    return _backend


def set_backend(obj):
    """TEST HOOK ONLY: install an object implementing the CudaKernels methods (see tests/fake_kernels.py)."""
    global _backend
    old = _backend
    _backend = obj
    return old
