"""TEST INFRASTRUCTURE ONLY — a pure-PyTorch stand-in for `b200st.kernels.CudaKernels`.

Two uses, both in tests/:
  * CPU (`-m "not gpu"`): lets the host-side orchestration (autograd functions, hand-written BPTT of the
    LAS decoder and BLSTM layers, mask plumbing, module mirror) be checked against the oracle on a
    machine without a GPU;
  * GPU (`-m gpu`): an independent per-kernel reference — every CUDA kernel is compared with the method of
    the same name here, run with torch ops on the same device.
It is never imported by the product package; `b200st.kernels.set_backend` is the only hook.
Each method restates the contract documented in include/b200st.h.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def philox_factor(rng, site, idx, p):
    """Dropout multiplier (0 or 1/(1-p)) of element indices `idx` (int64 tensor, any shape): Philox4x32-10 exactly as
    include/b200st.h documents for b200st_dropout -- key {seed_lo, seed_hi ^ step_hi}, counter {idx/4 lo, idx/4 hi,
    site, step_lo}, output word idx % 4, keep iff word >= p * 2^32.  numpy uint64 arithmetic."""
    seed, step = (int(v) for v in rng.cpu().tolist())
    M = np.uint64(0xffffffff)
    i = idx.cpu().numpy().astype(np.uint64)
    g, lane = i >> np.uint64(2), (i & np.uint64(3)).astype(np.int64)
    c0, c1 = g & M, g >> np.uint64(32)
    c2 = np.full_like(c0, np.uint64(site & 0xffffffff))
    c3 = np.full_like(c0, np.uint64(step & 0xffffffff))
    k0 = np.uint64(seed & 0xffffffff)
    k1 = np.uint64(((seed >> 32) ^ (step >> 32)) & 0xffffffff)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M, p1 >> np.uint64(32), p1 & M
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    words = np.stack([c0, c1, c2, c3], axis=-1)
    r = np.take_along_axis(words, lane[..., None], axis=-1)[..., 0]
    thresh = min(int(float(np.float32(p)) * 4294967296.0), 0xffffffff)
    keep = r >= np.uint64(thresh)
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(np.where(keep, scale, np.float32(0.0)).astype(np.float32)).to(idx.device)


class FakeKernels:
    name = 'fake-torch'

    def __init__(self):
        self.launches = 0
        self.adam_copies = {}

    def launch_count(self):
        return self.launches

    # -- GEMM -------------------------------------------------------------------------------------
    def gemm(self, a, b, *, trans_a=False, trans_b=False, out=None, out_dtype=None, bias=None,
             relu=False, residual=None, alpha=1.0, relu_gate=None):
        self.launches += 1
        A = a.float().transpose(-1, -2) if trans_a else a.float()
        Bm = b.float().transpose(-1, -2) if trans_b else b.float()
        y = alpha * torch.matmul(A, Bm)
        if bias is not None:
            y = y + bias
        if relu:
            y = torch.relu(y)
        if residual is not None:
            y = y + residual.float()
        if relu_gate is not None:
            y = torch.where(relu_gate.float() > 0, y, torch.zeros_like(y))
            residual = relu_gate
        if out is None:
            od = out_dtype or (residual.dtype if residual is not None else a.dtype)
            return y.to(od)
        out.copy_(y.to(out.dtype))
        return out

    def gemm2(self, a, b, a2, b2, *, trans_b=False, out=None, out_dtype=None, bias=None, residual=None, alpha=1.0):
        self.launches += 1
        Bm = b.float().t() if trans_b else b.float()
        B2 = b2.float().t() if trans_b else b2.float()
        y = alpha * (a.float() @ Bm + a2.float() @ B2)
        if bias is not None:
            y = y + bias
        if residual is not None:
            y = y + residual.float()
        if out is None:
            return y.to(out_dtype or (residual.dtype if residual is not None else a.dtype))
        out.copy_(y.to(out.dtype))
        return out

    # -- dropout ----------------------------------------------------------------------------------
    def dropout(self, x, p, rng, site, residual=None, out=None, ld_mask=None, col_off=0):
        self.launches += 1
        x2 = x if x.dim() == 2 else x.reshape(-1, x.size(-1))
        rows, cols = x2.shape
        ld = cols if ld_mask is None else ld_mask
        idx = torch.arange(rows, device=x.device)[:, None] * ld + col_off + torch.arange(cols, device=x.device)[None, :]
        y = x2.float() * philox_factor(rng, site, idx, p)
        if residual is not None:
            y = y + residual.reshape(rows, cols).float()
        y = y.to(x.dtype)
        if out is None:
            return y.view(x.shape) if x.dim() != 2 else y
        (out if out.dim() == 2 else out.view(rows, cols)).copy_(y)
        return out

    def rng_advance(self, rng):
        self.launches += 1
        rng[1] += 1

    # -- input stage --------------------------------------------------------------------------------
    def fbank_norm_pad(self, packed, offsets, lens, mu, sd, T_pad, out=None):
        self.launches += 1
        B, F = lens.numel(), packed.size(1)
        res = torch.zeros((B, T_pad, F), dtype=torch.float32, device=packed.device)
        for i in range(B):
            o, n = int(offsets[i]), int(lens[i])
            x = packed[o:o + n]
            if mu is not None:
                x = (x - mu[i]) / sd[i]
            res[i, :n] = x
        if out is not None:
            out.copy_(res)
            return out
        return res

    # -- fused clip + Adam ------------------------------------------------------------------------
    def opt_chunk(self):
        return 8192

    def opt_table_cols(self):
        return 7

    def clip_adam_step(self, tensors, table, blockmap, partials, scal, step, lr, *, max_grad_norm, beta1, beta2,
                       eps, weight_decay):
        """torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (amsgrad off) restated with torch ops on the tensors the
        device table points at; p.grad is left unscaled (the clip coefficient is applied on the fly)."""
        self.launches += 3
        params, grads, ms, vs = tensors
        norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
        clip = torch.ones((), device=norm.device)
        if max_grad_norm and max_grad_norm > 0:
            clip = torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0)
        step += 1
        t = float(step)
        bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
        scal.copy_(torch.stack([clip, lr[0] / bc1, torch.tensor(bc2 ** 0.5, device=norm.device), norm]))
        for p, g, m, v in zip(params, grads, ms, vs):
            g = g * clip
            if weight_decay:
                g = g + weight_decay * p.data
            m.lerp_(g, 1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (v.sqrt() / bc2 ** 0.5).add_(eps)
            p.data.addcdiv_(m, denom, value=-float(lr[0]) / bc1)
        for p, dests in getattr(self, 'adam_copies', {}).items():      # bf16 operand copies kept by the real kernel
            for d in dests:
                d.copy_(p.data.reshape(d.shape))

    # -- LayerNorm --------------------------------------------------------------------------------
    def set_gemm_sm_budget(self, n):
        return 0

    def gemm_lnbwd_ok(self, a, w, x, add=None):
        return a.dim() == 2 and w.dim() == 2 and w.size(1) == 512 and a.dtype == torch.bfloat16

    def gemm_lnbwd(self, a, w, x, gamma, mean, rstd, add=None):
        dy = a.float() @ w.float()                       # fp32 gradient of the LayerNorm output (not rounded to bf16)
        cols = x.size(-1)
        xf = x.float().reshape(-1, cols)
        xh = (xf - mean[:, None]) * rstd[:, None]
        g = dy * gamma
        dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        if add is not None:
            dx = dx + add.float().reshape(-1, cols)
        part = torch.cat([(dy * xh).sum(0), dy.sum(0)])[None, :]
        return dx.to(x.dtype), part

    def gemm_ln_ok(self, a, w, residual=None, bias=None):
        return a.dim() == 2 and w.dim() == 2 and w.size(0) == 512 and a.dtype == torch.bfloat16

    def gemm_ln(self, a, w, bias, residual, gamma, beta, eps, dropout=None):
        y = a.float() @ w.float().t()
        if bias is not None:
            y = y + bias
        if dropout is not None and dropout[0] > 0:
            y = self.dropout(y.to(a.dtype), dropout[0], dropout[1], dropout[2], residual=residual).float()
        elif residual is not None:
            y = y + residual.float()
        y = y.to(a.dtype)
        yn, mean, rstd = self.layernorm_fwd(y, gamma, beta, eps)
        return y, yn, mean, rstd

    def layernorm_fwd(self, x, gamma, beta, eps, save_stats=True):
        xf = x.float()
        mean = xf.mean(-1)
        var = xf.var(-1, unbiased=False)
        rstd = torch.rsqrt(var + eps)
        y = (xf - mean[..., None]) * rstd[..., None] * gamma + beta
        return y.to(x.dtype), mean.reshape(-1), rstd.reshape(-1)

    def layernorm_bwd(self, dy, x, gamma, mean, rstd, dgamma, dbeta, add=None):
        cols = x.size(-1)
        xf, dyf = x.float().reshape(-1, cols), dy.float().reshape(-1, cols)
        xh = (xf - mean[:, None]) * rstd[:, None]
        g = dyf * gamma
        dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        dgamma += (dyf * xh).sum(0)
        dbeta += dyf.sum(0)
        if add is not None:
            dx = dx + add.float().reshape(-1, cols)
        return dx.to(x.dtype).view(x.shape)

    def layernorm_bwd_partial(self, dy, x, gamma, mean, rstd, add=None):
        cols = x.size(-1)
        part = torch.zeros((1, 2 * cols), dtype=torch.float32, device=x.device)
        dx = self.layernorm_bwd(dy, x, gamma, mean, rstd, part[0, :cols], part[0, cols:], add=add)
        return dx, part

    # -- attention --------------------------------------------------------------------------------
    @staticmethod
    def _attn_drop(p_shape, dropout, device):
        if dropout is None or dropout[0] <= 0:
            return None
        dp, rng, site = dropout
        n = 1
        for v in p_shape:
            n *= v
        return philox_factor(rng, site, torch.arange(n, device=device).view(p_shape), dp)

    def mha_fwd(self, q, k, v, mask, n_head, temperature, want_probs=True, dropout=None):
        B, Lq, HD = q.shape
        Lk = k.size(1)
        d = HD // n_head
        qf = (q.float() / temperature).view(B, Lq, n_head, d).transpose(1, 2)
        kf = k.float().reshape(B, Lk, n_head, d).transpose(1, 2)
        vf = v.float().reshape(B, Lk, n_head, d).transpose(1, 2)
        s = torch.matmul(qf, kf.transpose(2, 3))
        if mask is not None:
            s = s.masked_fill(mask.unsqueeze(1) == 0, -1e9)
        p = torch.softmax(s, dim=-1)
        m = self._attn_drop(p.shape, dropout, q.device)
        o = torch.matmul(p if m is None else p * m, vf).transpose(1, 2).reshape(B, Lq, HD)
        return o.to(q.dtype), (p.to(q.dtype) if want_probs else None)

    def mha_bwd(self, dout, q, k, v, p, n_head, temperature, dq=None, dk=None, dv=None, dropout=None):
        outs = (dq, dk, dv)
        B, Lq, HD = q.shape
        Lk = k.size(1)
        d = HD // n_head
        do = dout.float().view(B, Lq, n_head, d).transpose(1, 2)
        qf = (q.float() / temperature).reshape(B, Lq, n_head, d).transpose(1, 2)
        kf = k.float().reshape(B, Lk, n_head, d).transpose(1, 2)
        vf = v.float().reshape(B, Lk, n_head, d).transpose(1, 2)
        pf = p.float()
        dp = torch.matmul(do, vf.transpose(2, 3))
        m = self._attn_drop(pf.shape, dropout, q.device)
        pd = pf
        if m is not None:
            dp, pd = dp * m, pf * m
        ds = pf * (dp - (dp * pf).sum(-1, keepdim=True))
        dq = torch.matmul(ds, kf) / temperature
        dk = torch.matmul(ds.transpose(2, 3), qf)
        dv = torch.matmul(pd.transpose(2, 3), do)
        back = lambda t, L: t.transpose(1, 2).reshape(B, L, HD).to(q.dtype)
        res = [back(dq, Lq), back(dk, Lk), back(dv, Lk)]
        for i, o in enumerate(outs):
            if o is not None:
                o.copy_(res[i]); res[i] = o
        return tuple(res)

    def mha_decode(self, q, k_cache, v_cache, Lk, n_head, temperature, anc=None, bdiv=1, mask=None, mask_bdiv=1):
        self.launches += 1
        n_hyp, HD = q.shape
        d = HD // n_head
        b = torch.arange(n_hyp, device=q.device)
        t = torch.arange(Lk, device=q.device)
        slot = anc[:Lk].long().t() if anc is not None else (b // bdiv)[:, None].expand(n_hyp, Lk)   # [n_hyp, Lk]
        kk = k_cache[slot, t[None, :]].float().view(n_hyp, Lk, n_head, d)
        vv = v_cache[slot, t[None, :]].float().view(n_hyp, Lk, n_head, d)
        qq = (q.float() / temperature).view(n_hyp, n_head, d)
        s = torch.einsum('bhd,bthd->bht', qq, kk)
        if mask is not None:
            m = mask.reshape(mask.size(0), -1)[:, :Lk][b // mask_bdiv]
            s = s.masked_fill(m[:, None, :] == 0, -1e9)
        p = torch.softmax(s, dim=-1)
        return torch.einsum('bht,bthd->bhd', p, vv).reshape(n_hyp, HD).to(q.dtype)

    # -- beam search bookkeeping (csrc/beam.cu), written with the reference's own op sequence (Seq2seq.py:337-393) --------
    def topk_logsoftmax(self, x, k):
        self.launches += 1
        score, pred = torch.log_softmax(x.float(), dim=-1).topk(k)
        return score.contiguous(), pred.contiguous()

    def beam_select(self, scores, cand_score, cand_pred, eos, len_map, penalty, pos, first, preds, anc, tokmask, k, done_u,
                    ticket, n_done):
        self.launches += 1
        n_hyp = scores.numel()
        B = n_hyp // k
        if first:
            scores.add_(cand_score.reshape(B, -1)[:, :k].contiguous().view(-1))
            pred_select = cand_pred.reshape(B, -1)[:, :k].contiguous().view(-1)
        else:
            eos_b = eos.bool()
            eos_exp = eos_b.reshape(-1, 1).repeat(1, k)
            eos_exp[:, 0] = False
            score_temp = scores.reshape(-1, 1) + cand_score.masked_fill(eos_b.reshape(-1, 1), 0).masked_fill(eos_exp, -1e9)
            lp = len_map.reshape(-1, 1) ** penalty
            score_temp = score_temp / lp
            score_select, sel = score_temp.reshape(B, -1).topk(k)
            scores.copy_(score_select.view(-1) * lp.view(-1))
            sel = sel + torch.arange(0, n_hyp * k, k * k, device=sel.device).reshape(B, 1)
            r_idxs, c_idxs = sel // k, sel % k
            pred_select = cand_pred[r_idxs, c_idxs].view(-1)
            rows = r_idxs.view(-1)
            preds[:, :pos] = preds[rows, :pos]
            if anc is not None:
                anc[:pos] = anc[:pos][:, rows]
            tokmask[:, :pos] = tokmask[rows, :pos]
        preds[:, pos] = pred_select
        new_eos = (pred_select == 3) | eos.bool()
        eos.copy_(new_eos.to(eos.dtype))
        len_map.add_(torch.ones_like(len_map).masked_fill(new_eos, 0))
        n_done.copy_(new_eos.sum())

    # -- LSTM cell --------------------------------------------------------------------------------
    def lstm_cell_fwd(self, gates, c_prev, residual=None, save_acts=True, h_out=None, c_out=None,
                      acts_out=None, res_out=None, gates_b=None, gates_c=None):
        pre = gates.float()
        for extra in (gates_b, gates_c):
            if extra is not None:
                pre = pre + extra.float()
        i, f, g, o = pre.chunk(4, dim=-1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        cp = 0 if c_prev is None else c_prev
        c = f * cp + i * g
        h = (o * torch.tanh(c))
        acts = torch.cat([i, f, g, o], -1)
        hq = h.to(gates.dtype)
        if h_out is not None:
            h_out.copy_(hq); hq = h_out
        if c_out is not None:
            c_out.copy_(c); c = c_out
        if acts_out is not None:
            acts_out.copy_(acts); acts = acts_out
        out_res = None
        if residual is not None:
            out_res = (h + residual.float()).to(gates.dtype)
            if res_out is not None:
                res_out.copy_(out_res); out_res = res_out
        return hq, c, acts, out_res

    def lstm_cell_bwd(self, dhs, dc_next, acts, c_prev, c, dtype, dgates_out=None):
        dh = sum(d.float() for d in dhs if d is not None)
        i, f, g, o = acts.chunk(4, dim=-1)
        tc = torch.tanh(c)
        cp = 0 if c_prev is None else c_prev
        dc = (0 if dc_next is None else dc_next) + dh * o * (1 - tc * tc)
        dg = torch.cat([dc * g * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - g * g),
                        dh * tc * o * (1 - o)], -1).to(dtype)
        if dgates_out is not None:
            dgates_out.copy_(dg); dg = dgates_out
        return dg, dc * f

    # -- persistent BLSTM recurrence ----------------------------------------------------------------
    @staticmethod
    def _out_index(out, t, pair, H, d, out_ld_t, out_ld_b, B):
        """Flat element offsets of out[(t/pair)*ld_t + b*ld_b + (t%pair)*2H + d*H + u] for all b,u."""
        b = torch.arange(B, device=out.device)[:, None]
        u = torch.arange(H, device=out.device)[None, :]
        return (t // pair) * out_ld_t + b * out_ld_b + (t % pair) * 2 * H + d * H + u

    def blstm_fwd(self, xproj, w_hh_f, w_hh_r, lens, out, out_ld_t, out_ld_b, pair, save=True):
        _, T, B, H4 = xproj.shape
        H = H4 // 4
        dev = xproj.device
        hs = torch.zeros((2, T + 1, B, H), dtype=xproj.dtype, device=dev)
        acts = torch.zeros((2, T, B, H4), dtype=torch.float32, device=dev)
        cs = torch.zeros((2, T, B, H), dtype=torch.float32, device=dev)
        flat = out.view(-1)
        ln = lens.to(dev).long()
        for d, w in ((0, w_hh_f), (1, w_hh_r)):
            h = torch.zeros(B, H, device=dev)
            c = torch.zeros(B, H, device=dev)
            for s in range(T):
                t = T - 1 - s if d else s
                gates = xproj[d, t].float() + h @ w.t()
                i, f, g, o = gates.chunk(4, -1)
                i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
                valid = (t < ln)[:, None]
                c_new = torch.where(valid, f * c + i * g, c)
                h_new = torch.where(valid, o * torch.tanh(c_new), h)
                h, c = h_new, c_new
                acts[d, t] = torch.cat([i, f, g, o], -1)
                cs[d, t] = c
                ho = torch.where(valid, h, torch.zeros_like(h)).to(xproj.dtype)
                flat[self._out_index(out, t, pair, H, d, out_ld_t, out_ld_b, B).view(-1)] = ho.view(-1)
                hs[d, t if d else t + 1] = ho
        if not save:
            return None, None, None
        return hs, acts, cs

    def blstm_bwd(self, dout, out_ld_t, out_ld_b, pair, acts, cs, w_hh_f, w_hh_r, lens, dtype):
        _, T, B, H = cs.shape
        dev = cs.device
        dg_all = torch.zeros((2, T, B, 4 * H), dtype=dtype, device=dev)
        flat = dout.reshape(-1)
        ln = lens.to(dev).long()
        for d, w in ((0, w_hh_f), (1, w_hh_r)):
            dhrec = torch.zeros(B, H, device=dev)
            dcrec = torch.zeros(B, H, device=dev)
            for s in range(T):
                t = s if d else T - 1 - s
                valid = (t < ln)[:, None]
                i, f, g, o = acts[d, t].chunk(4, -1)
                c_t = cs[d, t]
                tp = t + 1 if d else t - 1
                c_prev = cs[d, tp] if 0 <= tp < T else torch.zeros_like(c_t)
                do = flat[self._out_index(dout, t, pair, H, d, out_ld_t, out_ld_b, B).view(-1)].view(B, H).float()
                dh = do + dhrec
                tc = torch.tanh(c_t)
                dc = dcrec + dh * o * (1 - tc * tc)
                dg = torch.cat([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g),
                                dh * tc * o * (1 - o)], -1)
                dg = torch.where(valid, dg, torch.zeros_like(dg))
                dcrec = torch.where(valid, dc * f, torch.zeros_like(dc))
                dg_all[d, t] = dg.to(dtype)
                dhrec = dg @ w
        return dg_all

    # -- LAS attention / decode helpers -------------------------------------------------------------
    def las_stack_grad(self, w, x):
        return torch.einsum('sbt,sbd->btd', w.float(), x.float()).to(x.dtype)

    def las_attn_fwd(self, q, wk, vals, klens, ctx_out=None, probs_out=None):
        B, Tk, D = wk.shape
        s = torch.bmm(q.float().unsqueeze(1), wk.float().transpose(1, 2)).squeeze(1)
        if klens is not None:
            m = torch.arange(Tk, device=q.device)[None, :] >= klens.long()[:, None]
            s = s.masked_fill(m, -1e12)
        p = torch.softmax(s, dim=1)
        c = torch.bmm(p.unsqueeze(1), vals.float()).squeeze(1).to(q.dtype)
        if ctx_out is not None:
            ctx_out.copy_(c); c = ctx_out
        if probs_out is not None:
            probs_out.copy_(p); p = probs_out
        return c, p

    def las_attn_bwd(self, dctx, wk, vals, probs, dscore_out=None):
        dp = torch.bmm(dctx.float().unsqueeze(1), vals.float().transpose(1, 2)).squeeze(1)
        ds = probs * (dp - (dp * probs).sum(1, keepdim=True))
        dq = torch.bmm(ds.unsqueeze(1), wk.float()).squeeze(1).to(wk.dtype)
        if dscore_out is not None:
            dscore_out.copy_(ds); ds = dscore_out
        return ds, dq

    def argmax_rows(self, x, idx_out, lengths=None, step=0, embed=None, embed2=None):
        idx_out.copy_(x.float().argmax(dim=1))
        if lengths is not None:
            self.las_update_lengths(idx_out, lengths, step)
        if embed is not None:
            table, out = embed
            out.copy_(table[idx_out].to(out.dtype))
        if embed2 is not None:
            table2, out2 = embed2
            out2.copy_(table2[idx_out])
        return idx_out

    def las_update_lengths(self, sym, lengths, step):
        ended = ((sym == 3) | (sym == 0)) & (lengths > step)
        lengths[ended] = step + 1

    # -- embeddings / mix ---------------------------------------------------------------------------
    def embedding_fwd(self, ids, table, dtype, out=None):
        e = table[ids].to(dtype)
        if out is None:
            return e
        out.copy_(e.view(out.shape))
        return out

    def embedding_bwd(self, ids, dout, dtable, padding_idx):
        keep = ids != (padding_idx if padding_idx is not None else -1)
        dtable.index_add_(0, ids[keep], dout.float()[keep])
        return dtable

    def mix_gather_concat(self, ids, table, dyn):
        return torch.cat([table[ids].to(dyn.dtype), dyn], dim=1)

    # -- softmax / loss -----------------------------------------------------------------------------
    def log_softmax_fwd(self, x, want_argmax=False):
        y = torch.log_softmax(x.float(), dim=1).to(x.dtype)
        return y, (x.float().argmax(1) if want_argmax else None)

    def log_softmax_bwd(self, dy, y):
        dyf = dy.float()
        return (dyf - torch.exp(y.float()) * dyf.sum(1, keepdim=True)).to(y.dtype)

    def masked_nll_fwd(self, logp, target, mask):
        per = -logp.float().gather(1, target[:, None]).squeeze(1)
        if mask is not None:
            per = per * mask.to(per.dtype)
        return per.sum().reshape(1)

    def masked_nll_bwd(self, gscale, target, mask, rows, cols, dtype):
        d = torch.zeros((rows, cols), dtype=torch.float32, device=target.device)
        v = -gscale.expand(rows).clone()
        if mask is not None:
            v = v * mask.to(v.dtype)
        d.scatter_(1, target[:, None], v[:, None])
        return d.to(dtype)

    def softmax_nll_fused(self, logits, target, mask, scale, eps=0.0, inplace=False):
        x = logits.float()
        V = x.size(1)
        lse = torch.logsumexp(x, dim=1)
        q = torch.full_like(x, eps / V)
        q.scatter_add_(1, target[:, None], torch.full_like(x[:, :1], 1 - eps))
        per = lse - (q * x).sum(1)
        m = torch.ones_like(per) if mask is None else mask.to(per.dtype)
        loss = (per * m).sum().reshape(1)
        d = ((torch.softmax(x, 1) - q) * m[:, None] * scale).to(logits.dtype)
        if inplace:
            logits.copy_(d); d = logits
        return loss, d

    # -- glue ---------------------------------------------------------------------------------------
    def add(self, a, b, out=None):
        y = (a.float() + b.float()).to(a.dtype)
        if out is None:
            return y
        out.copy_(y)
        return out

    def add_posenc(self, x, pe):
        return (x.float() + pe[:x.size(1)]).to(x.dtype)

    def transpose01(self, x, out_dtype=None):
        return x.transpose(0, 1).contiguous().to(out_dtype or x.dtype)

    def cast(self, x, dtype, out=None):
        if out is None:
            return x.to(dtype)
        out.copy_(x.to(dtype))
        return out

    def colsum(self, x, out=None, accumulate=False):
        s = x.float().sum(0)
        if out is None:
            return s
        if accumulate:
            out += s
        else:
            out.copy_(s)
        return out

    def relu_bwd(self, dy, y):
        return torch.where(y.float() > 0, dy, torch.zeros_like(dy))

    def token_mask(self, ids, pad, causal):
        B, L = ids.shape
        m = (ids != pad).unsqueeze(1)
        if causal:
            m = m & torch.tril(torch.ones(L, L, dtype=torch.bool, device=ids.device)).unsqueeze(0)
        return m.to(torch.uint8).contiguous()

    def length_mask(self, lengths, L):
        return (torch.arange(L, device=lengths.device)[None, :] < lengths.long()[:, None]).unsqueeze(1).to(torch.uint8)
