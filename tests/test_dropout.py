"""Dropout > 0 in training mode (reference defaults: dropout 0.2, attention dropout 0.1 hard-wired at layers.py:207).

A dropout mask cannot be bit-compatible with PyTorch's generator, so parity is established by MASK INJECTION: the
product draws its counter-based masks (csrc/philox.cuh), the test reads back the mask of every site through the same
C-ABI call on a tensor of ones, and the CPU oracle (the reference's algorithm) is run with exactly those masks at the
reference's own nn.Dropout call sites.  Loss and every gradient must then agree to the fp32 contract (1e-4)."""
import re

import numpy as np
import pytest
import torch

from b200st import kernels, runtime as rt
from conftest import rel_err
from fake_kernels import FakeKernels, philox_factor
from helpers import build_model, train_step
from oracle import st_oracle as O

CFG = dict(enc_vocab_size=120, dec_vocab_size=120, enc_embedding_size=24, dec_embedding_size=24, max_seq_len_src=7,
           max_seq_len_tgt=9, num_heads=4, dim_model=32, dim_feedforward=48, enc_layers=2, dec_layers=2, acous_dim=12,
           acous_hidden_size=16)


def _set_dropout(model, p=0.2, p_emb=0.15, p_attn=0.1):
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Dropout):
            if name.endswith('embedding_dropout'):
                mod.p = p_emb
            elif name.endswith('attention.dropout'):
                mod.p = p_attn
            else:
                mod.p = p


def _mask_from_product(tag, shape, log, rng, device):
    """The multiplier tensor (0 or 1/(1-p)) of site `tag`, in the ORACLE's layout `shape`."""
    site, p = log[tag]
    k = kernels.K()

    def draw(rows, cols):
        return k.dropout(torch.ones(rows, cols, device=device), p, rng, site).cpu()
    m = re.fullmatch(r'las\.enc\.l(\d)', tag)
    if m and int(m.group(1)) < 4:            # product layout [T/2, B, 2 * 2H] (frame pairs already concatenated)
        B, T, H2 = shape
        return draw(T // 2 * B, 2 * H2).view(T // 2, B, 2, H2).permute(1, 0, 2, 3).reshape(B, T, H2)
    if tag == 'las.dec.emb':                 # product: [B, E] (free running: BOS row only) or [S, B, E] (teacher forcing)
        B, S_full, E = shape
        out = torch.ones(shape)
        if log.get('_las_teacher_forcing'):
            out[:, :S_full - 1] = draw((S_full - 1) * B, E).view(S_full - 1, B, E).permute(1, 0, 2)
        else:
            out[:, 0] = draw(B, E)
        return out
    n_last = shape[-1]
    return draw(int(np.prod(shape[:-1])), n_last).view(shape)


def _run_case(device):
    cfg = O.STConfig(**CFG)
    P = O.init_params(cfg, seed=21, scale=2.0)
    data = O.synthetic_batch(cfg, batch=3, frames=40, seed=22, ragged=True)
    model = build_model(cfg, P, device=device)
    _set_dropout(model)
    model.train()
    rt.manual_seed(77)
    rt.site_log = {}
    try:
        loss, out = train_step(model, data, device)
        loss.backward()
        log = dict(rt.site_log)
    finally:
        rt.site_log = None
    rng = rt.current_rng(torch.device(device) if device != 'cuda' else torch.device('cuda', torch.cuda.current_device()))
    # every reference dropout site was visited: 4 encoder layers, BOS embedding, (3 layers + context) x S LAS steps,
    # mix, target embedding, and per Transformer sub-layer attn + fc (+ ffn)
    S = cfg.max_seq_len_src - 1
    assert len(log) == 4 + 1 + 4 * S + 2 + cfg.enc_layers * 3 + cfg.dec_layers * 5, sorted(log)
    used = set()

    def drop(x, tag):
        used.add(tag)
        return x * _mask_from_product(tag, tuple(x.shape), log, rng, device).to(x.dtype)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    O.DROP = drop
    try:
        loss_ref, _ = O.train_step_st(Pg, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        loss_ref.backward()
    finally:
        O.DROP = None
    assert used == set(log)
    # and the masks do something: the same model without dropout gives a different loss
    loss0, _ = O.train_step_st(P, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
    assert abs(float(loss0) - float(loss_ref)) > 1e-3 * abs(float(loss_ref))
    assert abs(loss.get_loss() - float(loss_ref)) < 1e-4 * abs(float(loss_ref)), (loss.get_loss(), float(loss_ref))
    named = dict(model.named_parameters())
    gnorm = sum(float(v.grad.double().norm() ** 2) for v in Pg.values() if v.grad is not None) ** 0.5
    for name, v in Pg.items():
        if v.grad is None or float(v.grad.abs().sum()) == 0.0:
            continue
        got = named[name].grad
        assert got is not None, name
        err = float((got.double().cpu() - v.grad.double()).norm())
        assert err <= 1e-4 * max(float(v.grad.double().norm()), 1e-3 * gnorm), (name, err, float(v.grad.norm()))


def _run_las_teacher_forcing(device):
    """LAS.forward in training mode with teacher forcing: embedding dropout covers every input token (Dec.py:166)."""
    cfg = O.STConfig(**CFG)
    P = O.init_params(cfg, seed=31, scale=2.0)
    data = O.synthetic_batch(cfg, batch=3, frames=32, seed=32)
    model = build_model(cfg, P, device=device)
    _set_dropout(model)
    model.train()
    rt.manual_seed(5)
    rt.site_log = {}
    try:
        lens = [torch.tensor([n]) for n in data['acous_lens']]
        embs, logps, _, _ = model.las(data['acous_feats'].to(device), acous_lens=lens, tgt=data['src'].to(device),
                                      is_training=False, teacher_forcing_ratio=1.0, use_gpu=device != 'cpu')
        w = torch.randn(logps.shape, generator=torch.Generator().manual_seed(1))
        obj = (logps.float() * w.to(device)).sum() + embs.float().sum()
        obj.backward()
        log = dict(rt.site_log)
    finally:
        rt.site_log = None
    log['_las_teacher_forcing'] = True
    rng = rt.current_rng(embs.device)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    O.DROP = lambda x, tag: x * _mask_from_product(tag, tuple(x.shape), log, rng, device).to(x.dtype)
    try:
        enc = O.las_encoder(Pg, cfg, data['acous_feats'], data['acous_lens'])
        e_ref, lp_ref, _, _ = O.las_decoder(Pg, cfg, enc, data['acous_lens'], tgt=data['src'], teacher_forcing=True)
        ((lp_ref * w).sum() + e_ref.sum()).backward()
    finally:
        O.DROP = None
    assert rel_err(logps.cpu(), lp_ref) < 1e-4
    named = dict(model.named_parameters())
    for name, v in Pg.items():
        if v.grad is None or float(v.grad.abs().sum()) == 0.0:
            continue
        assert rel_err(named[name].grad.cpu(), v.grad) < 2e-4, name


def test_dropout_step_vs_oracle_with_injected_masks_cpu():
    old = kernels.set_backend(FakeKernels())
    try:
        _run_case('cpu')
    finally:
        kernels.set_backend(old)


def test_las_teacher_forcing_dropout_cpu():
    old = kernels.set_backend(FakeKernels())
    try:
        _run_las_teacher_forcing('cpu')
    finally:
        kernels.set_backend(old)


def test_eval_mode_ignores_dropout_cpu():
    old = kernels.set_backend(FakeKernels())
    try:
        cfg = O.STConfig(**CFG)
        P = O.init_params(cfg, seed=21)
        data = O.synthetic_batch(cfg, batch=2, frames=24, seed=3)
        model = build_model(cfg, P, device='cpu')
        _set_dropout(model)
        model.eval()
        rt.site_log = {}
        loss, _ = train_step(model, data, 'cpu')
        assert rt.site_log == {}
        rt.site_log = None
        ref, _ = O.train_step_st(P, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        assert abs(loss.get_loss() - float(ref)) < 1e-4 * abs(float(ref))
    finally:
        rt.site_log = None
        kernels.set_backend(old)


# ---- on the B200 -----------------------------------------------------------------------------------------------
@pytest.fixture
def fp32():
    rt.set_compute_dtype('fp32')
    yield
    rt.set_compute_dtype('fp32')


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('rows,cols', [(64, 512), (3, 24), (37, 29), (1, 7), (1000, 712)])
def test_dropout_kernel_matches_philox_restatement(dtype, rows, cols):
    k, f = kernels.CudaKernels(), FakeKernels()
    rng = torch.tensor([0x1234567890abcdef & 0x7fffffffffffffff, 41], dtype=torch.int64, device='cuda')
    x = torch.randn(rows, cols, device='cuda').to(dtype)
    res = torch.randn(rows, cols, device='cuda').to(dtype)
    for p in (0.1, 0.5):
        for site in (1, 77):
            assert torch.equal(k.dropout(x, p, rng, site), f.dropout(x, p, rng, site))
            got, ref = k.dropout(x, p, rng, site, residual=res), f.dropout(x, p, rng, site, residual=res)
            assert rel_err(got, ref) < (1e-6 if dtype == torch.float32 else 1e-2)
    # a column slice of a wider mask (the two halves of the embedding-passing concat, Seq2seq.py:188-195)
    wide = torch.randn(rows, cols + 8, device='cuda').to(dtype)
    full = k.dropout(wide, 0.3, rng, 5)
    part = k.dropout(wide[:, 8:], 0.3, rng, 5, ld_mask=cols + 8, col_off=8)
    assert torch.equal(part, full[:, 8:])
    assert torch.equal(part, f.dropout(wide[:, 8:], 0.3, rng, 5, ld_mask=cols + 8, col_off=8))


@pytest.mark.gpu
def test_dropout_statistics_and_step_advance():
    k = kernels.CudaKernels()
    rng = torch.tensor([99, 0], dtype=torch.int64, device='cuda')
    ones = torch.ones(4096, 512, device='cuda')
    a = k.dropout(ones, 0.2, rng, 3)
    keep = float((a != 0).float().mean())
    assert abs(keep - 0.8) < 2e-3 and abs(float(a.mean()) - 1.0) < 3e-3
    assert not torch.equal(a, k.dropout(ones, 0.2, rng, 4))          # another site: another mask
    k.rng_advance(rng)
    assert rng.tolist() == [99, 1]
    assert not torch.equal(a, k.dropout(ones, 0.2, rng, 3))          # another step: another mask


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,H,Lq,Lk,d', [(3, 4, 9, 9, 8), (2, 8, 50, 31, 64), (2, 8, 50, 50, 64)])
def test_mha_attention_dropout_matches_restatement(dtype, B, H, Lq, Lk, d):
    k, f = kernels.CudaKernels(), FakeKernels()
    g = torch.Generator().manual_seed(3)
    q, kk, v, do = (torch.randn(B, L, H * d, generator=g).cuda().to(dtype) for L in (Lq, Lk, Lk, Lq))
    mask = (torch.rand(B, 1, Lk, generator=g) > 0.2).to(torch.uint8).cuda()
    mask[:, :, 0] = 1
    rng = torch.tensor([7, 3], dtype=torch.int64, device='cuda')
    drop = (0.1, rng, 12)
    tol = 2e-5 if dtype == torch.float32 else 3e-2
    o, p = k.mha_fwd(q, kk, v, mask, H, d ** 0.5, dropout=drop)
    o_r, p_r = f.mha_fwd(q, kk, v, mask, H, d ** 0.5, dropout=drop)
    assert rel_err(o, o_r) < tol and rel_err(p, p_r) < tol
    o0, _ = k.mha_fwd(q, kk, v, mask, H, d ** 0.5)
    assert rel_err(o0, o_r) > 10 * tol                                # the mask is really applied
    got = k.mha_bwd(do, q, kk, v, p, H, d ** 0.5, dropout=drop)
    ref = f.mha_bwd(do, q, kk, v, p_r, H, d ** 0.5, dropout=drop)
    for a, b in zip(got, ref):
        assert rel_err(a, b) < tol


@pytest.mark.gpu
def test_dropout_step_vs_oracle_with_injected_masks_gpu(fp32):
    _run_case('cuda')


@pytest.mark.gpu
def test_las_teacher_forcing_dropout_gpu(fp32):
    _run_las_teacher_forcing('cuda')


@pytest.mark.gpu
def test_dropout_train_step_bf16_graph_replays_draw_new_masks():
    """bf16 + whole-step CUDA graph with the reference's default dropouts: replays give different (fresh-mask) losses
    that stay close to the no-dropout loss, and gradients are finite."""
    from b200st.graph import GraphedTrainStep
    from b200st.train_step import Trainer_ST
    rt.set_compute_dtype('bf16')
    try:
        cfg = O.STConfig(enc_vocab_size=300, dec_vocab_size=300, enc_embedding_size=40, dec_embedding_size=40,
                         max_seq_len_src=10, max_seq_len_tgt=13, num_heads=2, dim_model=128, dim_feedforward=256,
                         enc_layers=2, dec_layers=2, acous_dim=24, acous_hidden_size=256)
        P = O.init_params(cfg, seed=11)
        data = O.synthetic_batch(cfg, 16, 96, seed=21)
        m = build_model(cfg, P, device='cuda')
        _set_dropout(m)
        m.train()
        items = {'srcid': [data['src'].cuda()], 'tgtid': [data['tgt'].cuda()], 'acous_feat': [data['acous_feats'].cuda()],
                 'acouslen': [int(n) for n in data['acous_lens']]}
        g = GraphedTrainStep(m, Trainer_ST(use_gpu=True, batch_size=16), items)
        losses = [float(g()) for _ in range(4)]
        assert len(set(losses)) == 4, losses
        m.eval()
        ref, _ = O.train_step_st(P, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        for v in losses:
            assert abs(v - float(ref)) < 0.2 * abs(float(ref)), (losses, float(ref))
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    finally:
        rt.set_compute_dtype('fp32')


@pytest.mark.gpu
def test_dropout_fused_gemm_ln_path_matches_separate_kernels_bf16():
    """d_model = 512, bf16, the reference's default dropouts: with the sub-layer-closing GEMMs running dropout + skip +
    the next LayerNorm in their epilogue (csrc/gemm_ln.cu) the step gives the same loss and gradients as with the separate
    GEMM -> dropout(+residual) -> LayerNorm kernels, mask for mask (same seed, same sites)."""
    cfg = O.STConfig(enc_vocab_size=304, dec_vocab_size=304, enc_embedding_size=24, dec_embedding_size=24, max_seq_len_src=8,
                     max_seq_len_tgt=11, num_heads=8, dim_model=512, dim_feedforward=128, enc_layers=2, dec_layers=2,
                     acous_dim=16, acous_hidden_size=256)
    P = O.init_params(cfg, seed=3)
    data = O.synthetic_batch(cfg, batch=8, frames=64, seed=4, ragged=True)
    k = kernels.K()
    rt.set_compute_dtype('bf16')
    try:
        results = []
        for fused in (True, False):
            model = build_model(cfg, P, device='cuda')
            _set_dropout(model)
            model.train()
            rt.manual_seed(123)
            calls = {'n': 0}
            ok0 = k.gemm_ln_ok
            if fused:
                k.gemm_ln_ok = lambda *a, **kw: (calls.__setitem__('n', calls['n'] + 1), ok0(*a, **kw))[1]
            else:
                k.gemm_ln_ok = lambda *a, **kw: False
            try:
                loss, out = train_step(model, data, 'cuda')
                loss.backward()
            finally:
                del k.gemm_ln_ok
            torch.cuda.synchronize()
            assert (calls['n'] > 0) == fused
            results.append((loss.get_loss(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}))
    finally:
        rt.set_compute_dtype('fp32')
    (l1, g1), (l0, g0) = results
    assert abs(l1 - l0) < 5e-3 * abs(l0), (l1, l0)
    gn = sum(float(v.double().norm() ** 2) for v in g0.values()) ** 0.5
    for n, v in g0.items():
        if n.startswith('las.'):
            continue          # (the LAS forward is identical in both runs; its gradients are the Transformer's passed through BPTT)
        err = float((g1[n].double() - v.double()).norm())
        assert err <= 3e-2 * max(float(v.double().norm()), 1e-3 * gn), (n, err, float(v.norm()))
