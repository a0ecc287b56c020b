"""The persistent one-launch LAS decoder loop (csrc/las_decoder.cu, b200st_las_decoder_fwd) against (1) the step-by-step
kernels it replaces and (2) the fp32 CPU oracle (oracle.st_oracle.las_decoder = Dec.forward, Dec.py:130-233,320-438)."""
import pytest
import torch

from b200st import runtime
from conftest import rel_err
from helpers import build_model
from oracle import st_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _bf16():
    runtime.set_compute_dtype('bf16')
    yield
    runtime.las_persistent(True)
    runtime.set_compute_dtype('fp32')


def _case(B, Tk, V=304, E=24, S=9, seed=1):
    cfg = O.STConfig(enc_vocab_size=V, dec_vocab_size=V, enc_embedding_size=E, dec_embedding_size=E, max_seq_len_src=S + 1,
                     max_seq_len_tgt=8, num_heads=8, dim_model=512, dim_feedforward=64, enc_layers=1, dec_layers=1,
                     acous_dim=16, acous_hidden_size=256)
    P = O.init_params(cfg, seed=seed, scale=1.5)
    g = torch.Generator().manual_seed(seed + 1)
    enc = 0.5 * torch.randn(B, Tk, 512, generator=g)
    lens8 = torch.randint(max(1, Tk // 2), Tk + 1, (B,), generator=g)
    lens8[0] = Tk
    tgt = torch.randint(5, V, (B, S + 1), generator=g)
    tgt[:, 0] = 2
    m = build_model(cfg, P, device='cuda')
    return cfg, P, m, enc, lens8, tgt


def _run(m, enc, lens8, tgt, teacher, persistent, need_grad=False):
    runtime.las_persistent(persistent)
    dec = m.las.decoder
    x = enc.cuda().to(torch.bfloat16).requires_grad_(need_grad)
    import random
    random.seed(0)
    embs, logps, syms, lengths = dec.forward_device(x, lens8.to(torch.int32).cuda(), tgt=tgt.cuda() if teacher else None,
                                                    teacher_forcing_ratio=1.0 if teacher else 0.0, need_logps=True)
    return x, embs, logps, syms.squeeze(-1), lengths


def _oracle(cfg, P, enc, lens8, tgt, teacher):
    acous_lens = [int(n) * 8 - 1 for n in lens8]           # padded_len(n) / 8 == lens8
    e, lp, sy, ln = O.las_decoder(P, cfg, enc, acous_lens, tgt=tgt if teacher else None, teacher_forcing=teacher)
    return e, lp, sy.squeeze(-1), ln


@pytest.mark.parametrize('B,Tk', [(16, 12), (64, 126), (80, 37)])
def test_teacher_forced_matches_stepwise_and_oracle(B, Tk):
    cfg, P, m, enc, lens8, tgt = _case(B, Tk)
    with torch.no_grad():
        _, e1, l1, s1, n1 = _run(m, enc, lens8, tgt, True, True)
        _, e0, l0, s0, n0 = _run(m, enc, lens8, tgt, True, False)
    eo, lo, so, no = _oracle(cfg, P, enc.to(torch.bfloat16).float(), lens8, tgt, True)
    assert rel_err(e1.float().cpu(), e0.float().cpu()) < 1e-2 and rel_err(l1.float().cpu(), l0.float().cpu()) < 1e-2
    assert rel_err(e1.float().cpu(), eo) < 2e-2 and rel_err(l1.float().cpu(), lo) < 2e-2
    assert float((s1.cpu() == so).float().mean()) > 0.95
    # lengths follow the kernel's own symbols by the Dec.decode rule (Dec.py:334-340)
    S = s1.size(1)
    for b in range(B):
        hit = ((s1[b] == 3) | (s1[b] == 0)).nonzero()
        assert int(n1[b]) == (int(hit[0]) + 1 if hit.numel() else S + 1)


def test_free_running_feedback():
    cfg, P, m, enc, lens8, tgt = _case(32, 40, V=1000, S=12, seed=5)
    with torch.no_grad():
        _, e1, l1, s1, n1 = _run(m, enc, lens8, None, False, True)
    eo, lo, so, no = _oracle(cfg, P, enc.to(torch.bfloat16).float(), lens8, None, False)
    same = s1.cpu() == so
    assert float(same.float().mean()) > 0.8, float(same.float().mean())
    rows = same.all(dim=1)                                  # rows that never flipped a near tie follow the oracle to 2e-2
    assert int(rows.sum()) >= 16
    assert rel_err(e1.float().cpu()[rows], eo[rows]) < 2e-2
    # feeding the kernel's OWN symbols back as teacher-forcing tokens must reproduce its cell values exactly
    ids = torch.cat([torch.full((s1.size(0), 1), 2, dtype=torch.int64), s1.cpu()], dim=1)
    with torch.no_grad():
        _, e2, _, s2, _ = _run(m, enc, lens8, ids, True, True)
    assert torch.equal(s2, s1) and rel_err(e2.float(), e1.float()) < 1e-3


def test_backward_through_persistent_forward_matches_stepwise():
    cfg, P, m, enc, lens8, tgt = _case(16, 20)
    grads = {}
    for persistent in (True, False):
        m.zero_grad(set_to_none=True)
        x, embs, logps, _, _ = _run(m, enc, lens8, tgt, True, persistent, need_grad=True)
        w = torch.linspace(-1, 1, embs.numel(), device='cuda').view_as(embs)
        (embs.float() * w).sum().backward()
        grads[persistent] = {n: p.grad.detach().float().clone() for n, p in m.las.decoder.named_parameters() if p.grad is not None}
        grads[persistent]['enc'] = x.grad.detach().float().clone()
    assert set(grads[True]) == set(grads[False]) and len(grads[True]) > 10
    for n in grads[True]:
        assert rel_err(grads[True][n], grads[False][n]) < 2e-2, n
