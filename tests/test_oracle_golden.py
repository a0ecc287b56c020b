"""Pin the oracle (oracle/st_oracle.py) against outputs of the real reference (tests/golden/*.npz).

CPU only.  Tolerances: fp32 on both sides, same torch primitives, so agreement is ~1e-6; the contract
tolerance for the CUDA path (1e-4 relative, BASELINE.json north_star) is asserted in tests/test_gpu_*.py."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import st_oracle as O

TOL = 2e-5


def test_st_forward_loss(golden):
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    with torch.no_grad():
        loss, out = O.train_step_st(P, cfg, I['src'], I['tgt'], I['acous_feats'], I['acous_lens'])
    assert rel_err(out['logps_st'], golden['st/logps_st']) < TOL
    assert rel_err(out['emb_st'], golden['st/emb_st']) < TOL
    assert abs(float(loss) - float(golden['st/loss'])) < TOL * abs(float(golden['st/loss']))
    assert torch.equal(out['preds_st'], golden['st/preds_st'])


def test_st_gradients(golden):
    P, cfg, I = golden.params(requires_grad=True), golden.cfg, golden.inputs()
    loss, _ = O.train_step_st(P, cfg, I['src'], I['tgt'], I['acous_feats'], I['acous_lens'])
    loss.backward()
    ref = golden.group('st_grad')
    no_grad = set(str(s) for s in golden.z['st/no_grad_params'])
    for name, g in ref.items():
        assert P[name].grad is not None, name
        assert rel_err(P[name].grad, g) < 5e-5, name
    for name in no_grad:                      # enc_src.enc.*, dec_tgt.dec.*, las.decoder.acous_out.*
        assert P[name].grad is None or float(P[name].grad.abs().sum()) == 0.0, name


def test_las_free_running(golden):
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    with torch.no_grad():
        enc = O.las_encoder(P, cfg, I['acous_feats'], I['acous_lens'])
        embs, logps, syms, lengths = O.las_decoder(P, cfg, enc, I['acous_lens'])
    assert rel_err(enc, golden['las/enc_out']) < TOL
    assert torch.equal(syms, golden['las/symbols'])
    assert list(lengths) == [int(v) for v in golden['las/lengths']]
    assert rel_err(embs, golden['las/embs']) < TOL
    assert rel_err(logps, golden['las/logps']) < TOL


def test_las_encoder_loops_match_fused(golden):
    """First-principles packed-BLSTM restatement == torch's fused LSTM on the same weights."""
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    with torch.no_grad():
        a = O.las_encoder(P, cfg, I['acous_feats'], I['acous_lens'], loops=True)
        b = O.las_encoder(P, cfg, I['acous_feats'], I['acous_lens'], loops=False)
    assert rel_err(a, b) < 1e-5
    assert rel_err(a, golden['las/enc_out']) < TOL


def test_hoisted_keys_same_numbers(golden):
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    with torch.no_grad():
        enc = O.las_encoder(P, cfg, I['acous_feats'], I['acous_lens'])
        a = O.las_decoder(P, cfg, enc, I['acous_lens'], hoist_keys=False)
        b = O.las_decoder(P, cfg, enc, I['acous_lens'], hoist_keys=True)
    assert torch.equal(a[2], b[2]) and a[3] == b[3]
    assert rel_err(a[0], b[0]) < 1e-6


def test_greedy_eval_ids_exact(golden):
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    out = O.forward_eval_st(P, cfg, I['acous_feats'], I['acous_lens'])
    assert torch.equal(out['preds_st'], golden['eval/preds_st'])


@pytest.mark.parametrize('beam', [1, 3])
def test_translate_ids_exact(golden, beam):
    P, cfg, I = golden.params(), golden.cfg, golden.inputs()
    ids = O.forward_translate_st(P, cfg, I['acous_feats'], I['acous_lens'], beam_width=beam,
                                 penalty_factor=1, max_seq_len=cfg.max_seq_len_tgt)
    assert torch.equal(ids, golden[f'translate/beam{beam}'])


def test_mt_mode(golden):
    P, cfg, I = golden.params(requires_grad=True), golden.cfg, golden.inputs()
    out = O.forward_train_mt(P, cfg, I['src'], I['tgt'], I['emb_dyn_ave'])
    loss = O.masked_nll(out['logps_mt'], I['tgt'])
    loss.backward()
    assert rel_err(out['logps_mt'], golden['mt/logps_mt']) < TOL
    assert abs(float(loss) - float(golden['mt/loss'])) < TOL * abs(float(golden['mt/loss']))
    for name, n in golden.group('mt_gradnorm').items():
        assert abs(float(P[name].grad.norm()) - float(n)) < 1e-4 * float(n) + 1e-7, name


def test_asr_mode_teacher_forced(golden):
    P, cfg, I = golden.params(requires_grad=True), golden.cfg, golden.inputs()
    out = O.forward_train_asr(P, cfg, I['src'], golden['asr/aug_feats'], I['acous_lens'])
    lp = out['logps_asr']
    tgt = I['src']
    mask = tgt[:, 1:].ne(O.PAD).reshape(-1)
    per = torch.nn.functional.nll_loss(lp.reshape(-1, lp.size(-1)), tgt[:, 1:].reshape(-1), reduction='none')
    loss = per.masked_select(mask).sum() / mask.sum()
    loss.backward()
    assert rel_err(lp, golden['asr/logps_asr']) < TOL
    assert abs(float(loss) - float(golden['asr/loss'])) < TOL * abs(float(golden['asr/loss']))
    assert list(out['lengths_asr']) == [int(v) for v in golden['asr/lengths']]
    for name, n in golden.group('asr_gradnorm').items():
        assert abs(float(P[name].grad.norm()) - float(n)) < 1e-4 * float(n) + 1e-7, name


# ---- the fixture at the benchmark's kernel shapes (H = 256 per direction, 8 heads x d_k = 64, B = 16; seeded weights)
def test_h256_oracle_matches_reference_fixture():
    from conftest import SeededGolden
    g = SeededGolden('st_h256')
    P, cfg, I = g.params(requires_grad=True), g.cfg, g.inputs()
    loss, out = O.train_step_st(P, cfg, I['src'], I['tgt'], I['acous_feats'], I['acous_lens'])
    loss.backward()
    assert rel_err(out['logps_st'], g['st/logps_st']) < TOL
    assert rel_err(out['emb_st'], g['st/emb_st']) < TOL
    assert abs(float(loss) - float(g['st/loss'])) < TOL * abs(float(g['st/loss']))
    assert torch.equal(out['preds_st'], g['st/preds_st'])
    assert torch.equal(out['preds_asr'], g['las/symbols'])
    assert list(out['lengths_asr']) == [int(v) for v in g['las/lengths']]
    g.grad_check({k: v.grad for k, v in P.items()}, 1e-4)
    Pn = {k: v.detach() for k, v in P.items()}
    for beam in (1, 5):
        ids = O.forward_translate_st(Pn, cfg, I['acous_feats'], I['acous_lens'], beam_width=beam, penalty_factor=1,
                                     max_seq_len=cfg.max_seq_len_tgt)
        assert torch.equal(ids, g[f'translate/beam{beam}']), beam
