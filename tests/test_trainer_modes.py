"""f-4 slices: the joint ASR + ST trainer step (reference trainer/trainer_asr_st.py:253-357) and whole-module checkpoint
pickles (reference modules/checkpoint.py:76,160-164 saves / loads `model` with torch.save / torch.load)."""
import io
import random

import pytest
import torch

from b200st import kernels
from fake_kernels import FakeKernels
from helpers import build_model
from oracle import st_oracle as O

CFG = dict(enc_vocab_size=90, dec_vocab_size=90, enc_embedding_size=16, dec_embedding_size=16, max_seq_len_src=7,
           max_seq_len_tgt=9, num_heads=2, dim_model=32, dim_feedforward=48, enc_layers=2, dec_layers=2, acous_dim=12,
           acous_hidden_size=16)


@pytest.fixture
def fake_backend():
    old = kernels.set_backend(FakeKernels())
    yield
    kernels.set_backend(old)


def _case(device='cpu'):
    cfg = O.STConfig(**CFG)
    P = O.init_params(cfg, seed=3, scale=2.0)
    data = O.synthetic_batch(cfg, batch=4, frames=40, seed=5, ragged=True)
    m = build_model(cfg, P, device=device, mode='ASR_ST')
    items = {'srcid': [data['src'].to(device)], 'tgtid': [data['tgt'].to(device)], 'acous_feat': [data['acous_feats'].to(device)],
             'acouslen': [int(n) for n in data['acous_lens']], 'srclen': [cfg.max_seq_len_src] * 4,
             'tgtlen': [cfg.max_seq_len_tgt] * 4}
    return cfg, P, data, m, items


def _manual(m, data, coeff, device):
    """The two-loss sum assembled by hand from forward_train('ASR_ST') outputs (trainer_asr_st.py:306-346)."""
    src, tgt = data['src'].to(device), data['tgt'].to(device)
    lens = [torch.tensor([n]) for n in data['acous_lens']]
    out = m.forward_train(src, tgt=tgt, acous_feats=data['acous_feats'].to(device), acous_lens=lens, mode='ASR_ST',
                          use_gpu=device != 'cpu')
    def nll(logps, target):
        mask = target.ne(0)
        picked = -logps.float().gather(2, target.unsqueeze(2)).squeeze(2)
        return (picked * mask).sum() / mask.sum()
    de = coeff['nll_st'] * nll(out['logps_st'][:, :-1], tgt[:, 1:])
    en = coeff['nll_asr'] * nll(out['logps_asr'], src[:, 1:])
    (de + en).backward()
    return float(de), float(en)


def _check_trainer(device):
    from b200st.train_step import Trainer_ASR_ST
    coeff = {'nll_asr': 0.3, 'nll_st': 1.0}
    cfg, P, data, m, items = _case(device)
    m.train()
    random.seed(11)                      # SpecAug (Enc.py:87-117) and the teacher-forcing draw use Python's generator
    res = Trainer_ASR_ST(use_gpu=device != 'cpu', batch_size=4, loss_coeff=coeff)._train_batch(m, items)
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    m2 = build_model(cfg, P, device=device, mode='ASR_ST')
    m2.train()
    random.seed(11)
    de, en = _manual(m2, data, coeff, device)
    assert abs(res['nll_loss_de'] - de) < 1e-5 * abs(de) and abs(res['nll_loss_en'] - en) < 1e-5 * abs(en)
    named = dict(m2.named_parameters())
    assert 'las.decoder.acous_out.weight' in grads            # the ASR loss reaches the LAS output layer
    for n, g in grads.items():
        ref = named[n].grad
        assert ref is not None and float((g - ref).norm()) <= 1e-4 * max(float(ref.norm()), 1e-6), n
    # gradient accumulation over two minibatches == one minibatch of everything when every utterance has the same number
    # of target / source tokens is NOT assumed here: just check the partition runs and accumulates
    m3 = build_model(cfg, P, device=device, mode='ASR_ST')
    m3.train()
    random.seed(11)
    r2 = Trainer_ASR_ST(use_gpu=device != 'cpu', batch_size=4, minibatch_partition=2, loss_coeff=coeff)._train_batch(m3, items)
    assert r2['nll_loss_de'] > 0 and r2['nll_loss_en'] > 0


def test_asr_st_trainer_step_cpu(fake_backend):
    _check_trainer('cpu')


def test_whole_module_pickle_roundtrip_cpu(fake_backend):
    """checkpoint.py:76 pickles the whole model: class paths resolve, parameters survive, inference caches are dropped."""
    cfg, P, data, m, _ = _case('cpu')
    m.eval()
    lens = [torch.tensor([n]) for n in data['acous_lens']]
    a = m.forward_translate(acous_feats=data['acous_feats'].clone(), acous_lens=lens, beam_width=2, penalty_factor=1,
                            use_gpu=False, max_seq_len=9, mode='ST')
    assert hasattr(m, '_beam')
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert type(m2).__module__ == 'models.Seq2seq' and not hasattr(m2, '_beam')
    for (n1, p1), (n2, p2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2)
    b = m2.forward_translate(acous_feats=data['acous_feats'].clone(), acous_lens=lens, beam_width=2, penalty_factor=1,
                             use_gpu=False, max_seq_len=9, mode='ST')
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_asr_st_trainer_step_gpu():
    from b200st import runtime
    runtime.set_compute_dtype('fp32')
    _check_trainer('cuda')


@pytest.mark.parametrize('boost', [0.0, 1.5, 4.0, 40.0])
def test_beam_search_early_exit_and_width_cpu(fake_backend, boost):
    _early_exit_case('cpu', boost)


@pytest.mark.gpu
@pytest.mark.parametrize('boost', [0.0, 4.0, 40.0])
def test_beam_search_early_exit_and_width_gpu_graphs(boost):
    from b200st import runtime
    runtime.set_compute_dtype('fp32')
    _early_exit_case('cuda', boost)


def _early_exit_case(device, boost):
    """The static-buffer, KV-cached search loop against the reference-shaped recompute loop when hypotheses hit EOS:
    `boost` pushes the EOS logit (through the decoder's final LayerNorm bias) so that the all-EOS early exit
    (Seq2seq.py:388-393) fires after 0 .. max steps; output shape (the `reshape(batch, -1)[:, :max_seq_len]` quirk) and
    every token id must agree for greedy and beam search."""
    cfg, P, data, m, _ = _case(device)
    m.eval()
    with torch.no_grad():
        w = m.out_tgt.weight[3]                                  # EOS row
        m.dec_tgt.norm.bias.add_(boost * w / w.pow(2).sum())
    lens = [torch.tensor([n]) for n in data['acous_lens']]
    widths = set()
    for k in (1, 3):
        outs = {}
        for cached in (True, False):
            m.decode_cache = cached
            for _ in range(2 if cached else 1):      # on the GPU the second call replays the captured graphs
                outs[cached] = m.forward_translate(acous_feats=data['acous_feats'].clone().to(device), acous_lens=lens,
                                                   beam_width=k, penalty_factor=1, use_gpu=device != 'cpu', max_seq_len=9,
                                                   mode='ST')
        assert outs[True].shape == outs[False].shape, (k, outs[True].shape, outs[False].shape)
        assert torch.equal(outs[True], outs[False]), (k, boost)
        widths.add(outs[True].size(1))
    if boost >= 40.0:
        assert min(widths) < 9          # greedy: every hypothesis ended at once, the loop stopped early


# ------------------------------------------------------------------------------------------------
# MT and ASR trainer steps (trainer_mt.py:199-282, trainer_asr.py:199-283) against the REAL reference's goldens
# ------------------------------------------------------------------------------------------------
def test_mt_and_asr_trainer_steps_match_reference_goldens(golden, fake_backend):
    from b200st.train_step import Trainer_ASR, Trainer_MT
    I = golden.inputs()
    m = build_model(golden.cfg, golden.params())
    m.EMB_DYN_AVE = golden['in/emb_dyn_ave']
    m.train()
    res = Trainer_MT(use_gpu=False, batch_size=I['src'].size(0))._train_batch(
        m, {'srcid': [I['src']], 'tgtid': [I['tgt']], 'srclen': None, 'tgtlen': None})
    assert set(res) == {'nll_loss_de'}
    assert abs(res['nll_loss_de'] - float(golden['mt/loss'])) < 2e-5 * abs(float(golden['mt/loss']))
    named = dict(m.named_parameters())
    for name, n in golden.group('mt_gradnorm').items():
        assert abs(float(named[name].grad.norm()) - float(n)) < 2e-4 * float(n) + 1e-7, name
    m.zero_grad()
    m.las.encoder.spec_aug = False          # the golden's features are the already-augmented copy (make_golden.py)
    res = Trainer_ASR(use_gpu=False, batch_size=I['src'].size(0))._train_batch(
        m, {'srcid': [I['src']], 'acous_feat': [golden['asr/aug_feats']], 'acouslen': I['acous_lens']})
    assert res['nll_loss_de'] == 0 and abs(res['nll_loss_en'] - float(golden['asr/loss'])) < 2e-5 * abs(float(golden['asr/loss']))
    for name, n in golden.group('asr_gradnorm').items():
        assert abs(float(named[name].grad.norm()) - float(n)) < 2e-4 * float(n) + 1e-7, name


# ------------------------------------------------------------------------------------------------
# a checkpoint WRITTEN BY THE REFERENCE loads into this repo's classes (checkpoint.py:54-180), and Trainer.train's
# load_mode / load_freeze path (trainer_base.py:244-313, trainer_st.py:30) trains with las.* frozen
# ------------------------------------------------------------------------------------------------
def test_reference_written_checkpoint_loads_translates_and_fine_tunes_frozen(tmp_path):
    import json
    import os
    import subprocess
    import sys
    ref = '/root/reference'
    if not os.path.isdir(os.path.join(ref, 'models')):
        pytest.skip('build container only: needs the reference to write the checkpoint')
    drv = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ckpt_driver.py')
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    w = subprocess.run([sys.executable, drv, '--write', str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert w.returncode == 0, w.stderr[-3000:]
    assert '/root/reference/models/Seq2seq.py' in json.loads(w.stdout.strip().splitlines()[-1])['cls']
    r = subprocess.run([sys.executable, drv, '--load', str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert 'speech-translation-joint-embedding-passing_b200/models/Seq2seq.py' in res['cls_file']
    assert 'speech-translation-joint-embedding-passing_b200/modules/optim.py' in res['optimizer_cls_file']
    assert (res['epoch'], res['step']) == (3, 77)
    assert res['beam1_equal_golden'] and res['beam3_equal_golden']
    assert abs(res['loss'] - res['golden_loss']) < 2e-5 * abs(res['golden_loss'])
    assert res['frozen_with_grad'] == [] and res['worst_unfrozen_grad_err'] < 5e-5


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', ['fp32', 'bf16'])
def test_frozen_las_step_launches_no_lstm_backward_and_matches_oracle(dtype):
    """The fine-tuning recipe (trainer_st.py:30 load_freeze=True): las.* has requires_grad=False.  No recurrence / LAS-decoder
    backward kernel may be launched, and the loss and every remaining gradient equal the oracle's."""
    from b200st import runtime
    from b200st.kernels import K
    from b200st.train_step import Trainer_ST
    from test_gpu_parity import _grad_check
    runtime.set_compute_dtype(dtype)
    try:
        cfg = O.STConfig(enc_vocab_size=304, dec_vocab_size=304, enc_embedding_size=24, dec_embedding_size=24,
                         max_seq_len_src=8, max_seq_len_tgt=11, num_heads=2, dim_model=128, dim_feedforward=96,
                         enc_layers=2, dec_layers=2, acous_dim=16, acous_hidden_size=256)
        P = O.init_params(cfg, seed=3)
        data = O.synthetic_batch(cfg, batch=16, frames=61, seed=4, ragged=True)
        Pg = {k: v.clone().requires_grad_(not k.startswith('las.')) for k, v in P.items()}
        loss_ref, _ = O.train_step_st(Pg, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        loss_ref.backward()
        m = build_model(cfg, P, device='cuda')
        m.train()
        for n, p in m.named_parameters():
            if n.startswith('las.'):
                p.requires_grad = False
        calls = {'blstm_bwd': 0, 'lstm_cell_bwd': 0, 'las_attn_bwd': 0}
        k = K()
        orig = {n: getattr(k, n) for n in calls}
        for n in calls:
            def wrapped(*a, __n=n, **kw):
                calls[__n] += 1
                return orig[__n](*a, **kw)
            setattr(k, n, wrapped)
        try:
            items = {'srcid': [data['src'].cuda()], 'tgtid': [data['tgt'].cuda()], 'acous_feat': [data['acous_feats'].cuda()],
                     'acouslen': data['acous_lens']}
            loss = float(Trainer_ST(use_gpu=True, batch_size=16)._train_batch_device(m, items))
        finally:
            for n in calls:
                delattr(k, n)
        assert calls == {'blstm_bwd': 0, 'lstm_cell_bwd': 0, 'las_attn_bwd': 0}, calls
        tol = 1e-4 if dtype == 'fp32' else 2e-2
        assert abs(loss - float(loss_ref)) < tol * abs(float(loss_ref))
        named = dict(m.named_parameters())
        assert all(named[n].grad is None for n in named if n.startswith('las.'))
        ref = {n: v.grad for n, v in Pg.items() if v.grad is not None and float(v.grad.abs().sum()) > 0}
        assert ref and not any(n.startswith('las.') for n in ref)
        if dtype == 'fp32':
            _grad_check(named, ref, 1e-4)
        else:       # bf16: free-running LAS symbols may flip on near ties (nothing is pinned here): global bound only
            gn = sum(float(g.double().norm() ** 2) for g in ref.values()) ** 0.5
            dn = sum(float((named[n].grad.double().cpu() - g.double()).norm() ** 2) for n, g in ref.items()) ** 0.5
            assert dn / gn < 5e-2, dn / gn
    finally:
        runtime.set_compute_dtype('fp32')
