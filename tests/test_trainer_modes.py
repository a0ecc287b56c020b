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
