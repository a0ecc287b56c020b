"""BASELINE.json's FULL sizes (configs[2]: B 64, 1000 frames, V 10k, 6+6 layers): size-independent properties of the CUDA
path -- the two numerical modes agree with each other within the bf16 contract, the whole-step CUDA graph reproduces the
eager step, and the KV-cached / graphed inference loop returns the token ids of the reference-shaped recompute loop.
Parity with the ORACLE at these sizes is tests/test_gpu_oracle_fullsize.py."""
import pytest
import torch

import bench
from b200st import runtime
from oracle import st_oracle as O

pytestmark = pytest.mark.gpu


def _items(batch, frames, seed=333):
    cfg = bench.st_config()
    host = O.synthetic_batch(cfg, batch, frames, seed=seed)
    dev = torch.device('cuda')
    return cfg, host, {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)],
                       'acous_feat': [host['acous_feats'].to(dev)], 'acouslen': host['acous_lens']}


def _step(dtype, items, cfg):
    from b200st.train_step import Trainer_ST
    runtime.set_compute_dtype(dtype)
    model = bench.build_model(cfg, torch.device('cuda'))
    loss = float(Trainer_ST(use_gpu=True, batch_size=64)._train_batch_device(model, items))
    g2 = sum(float(p.grad.double().pow(2).sum()) for p in model.parameters() if p.grad is not None)
    gdec = model.out_tgt.weight.grad.detach().clone()
    return loss, g2 ** 0.5, gdec, model


def test_full_size_bf16_step_agrees_with_fp32_step():
    try:
        cfg, _, items = _items(64, 1000)
        l32, n32, g32, _ = _step('fp32', items, cfg)
        l16, n16, g16, _ = _step('bf16', items, cfg)
        assert abs(l16 - l32) < 2e-2 * abs(l32), (l16, l32)                      # north-star bf16 tolerance
        assert abs(n16 - n32) < 2e-2 * n32, (n16, n32)
        assert float((g16 - g32).norm() / g32.norm()) < 2e-2
        assert abs(l32 - 9.21) < 0.5                                             # ~ log(10000): random-init weights
    finally:
        runtime.set_compute_dtype('fp32')


def test_full_size_graph_replay_equals_eager_step_bf16():
    from b200st.graph import GraphedTrainStep
    from b200st.train_step import Trainer_ST
    try:
        cfg, _, items = _items(64, 1000)
        le, ne, ge, model = _step('bf16', items, cfg)
        model.zero_grad(set_to_none=True)
        g = GraphedTrainStep(model, Trainer_ST(use_gpu=True, batch_size=64), items)
        lg = float(g())
        ng = sum(float(p.grad.double().pow(2).sum()) for p in model.parameters() if p.grad is not None) ** 0.5
        assert abs(lg - le) < 1e-3 * abs(le) and abs(ng - ne) < 1e-2 * ne, (lg, le, ng, ne)   # split-K atomics: not bit-exact
    finally:
        runtime.set_compute_dtype('fp32')


def test_full_size_cached_graphed_translate_equals_recompute_fp32():
    runtime.set_compute_dtype('fp32')
    cfg, host, _ = _items(16, 1000, seed=7)
    model = bench.build_model(cfg, torch.device('cuda')).eval()
    feats = host['acous_feats'].cuda()
    for k in (1, 3):
        out = {}
        for cached in (True, False):
            model.decode_cache = cached
            for _ in range(2 if cached else 1):      # second call replays the captured graphs
                out[cached] = model.forward_translate(acous_feats=feats.clone(), acous_lens=host['acous_lens'], beam_width=k,
                                                      penalty_factor=1, use_gpu=True, max_seq_len=24, mode='ST')
        assert torch.equal(out[True], out[False]), k
