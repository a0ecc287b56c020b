"""Checkpoint compatibility with the reference (checkpoint.py:54-180, trainer_base.py:244-313).
  --write DIR : the UNMODIFIED reference builds its model (golden st_small weights), takes one optimizer step and saves a
                checkpoint with its own `Checkpoint.save` (whole-module pickle + optimizer state).
  --load DIR  : with this repo's modules in front (b200st.dropin) the reference's own `Checkpoint.load` unpickles that file
                into THIS repo's classes; the loaded model translates / trains, and its parameters are copied into a fresh
                model the way `Trainer.train` does for `load_mode` with `load_freeze=True` (las.* frozen).
Build-container only.  TEST INFRASTRUCTURE ONLY."""
import argparse
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from dropin_driver import PKG, shims  # noqa: E402


def build(Seq2seq, z, cfg):
    m = Seq2seq(cfg['V'], cfg['V'], share_embedder=False, enc_embedding_size=cfg['E'], dec_embedding_size=cfg['E'],
                max_seq_len_src=cfg['S'], max_seq_len_tgt=cfg['L'], num_heads=cfg['heads'], dim_model=cfg['dim_model'],
                dim_feedforward=cfg['FF'], enc_layers=cfg['layers'], dec_layers=cfg['layers'], embedding_dropout=0.0,
                dropout=0.0, acous_dim=cfg['F'], acous_hidden_size=cfg['H'], mode='ST', load_mode='null')
    for mod in m.modules():
        if type(mod).__name__ == 'ScaledDotProductAttention':
            mod.dropout.p = 0.0
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--write')
    ap.add_argument('--load')
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--golden', default='st_small')
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    z = np.load(os.path.join(HERE, 'golden', args.golden + '.npz'))
    shims(np.array(z['in/emb_dyn_ave']))
    tl = torch.load                                        # torch >= 2.6 defaults to weights_only=True, which rejects the
    torch.load = lambda *a, **k: tl(*a, **dict(k, weights_only=False))     # reference's whole-module pickles (checkpoint.py:160-164)
    cfg = {k[4:]: int(z[k]) for k in z.files if k.startswith('cfg/')}
    sd = {k[6:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith('param/')}
    src, tgt = torch.from_numpy(np.array(z['in/src'])), torch.from_numpy(np.array(z['in/tgt']))
    feats = torch.from_numpy(np.array(z['in/acous_feats']))
    lens = [torch.tensor([int(v)]) for v in z['in/acous_lens']]
    if args.write:
        sys.path.insert(0, args.reference)
        from models.Seq2seq import Seq2seq
        from modules.checkpoint import Checkpoint
        from modules.optim import Optimizer
        m = build(Seq2seq, z, cfg)
        m.load_state_dict(sd, strict=False)
        opt = Optimizer(torch.optim.Adam(m.parameters(), lr=1e-3), max_grad_norm=1.0)
        path = Checkpoint(model=m, optimizer=opt, epoch=3, step=77, input_vocab={'a': 5}, output_vocab={'b': 6}).save(args.write)
        print(json.dumps({'path': path, 'cls': type(m).__module__ + ':' + os.path.realpath(sys.modules[type(m).__module__].__file__)}))
        return
    sys.path[:0] = [PKG, HERE]
    from b200st import dropin, kernels
    dropin.install(args.reference)
    from fake_kernels import FakeKernels
    kernels.set_backend(FakeKernels())
    from models.Seq2seq import Seq2seq
    from modules.checkpoint import Checkpoint               # the reference's own file
    from modules.loss import NLLLoss
    ck = Checkpoint.load(Checkpoint.get_latest_checkpoint(args.load))
    loaded = ck.model
    res = {'cls_file': os.path.realpath(sys.modules[type(loaded).__module__].__file__), 'epoch': ck.epoch, 'step': ck.step,
           'optimizer_cls_file': os.path.realpath(sys.modules[type(ck.optimizer).__module__].__file__)}
    # (1) the unpickled module, as translate.py uses it (translate.py:310-330)
    loaded.eval()
    with torch.no_grad():
        for k in (1, 3):
            ids = loaded.forward_translate(acous_feats=feats.clone(), acous_lens=lens, beam_width=k, penalty_factor=1,
                                           use_gpu=False, max_seq_len=cfg['L'], mode='ST')
            res[f'beam{k}_equal_golden'] = bool(torch.equal(ids, torch.from_numpy(np.array(z[f'translate/beam{k}']))))
    # (2) Trainer.train's load_mode path (trainer_base.py:244-313): copy parameters by name into a fresh model, freeze las.*
    fresh = build(Seq2seq, z, cfg)
    for name, param in fresh.named_parameters():
        for load_name, load_param in loaded.named_parameters():
            if name == load_name:
                assert param.data.size() == load_param.data.size()
                param.data = load_param.data
                if name.startswith('las.'):
                    param.requires_grad = False             # load_freeze=True (trainer_st.py:30)
    fresh.train()
    out = fresh.forward_train(src, tgt=tgt, acous_feats=feats.clone(), acous_lens=lens, mode='ST', use_gpu=False)
    lp = out['logps_st'][:, :-1, :]
    loss = NLLLoss(); loss.reset()
    keep = tgt.ne(0)[:, 1:]
    loss.eval_batch_with_mask(lp.reshape(-1, lp.size(-1)), tgt[:, 1:].reshape(-1), keep.reshape(-1))
    loss.norm_term = 1.0 * torch.sum(keep); loss.normalise(); loss.backward()
    res['loss'], res['golden_loss'] = loss.get_loss(), float(z['st/loss'])
    worst, frozen_with_grad = 0.0, []
    for n, p in fresh.named_parameters():
        key = 'st_grad/' + n
        if n.startswith('las.'):
            if p.grad is not None:
                frozen_with_grad.append(n)
        elif key in z.files:
            g = torch.from_numpy(np.array(z[key]))
            worst = max(worst, float((p.grad - g).norm() / (g.norm() + 1e-12)))
    res['worst_unfrozen_grad_err'], res['frozen_with_grad'] = worst, frozen_with_grad
    print(json.dumps(res))


if __name__ == '__main__':
    main()
