"""Drives the reference's UNMODIFIED `train`, `translate` and `trainer.trainer_st` modules either on the reference's
own modules (--impl reference) or on this repo's modules through b200st.dropin (--impl b200, kernels replaced by the
torch stand-in of tests/fake_kernels.py: this runs in the build container, which has no GPU).  Writes result.json.

Build-container only (needs /root/reference); tests/test_dropin_reference.py runs it twice and compares.
TEST INFRASTRUCTURE ONLY."""
import argparse
import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200')


def shims(dyn_ave):
    """packages the reference imports at top level that this image lacks; none touches arithmetic."""
    for name in ['bpemb', 'matplotlib', 'matplotlib.pyplot', 'torchtext']:
        sys.modules[name] = types.ModuleType(name)
    sys.modules['bpemb'].BPEmb = object
    sys.modules['matplotlib'].use = lambda *a, **k: None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    np_load = np.load
    np.load = lambda p, *a, **k: (dyn_ave if str(p).endswith('dyn_emb_ave.npy') else np_load(p, *a, **k))
    mf = torch.Tensor.masked_fill                        # torch >= 2 rejects the reference's uint8 masks
    torch.Tensor.masked_fill = lambda self, m, v: mf(self, m.bool() if m.dtype == torch.uint8 else m, v)


class _Iter:
    """What `iter(DataLoader)` looked like to the reference: len() and .next() (translate.py:101-110)."""

    def __init__(self, batches):
        self.batches, self.i = batches, 0

    def __len__(self):
        return len(self.batches)

    def next(self):
        self.i += 1
        return self.batches[self.i - 1]

    __next__ = next


class _Loader:
    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return _Iter(self.batches)


class FakeTestSet:
    """The slice of utils.dataset.Dataset that translate.translate touches."""

    def __init__(self, batches, vocab):
        self.iter_loader = _Loader(batches)
        self.tgt_id2word = {i: ('<pad>' if i == 0 else '</s>' if i == 3 else '<spc>' if i == 4 else f'w{i}')
                            for i in range(vocab)}
        self.src_id2word = dict(self.tgt_id2word)

    def construct_batches(self, is_train=False):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', required=True, choices=['reference', 'b200'])
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--golden', default='st_small')
    ap.add_argument('--out', required=True)
    ap.add_argument('--device', default='cpu', choices=['cpu', 'cuda'],
                    help="cuda: --impl b200 runs the REAL kernels (fp32 mode), --impl reference stock PyTorch CUDA")
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    z = np.load(os.path.join(HERE, 'golden', args.golden + '.npz'))
    shims(np.array(z['in/emb_dyn_ave']))
    if args.impl == 'b200':
        sys.path[:0] = [PKG, HERE]
        from b200st import dropin, kernels
        dropin.install(args.reference)
        if args.device == 'cpu':
            from fake_kernels import FakeKernels
            kernels.set_backend(FakeKernels())
        else:
            from b200st import runtime
            runtime.set_compute_dtype('fp32')
    else:
        sys.path.insert(0, args.reference)

    # ---- the reference's entry points, unmodified (train.py:9-16, translate.py:11-19, trainer_st.py:12-18)
    import train                                                    # noqa: F401
    import translate
    from trainer.trainer_st import Trainer_ST
    from modules.optim import Optimizer
    from models.Seq2seq import Seq2seq
    origin = {m: os.path.realpath(sys.modules[m].__file__) for m in
              ('train', 'translate', 'trainer.trainer_st', 'trainer.trainer_base', 'utils.misc', 'utils.dataset',
               'modules.checkpoint', 'modules.loss', 'modules.optim', 'modules.layers', 'models.Seq2seq', 'models.Dec')}

    cfg = {k[4:]: int(z[k]) for k in z.files if k.startswith('cfg/')}
    torch.manual_seed(1); np.random.seed(1); random.seed(1)
    model = Seq2seq(cfg['V'], cfg['V'], share_embedder=False, enc_embedding_size=cfg['E'], dec_embedding_size=cfg['E'],
                    max_seq_len_src=cfg['S'], max_seq_len_tgt=cfg['L'], num_heads=cfg['heads'], dim_model=cfg['dim_model'],
                    dim_feedforward=cfg['FF'], enc_layers=cfg['layers'], dec_layers=cfg['layers'], embedding_dropout=0.0,
                    dropout=0.0, acous_dim=cfg['F'], acous_hidden_size=cfg['H'], mode='ST', load_mode='null')
    sd = {k[6:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith('param/')}
    model.load_state_dict(sd, strict=False)
    gpu = args.device == 'cuda'
    dev = torch.device(args.device)
    if gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        model = model.to(dev)
    for mod in model.modules():                                     # the hidden attention dropout (layers.py:207)
        if type(mod).__name__ == 'ScaledDotProductAttention':
            mod.dropout.p = 0.0
    src, tgt = torch.from_numpy(np.array(z['in/src'])), torch.from_numpy(np.array(z['in/tgt']))
    feats = torch.from_numpy(np.array(z['in/acous_feats']))
    lens = [int(v) for v in z['in/acous_lens']]
    B = src.size(0)
    batch_items = {'srcid': [src], 'srclen': [src.size(1)] * B, 'tgtid': [tgt], 'tgtlen': [tgt.size(1)] * B,
                   'acous_feat': [feats], 'acouslen': [torch.tensor([n]) for n in lens]}

    # ---- Trainer_ST(expt_dir=...)._train_batch, three optimizer steps (trainer_st.py:211-299, trainer_base.py:420-426)
    t = Trainer_ST(expt_dir=os.path.join(args.out, 'expt'), load_dir=None, load_mode='null', batch_size=B, use_gpu=gpu,
                   learning_rate=1e-3, learning_rate_init=1e-3, lr_warmup_steps=0, max_grad_norm=1.0,
                   loss_coeff={'nll_asr': 1.0, 'nll_mt': 1.0, 'nll_st': 1.0}, minibatch_partition=1)
    t.optimizer = Optimizer(torch.optim.Adam(model.parameters(), lr=t.learning_rate_init), max_grad_norm=t.max_grad_norm)
    model.train()
    losses = [float(t._train_batch(model, batch_items, None, i, 3)['nll_loss_de']) for i in range(3)]
    wsum = {n: float(p.detach().double().abs().sum()) for n, p in model.named_parameters()}

    # ---- translate.translate, beam 1 and 3, history HYP and REF (translate.py:56-197)
    out_txt = {}
    model.load_state_dict(sd, strict=False)                         # back to the golden weights
    for history in ('HYP', 'REF'):
        for k in (1, 3):
            d = os.path.join(args.out, f'tr_{history}_{k}')
            os.makedirs(d, exist_ok=True)
            ts = FakeTestSet([batch_items], cfg['V'])
            translate.translate(ts, model, d, gpu, cfg['L'], k, dev, gen_mode='ST', history=history)
            out_txt[f'{history}_{k}'] = open(os.path.join(d, 'translate.txt'), encoding='utf8').read().split('\n')
    json.dump({'impl': args.impl, 'origin': origin, 'losses': losses, 'wsum': wsum, 'translate': out_txt,
               'golden_loss': float(z['st/loss'])}, open(os.path.join(args.out, 'result.json'), 'w'))


if __name__ == '__main__':
    main()
