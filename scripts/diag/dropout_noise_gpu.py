import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/speech-translation-joint-embedding-passing_b200'); sys.path.insert(0, '/root/repo/tests')
import torch
import test_dropout as T
from b200st import kernels, runtime as rt
from fake_kernels import FakeKernels
from helpers import build_model, train_step
from oracle import st_oracle as O
old = None
rt.set_compute_dtype('fp32')
cfg = O.STConfig(**T.CFG)
P = O.init_params(cfg, seed=21, scale=2.0)
data = O.synthetic_batch(cfg, batch=3, frames=40, seed=22, ragged=True)
model = build_model(cfg, P, device='cuda'); T._set_dropout(model); model.train()
rt.manual_seed(77); rt.site_log = {}
loss, out = train_step(model, data, 'cuda'); loss.backward(); log = dict(rt.site_log); rt.site_log = None
rng = rt.current_rng(torch.device('cuda', 0))
def run(dt):
    Pg = {k: v.clone().to(dt).requires_grad_(True) for k, v in P.items()}
    O.DROP = lambda x, tag: x * T._mask_from_product(tag, tuple(x.shape), log, rng, 'cuda').to(x.dtype)
    try:
        l, _ = O.train_step_st(Pg, cfg, data['src'], data['tgt'], data['acous_feats'].to(dt), data['acous_lens'])
        l.backward()
    finally:
        O.DROP = None
    return float(l), {k: v.grad.double() for k, v in Pg.items() if v.grad is not None}
l32, g32 = run(torch.float32)
l64, g64 = run(torch.float64)
named = dict(model.named_parameters())
print('loss', l32, l64, loss.get_loss())
worst = []
for k in g64:
    n = float(g64[k].norm())
    if n == 0: continue
    e_or = float((g32[k] - g64[k]).norm()) / n
    e_pr = float((named[k].grad.double().cpu() - g64[k]).norm()) / n
    e_pp = float((named[k].grad.double().cpu() - g32[k]).norm()) / n
    worst.append((e_or, e_pr, e_pp, k))
worst.sort(reverse=True)
for w in sorted(worst, key=lambda w: -w[1])[:12]: print('oracle32-vs-64 %.2e  product(cuda fp32)-vs-64 %.2e  product-vs-oracle32 %.2e  %s' % w)
