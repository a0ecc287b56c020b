"""Flake hunt: the dropout mask-injection step (tests/test_dropout.py::_run_case) repeated N times in one process on
cuda; prints every parameter whose gradient error vs the fp32 oracle exceeds 5e-5 of its norm.
usage: python scripts/diag/dropout_flake_loop.py [N]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_dropout as T
from b200st import runtime as rt
from helpers import build_model, train_step
from oracle import st_oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rt.set_compute_dtype('fp32')
cfg = O.STConfig(**T.CFG)
P = O.init_params(cfg, seed=21, scale=2.0)
data = O.synthetic_batch(cfg, batch=3, frames=40, seed=22, ragged=True)
ref = None
first = None
for it in range(N):
    model = build_model(cfg, P, device='cuda'); T._set_dropout(model); model.train()
    rt.manual_seed(77); rt.site_log = {}
    loss, out = train_step(model, data, 'cuda'); loss.backward(); log = dict(rt.site_log); rt.site_log = None
    torch.cuda.synchronize()
    grads = {k: v.grad.double().cpu() for k, v in model.named_parameters() if v.grad is not None}
    if ref is None:
        rng = rt.current_rng(torch.device('cuda', 0))
        Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        O.DROP = lambda x, tag: x * T._mask_from_product(tag, tuple(x.shape), log, rng, 'cuda').to(x.dtype)
        l, _ = O.train_step_st(Pg, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        l.backward(); O.DROP = None
        ref = {k: v.grad.double() for k, v in Pg.items() if v.grad is not None}
        first = grads
    bad = []
    for k, g in ref.items():
        n = float(g.norm())
        if n == 0 or k not in grads: continue
        e = float((grads[k] - g).norm()) / n
        d = float((grads[k] - first[k]).norm()) / n
        if e > 5e-5 or d > 1e-6: bad.append((k, e, d))
    print(f'iter {it}: loss {loss.get_loss():.7f}; {len(bad)} suspicious', flush=True)
    for k, e, d in bad[:6]:
        diff = (grads[k] - first[k]).abs()
        print(f'    {k}: err vs oracle {e:.2e}, diff vs iteration 0 {d:.2e}, {int((diff > 1e-6 * float(ref[k].abs().max())).sum())} elements differ, max at {int(diff.argmax())}')
