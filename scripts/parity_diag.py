"""Per-parameter gradient error table at configs[2] full size: product fp32 / bf16 (LAS symbols pinned) vs the oracle on
cuda in fp64 ("truth") and fp32, plus the oracle under torch.autocast(bf16) as the stock-PyTorch bf16 noise floor.
Diagnostic only.   python scripts/parity_diag.py [--batch 64] [--frames 1000]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import bench
from b200st import runtime
from helpers import build_model, train_step
from oracle import st_oracle as O
from oracle_cuda import oracle_on_cuda, params_to
from test_gpu_oracle_fullsize import _force_las_symbols

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--frames', type=int, default=1000)
ap.add_argument('--top', type=int, default=20)
args = ap.parse_args()
cfg = bench.st_config()
P = O.init_params(cfg, seed=333)
data = O.synthetic_batch(cfg, args.batch, args.frames, seed=334)


def oracle(dtype, autocast=False, force=None):
    with oracle_on_cuda() as dev:
        Pg = {k: v.detach().to(dev, dtype).clone().requires_grad_(True) for k, v in P.items()}
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            if force is None:
                loss, out = O.train_step_st(Pg, cfg, data['src'].to(dev), data['tgt'].to(dev),
                                            data['acous_feats'].to(dev, dtype), data['acous_lens'])
            else:       # LAS teacher-forced on `force` [B, S]; the rest of forward_train_st unchanged
                tgt_d, src_d = data['tgt'].to(dev), data['src'].to(dev)
                tgt_mask, emb_tgt = O.target_embeddings(Pg, cfg, tgt_d)
                enc_ac = O.las_encoder(Pg, cfg, data['acous_feats'].to(dev, dtype), data['acous_lens'])
                ids = torch.cat([torch.full((force.size(0), 1), 2, dtype=torch.int64), force.to(dev)], 1)
                emb_dyn, lp_asr, preds_asr, lengths = O.las_decoder(Pg, cfg, enc_ac.float() if autocast else enc_ac, data['acous_lens'], tgt=ids, teacher_forcing=True)
                emb_src = O.mix_embeddings(Pg, src_d[:, 1:], emb_dyn)
                smask = O.length_mask(lengths, emb_src.size(1))
                enc_out, _ = O.tf_encoder(Pg, cfg, emb_src, smask)
                _, _, logps, _, _ = O.translation_decoder(Pg, cfg, emb_tgt, enc_out, tgt_mask, smask)
                loss = O.masked_nll(logps.float(), tgt_d)
                out = {'preds_asr': preds_asr}
        loss.backward()
        g = {k: v.grad.detach().double().cpu() for k, v in Pg.items() if v.grad is not None and float(v.grad.abs().sum()) > 0}
        res = float(loss), g, out['preds_asr'].squeeze(-1).cpu()
        del Pg, out, loss
    torch.cuda.empty_cache()
    return res


def table(name, g, ref, loss, loss_ref):
    gn = sum(float(v.norm() ** 2) for v in ref.values()) ** 0.5
    rows = []
    miss = [k for k in ref if k not in g]
    if miss:
        print(f'   ({len(miss)} parameters without a finite gradient, e.g. {miss[:2]})')
    ref = {k: r for k, r in ref.items() if k in g}
    for k, r in ref.items():
        e = float((g[k] - r).norm())
        rows.append((e / max(float(r.norm()), 1e-3 * gn), e / max(float(r.norm()), 1e-30), float(r.norm()), k))
    rows.sort(reverse=True)
    tot = sum(float((g[k] - r).norm() ** 2) for k, r in ref.items()) ** 0.5 / gn
    print(f'--- {name}: loss {loss:.7f} (ref {loss_ref:.7f}, rel {abs(loss - loss_ref) / abs(loss_ref):.2e}); global grad L2 rel {tot:.2e}')
    import collections
    hist = collections.Counter(min(int(a / 1e-2), 9) for a, _, _, _ in rows) if rows[0][0] > 5e-3 else None
    if hist:
        print('   histogram of per-parameter error in units of 1e-2:', dict(sorted(hist.items())))
    for a, b, n, k in rows[:args.top]:
        print(f'   {a:9.2e} (raw {b:9.2e}) |g|={n:9.2e}  {k}')


l64, g64, sym64 = oracle(torch.float64)
l32, g32, sym32 = oracle(torch.float32)
print('LAS symbols fp32 oracle == fp64 oracle:', bool((sym32 == sym64).all()), int((sym32 != sym64).sum()))
table('oracle fp32 vs oracle fp64', g32, g64, l32, l64)
try:
    lac, gac, symac = oracle(torch.float32, autocast=True, force=sym64)
    args.top = 400
    table('oracle autocast-bf16 (pinned symbols) vs oracle fp64', gac, g64, lac, l64)
    args.top = 20
except Exception as ex:
    print('autocast oracle failed:', type(ex).__name__, ex)


def product(dtype, force):
    runtime.set_compute_dtype(dtype)
    m = build_model(cfg, P, device='cuda'); m.train()
    if force:
        _force_las_symbols(m, sym64)
    loss, out = train_step(m, data, 'cuda')
    loss.backward()
    torch.cuda.synchronize()
    g = {k: v.grad.detach().double().cpu() for k, v in m.named_parameters() if v.grad is not None}
    return loss.get_loss(), g


l, g = product('fp32', False)
table('product fp32 vs oracle fp64', g, g64, l, l64)
table('product fp32 vs oracle fp32', g, g32, l, l32)
args.top = 400
l, g = product('bf16', True)
table('product bf16 (pinned symbols) vs oracle fp64', g, g64, l, l64)
