"""Per-shape GEMM time table of one cfg-3 training step (GPU): python scripts/gemm_table.py [bf16|fp32]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from b200st import runtime
from b200st.kernels import K
from oracle import st_oracle as O
from b200st.train_step import Trainer_ST
dt = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
runtime.set_compute_dtype(dt)
cfg = bench.st_config()
dev = torch.device('cuda')
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feats'].to(dev)],
         'acouslen': host['acous_lens']}
tr = Trainer_ST(use_gpu=True, batch_size=64)
for _ in range(3):
    tr._train_batch(model, items); model.zero_grad(set_to_none=True)
k = K()
orig = k.gemm
rec = []
def wrapped(a, b, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(a, b, **kw); e1.record()
    ta, tb = kw.get('trans_a', False), kw.get('trans_b', False)
    a2 = a[0] if a.dim() == 3 else a; b2 = b[0] if b.dim() == 3 else b
    M, Kk = (a2.size(1), a2.size(0)) if ta else (a2.size(0), a2.size(1))
    N = b2.size(0) if tb else b2.size(1)
    rec.append(((M, N, Kk, int(ta), int(tb), a.size(0) if a.dim() == 3 else 1, str(r.dtype)[6:]), e0, e1))
    return r
k.gemm = wrapped
tr._train_batch(model, items)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for key, e0, e1 in rec:
    agg[key][0] += 1; agg[key][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f'total gemm ms {tot:.2f} over {len(rec)} calls')
for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    M, N, Kk, ta, tb, nb, od = key
    fl = 2.0 * M * N * Kk * nb * n
    print(f'M={M:6d} N={N:6d} K={Kk:6d} ta={ta} tb={tb} batch={nb:3d} out={od:9s} calls={n:4d} ms={ms:8.3f} avg_us={ms/n*1e3:8.1f} TFLOP/s={fl/ms/1e9:8.1f}')
