"""Runs the tcgen05 attention core alone at the benchmark's decoder self-attention shape (for ncu / timing):
python scripts/run_mha_once.py [B] [Lq] [Lk]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 50
Lk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
k = CudaKernels()
torch.manual_seed(0)
kv = torch.randn(B, Lk, 1024, device='cuda').bfloat16()
q = torch.randn(B, Lq, 512, device='cuda').bfloat16()
mask = torch.ones(B, Lq, Lk, dtype=torch.uint8, device='cuda').tril()
do = torch.randn(B, Lq, 512, device='cuda').bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for name, fn in (('mha_fwd', lambda: k.mha_fwd(q, kv[:, :, :512], kv[:, :, 512:], mask, 8, 8.0)),):
    o, p = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name}: cold (L2 flushed) median {sorted(ts)[5]:.1f} us; back-to-back {e0.elapsed_time(e1) * 20:.1f} us')
o, p = k.mha_fwd(q, kv[:, :, :512], kv[:, :, 512:], mask, 8, 8.0)
dq, dk, dv = k.mha_bwd(do, q, kv[:, :, :512], kv[:, :, 512:], p, 8, 8.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    k.mha_bwd(do, q, kv[:, :, :512], kv[:, :, 512:], p, 8, 8.0)
e1.record(); torch.cuda.synchronize()
print(f'mha_bwd: back-to-back {e0.elapsed_time(e1) * 20:.1f} us')
