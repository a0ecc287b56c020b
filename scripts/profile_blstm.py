"""Runs the tensor-core BLSTM recurrence kernels alone (for ncu): python scripts/profile_blstm.py [T] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
T = int(sys.argv[1]) if len(sys.argv) > 1 else 252
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H = 256
k = CudaKernels()
torch.manual_seed(0)
xproj = torch.randn(2, T, B, 4 * H, device='cuda').bfloat16()
wf = torch.randn(4 * H, H, device='cuda') / 16
wr = torch.randn(4 * H, H, device='cuda') / 16
lens = torch.full((B,), T, dtype=torch.int32, device='cuda')
out = torch.empty(T // 2, B, 4 * H, device='cuda', dtype=torch.bfloat16)
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    hs, acts, cs = k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2)
    e1.record()
    dg = k.blstm_bwd(torch.randn_like(out), B * 4 * H, 4 * H, 2, acts, cs, wf, wr, lens, torch.bfloat16)
    e2.record()
    torch.cuda.synchronize()
    print(f'T={T} B={B}: fwd {e0.elapsed_time(e1):.3f} ms ({e0.elapsed_time(e1) / T * 1e3:.2f} us/step)  '
          f'bwd {e1.elapsed_time(e2):.3f} ms ({e1.elapsed_time(e2) / T * 1e3:.2f} us/step)')
