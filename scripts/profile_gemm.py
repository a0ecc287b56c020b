"""One big GEMM (the l2 BLSTM input projection, M=32256 N=1024 K=1024) for an ncu --set full capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
k = CudaKernels()
a = torch.randn(32256, 1024, device='cuda').bfloat16(); w = torch.randn(1024, 1024, device='cuda').bfloat16()
bias = torch.randn(1024, device='cuda')
for _ in range(5):
    y = k.gemm(a, w, trans_b=True, bias=bias)
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()))
