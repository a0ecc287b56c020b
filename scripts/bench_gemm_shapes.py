"""Back-to-back timing of individual GEMM shapes (GPU-bound: 200 launches queued, CUDA events around the batch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
k = CudaKernels()
shapes = [(64, 10000, 512, 0, 1), (64, 2048, 512, 0, 1), (64, 2048, 200, 0, 1), (64, 512, 512, 0, 1), (64, 512, 2048, 0, 0),
          (64, 200, 2048, 0, 0), (64, 512, 512, 0, 0), (3200, 512, 512, 0, 1), (3200, 512, 512, 0, 0), (1984, 512, 512, 0, 1),
          (3200, 1024, 512, 0, 1), (3200, 10000, 512, 0, 1), (32256, 1024, 1024, 0, 1), (32256, 1024, 1024, 0, 0),
          (64512, 1024, 80, 0, 1), (1024, 1024, 32256, 1, 0), (1024, 256, 64512, 1, 0), (512, 512, 3200, 1, 0), (2048, 512, 3200, 1, 0), (512, 2048, 2048, 1, 0),
          (512, 512, 2048, 1, 0), (2048, 1024, 64, 1, 0), (2048, 512, 64, 1, 0)]
import itertools
big = [(32256, 1024, 1024, 0, 1), (32256, 1024, 1024, 0, 0), (16128, 1024, 1024, 0, 1), (64512, 1024, 80, 0, 1), (3200, 10000, 512, 0, 1),
       (3200, 512, 10000, 0, 0), (8064, 1024, 1024, 0, 1), (4032, 1024, 1024, 0, 1)]
modes = [(s, 3) for s in shapes + big[2:]] + [(s, 1) for s in big] + [(s, 0) for s in big]
for (M, N, K, ta, tb), persist in modes:
    k.set_gemm_persistent(persist)
    a = torch.randn((K, M) if ta else (M, K), device='cuda').bfloat16()
    b = torch.randn((N, K) if tb else (K, N), device='cuda').bfloat16()
    od = torch.float32 if ta else torch.bfloat16
    out = torch.empty(M, N, device='cuda', dtype=od)
    for _ in range(5):
        k.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
    torch.cuda.synchronize()
    n = 100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)          # let the host run ahead so the launches are back to back on the GPU
    e0.record()
    for _ in range(n):
        k.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f'M={M:6d} N={N:6d} K={K:6d} ta={ta} tb={tb} persist={persist}  {us:8.2f} us  {2.0*M*N*K/us/1e6:8.1f} TFLOP/s', flush=True)
