"""Where does the CTA-pair GEMM lose time?  Times 32256x1024x1024 with probe modes: normal, no TMA (MMA + epilogue only),
no epilogue stores, neither."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
k = CudaKernels()
for (M, N, K) in ((32256, 1024, 1024), (32256, 1024, 4096), (8192, 8192, 8192)):
    a = torch.randn(M, K, device='cuda').bfloat16(); w = torch.randn(N, K, device='cuda').bfloat16()
    out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    for mode, name in ((3, 'normal'), (3 | 4, 'no TMA'), (3 | 8, 'no stores'), (3 | 12, 'no TMA, no stores'), (1, 'persist 1-CTA'), (0, '1 tile/CTA')):
        k.set_gemm_persistent(mode)
        for _ in range(5):
            k.gemm(a, w, trans_b=True, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(20_000_000)
        e0.record()
        for _ in range(50):
            k.gemm(a, w, trans_b=True, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print(f'M={M} N={N} K={K} {name:22s} {us:8.2f} us {2.0*M*N*K/us/1e6:8.1f} TFLOP/s', flush=True)
    k.set_gemm_persistent(3)
    y = torch.matmul(a, w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        torch.matmul(a, w.t(), out=y)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f'M={M} N={N} K={K} {"cuBLAS (torch.matmul)":22s} {us:8.2f} us {2.0*M*N*K/us/1e6:8.1f} TFLOP/s', flush=True)
