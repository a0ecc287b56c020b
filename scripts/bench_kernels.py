"""Back-to-back (launch-overhead-free) timings of the small kernels of the cfg-3 step at their training shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
k = CudaKernels()
bf = torch.bfloat16
dev = 'cuda'


def timeit(name, fn, n=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:44s} {e0.elapsed_time(e1) / n * 1e3:8.2f} us', flush=True)


R, D = 3200, 512
x, dy = torch.randn(R, D, device=dev).to(bf), torch.randn(R, D, device=dev).to(bf)
g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
y, mean, rstd = k.layernorm_fwd(x, g, b, 1e-6)
dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
timeit('layernorm_fwd 3200x512', lambda: k.layernorm_fwd(x, g, b, 1e-6))
timeit('layernorm_bwd 3200x512', lambda: k.layernorm_bwd(dy, x, g, mean, rstd, dg, db))
timeit('layernorm_bwd+add 3200x512', lambda: k.layernorm_bwd(dy, x, g, mean, rstd, dg, db, add=dy))
B, H, L = 64, 8, 50
q, kk, v = (torch.randn(B, L, 512, device=dev).to(bf) for _ in range(3))
mask = torch.ones(B, L, L, dtype=torch.uint8, device=dev).tril()
o, p = k.mha_fwd(q, kk, v, mask, H, 8.0)
timeit('mha_fwd B64 H8 L50', lambda: k.mha_fwd(q, kk, v, mask, H, 8.0))
timeit('mha_bwd B64 H8 L50', lambda: k.mha_bwd(o, q, kk, v, p, H, 8.0))
Tk = 126
qq, wk, vals = torch.randn(B, 512, device=dev).to(bf), torch.randn(B, Tk, 512, device=dev).to(bf), torch.randn(B, Tk, 512, device=dev).to(bf)
kl = torch.full((B,), Tk, dtype=torch.int32, device=dev)
cx, pp = k.las_attn_fwd(qq, wk, vals, kl)
timeit('las_attn_fwd B64 Tk126', lambda: k.las_attn_fwd(qq, wk, vals, kl))
timeit('las_attn_bwd B64 Tk126', lambda: k.las_attn_bwd(cx, wk, vals, pp))
gates, cp = torch.randn(B, 2048, device=dev).to(bf), torch.randn(B, 512, device=dev)
timeit('lstm_cell_fwd B64 H512', lambda: k.lstm_cell_fwd(gates, cp))
big = torch.randn(3200, 2048, device=dev).to(bf)
timeit('colsum 3200x2048', lambda: k.colsum(big))
timeit('relu_bwd 3200x2048', lambda: k.relu_bwd(big, big))
logits = torch.randn(B, 10000, device=dev).to(bf)
idx = torch.empty(B, dtype=torch.int64, device=dev)
timeit('argmax_rows 64x10000', lambda: k.argmax_rows(logits, idx))
