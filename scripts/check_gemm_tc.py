"""GPU debugging aid: tcgen05 GEMM vs torch for every transpose form / tail / epilogue (run under gpurun)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'))
import torch
from b200st.kernels import CudaKernels
k = CudaKernels()
k.set_gemm_backend(2)
torch.manual_seed(0)
bad = 0
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
cases = [(128, 128, 64), (128, 128, 256), (256, 256, 512), (100, 72, 80), (3200, 512, 512), (64, 2048, 712),
         (64, 10000, 512), (1024, 80, 4096), (1024, 1024, 64512), (3136, 512, 10000), (37, 24, 8)]
for ta, tb in ((False, True), (False, False), (True, False), (True, True)):
    for (M, N, K) in cases:
        for od in (torch.bfloat16, torch.float32):
            a = torch.randn((K, M) if ta else (M, K), device='cuda').bfloat16()
            b = torch.randn((N, K) if tb else (K, N), device='cuda').bfloat16()
            if (a.stride(0) % 8) or (b.stride(0) % 8):
                continue
            ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float())
            try:
                y = k.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=od)
                torch.cuda.synchronize()
                e = rel(y, ref)
            except Exception as ex:
                e = float('nan'); print('EXC', ex)
            ok = e < (1e-2 if od == torch.bfloat16 else 2e-3)
            bad += (not ok)
            print(f'ta={int(ta)} tb={int(tb)} M={M} N={N} K={K} out={str(od)[6:]:9s} rel={e:.3e} {"ok" if ok else "FAIL"}', flush=True)
# epilogue: bias + relu + residual, strided views
a = torch.randn(300, 712, device='cuda').bfloat16(); w = torch.randn(512, 712, device='cuda').bfloat16()
bias = torch.randn(512, device='cuda'); res = torch.randn(300, 512, device='cuda').bfloat16()
y = k.gemm(a, w, trans_b=True, bias=bias, residual=res, alpha=0.5)
ref = 0.5 * (a.float() @ w.float().t()) + bias + res.float()
print('bias+res', rel(y, ref)); bad += rel(y, ref) > 1e-2
y = k.gemm(a, w, trans_b=True, bias=bias, relu=True)
ref = torch.relu(a.float() @ w.float().t() + bias)
print('bias+relu', rel(y, ref)); bad += rel(y, ref) > 1e-2
y = k.gemm(a[:, :200], w[:, :200], trans_b=True); ref = a[:, :200].float() @ w[:, :200].float().t()
print('col-slice A,B (ld>K)', rel(y, ref)); bad += rel(y, ref) > 1e-2
y = k.gemm(a[:, 200:], w[:, 200:], trans_b=True); ref = a[:, 200:].float() @ w[:, 200:].float().t()
print('col-slice offset', rel(y, ref)); bad += rel(y, ref) > 1e-2
out = torch.zeros(300, 1024, device='cuda', dtype=torch.bfloat16)
k.gemm(a, w, trans_b=True, out=out[:, 256:768]); print('out slice', rel(out[:, 256:768], a.float() @ w.float().t()), float(out[:, :256].abs().sum()))
k.gemm(a, w, trans_b=True, residual=out[:, 256:768], out=out[:, 256:768]); print('accumulate in place', rel(out[:, 256:768], 2 * (a.float() @ w.float().t())))
print('FAILURES', bad)
