"""In-kernel timeline of the persistent LAS decoder forward at the configs[2] shape (B 64, Tk 126, V 10k, S 31).
    python scripts/profile_las_decoder.py"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import bench
from b200st import runtime
from b200st.kernels import K
runtime.set_compute_dtype('bf16')
cfg = bench.st_config()
m = bench.build_model(cfg, torch.device('cuda')).eval()
B, Tk, S = 64, 126, 31
enc = (0.5 * torch.randn(B, Tk, 512, device='cuda')).to(torch.bfloat16)
klens = torch.full((B,), Tk, dtype=torch.int32, device='cuda')
dec = m.las.decoder
tl = torch.zeros(S * 16 + 1, dtype=torch.int64, device='cuda')
with torch.no_grad():
    for _ in range(3):
        dec.forward_device(enc, klens, need_logps=False)
    torch.cuda.synchronize()
    K().lib.b200st_las_decoder_timeline(ctypes.c_void_p(tl.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec.forward_device(enc, klens, need_logps=False); e1.record()
    torch.cuda.synchronize()
    K().lib.b200st_las_decoder_timeline(None)
    for persistent in (True, False):
        runtime.las_persistent(persistent)
        dec.forward_device(enc, klens, need_logps=False)
        torch.cuda.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(5):
            dec.forward_device(enc, klens, need_logps=False)
        e3.record(); torch.cuda.synchronize()
        print(f'forward_device persistent={persistent}: {e2.elapsed_time(e3) / 5:.3f} ms (eager, incl. key projection / table GEMM / transposes)')
t = tl.cpu().tolist()
start = t[S * 16]
rows = [t[s * 16:s * 16 + 16] for s in range(S)]
print(f'kernel start -> step 0 start (weight staging): {(rows[0][0] - start) / 1e3:.1f} us; last step end - start: {(rows[-1][6] - start) / 1e3:.1f} us')
names = ['tokens+L0', 'L1', 'L2', 'attention', 'att barrier', 'ffn', 'vocab+argmax']
import statistics
seg = lambda r: [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[7] - r[3], r[4] - r[7], r[5] - r[4], r[6] - r[5]]
med = [statistics.median(seg(r)[i] for r in rows[2:]) for i in range(7)]
print('median ns per phase (steps 2..): ' + ', '.join(f'{n} {v:.0f}' for n, v in zip(names, med)) + f' | step {sum(med):.0f}')
fine = lambda r: [r[8] - r[1], r[9] - r[8], r[10] - r[9], r[11] - r[10], r[2] - r[11]]
medf = [statistics.median(fine(r)[i] for r in rows[2:]) for i in range(5)]
print('layer-1 -> layer-2 hand-over, median ns: fresh-half GEMM %d, cell %d, arrive (sync + fence + atomic) %d, recurrent-half GEMM %d, barrier wait %d' % tuple(medf))
