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
for backend, bname in ((0, 'default: tcgen05 (lstm_tc.cu)'), (3, 'register-resident warp MMA (lstm_rg.cu)')):
  k.set_blstm_backend(backend)
  print('backend', bname)
  for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    dout = torch.randn_like(out)
    e0.record()
    hs, acts, cs = k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2)
    e1.record()
    dg = k.blstm_bwd(dout, B * 4 * H, 4 * H, 2, acts, cs, wf, wr, lens, torch.bfloat16)
    e2.record()
    torch.cuda.synchronize()
    print(f'  T={T} B={B}: fwd {e0.elapsed_time(e1):.3f} ms ({e0.elapsed_time(e1) / T * 1e3:.2f} us/step)  '
            f'bwd {e1.elapsed_time(e2):.3f} ms ({e1.elapsed_time(e2) / T * 1e3:.2f} us/step)')

# ---- forward without the saved state (inference mode: only the layer output is written) = how much of a step is global stores
for backend, bname in ((0, 'tcgen05'), (3, 'register-resident HMMA')):
    k.set_blstm_backend(backend)
    for it in range(2):
        e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record(); k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2, save=False); e1.record()
        torch.cuda.synchronize()
    print(f'  {bname}: fwd without saved state {e0.elapsed_time(e1):.3f} ms ({e0.elapsed_time(e1) / T * 1e3:.2f} us/step)')
# ---- in-kernel timeline of steps 64..71 (cycles between recorded points, first CTA)
import ctypes
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
k.lib.b200st_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
# register-resident HMMA kernels: points = loop top, h landed, HMMA done, activations + cell done, sent, stores issued
k.set_blstm_backend(3)
hs, acts, cs = k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2)
for name, fn in (('fwd_rg', lambda: k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2)),
                 ('fwd_rg_nosave', lambda: k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2, save=False)),
                 ('bwd_rg', lambda: k.blstm_bwd(torch.randn_like(out), B * 4 * H, 4 * H, 2, acts, cs, wf, wr, lens, torch.bfloat16))):
    buf.zero_(); fn(); torch.cuda.synchronize()
    full = buf.cpu().view(8, 16)
    npts = int((full[0, :6] != 0).sum())
    if npts < 2:
        print(name, 'no timeline recorded'); continue
    if full[0, 6] != 0:
        print(name, 'tail detail (cycles after the send): x consume', int((full[:, 6] - full[:, 4]).median()), 'x loads issued', int((full[:, 7] - full[:, 6]).median()),
              'acts/cs stores', int((full[:, 8] - full[:, 7]).median()), 'out/hs stores', int((full[:, 5] - full[:, 8]).median()))
    tl = full[:, :npts]
    d = (tl[:, 1:] - tl[:, :-1]).float()
    step = (tl[1:, 0] - tl[:-1, 0]).float()
    print(name, 'cycles between points (median over steps 64..71):', [int(x) for x in d.median(0).values.tolist()],
          '| last point -> next h landed', int((tl[1:, 1] - tl[:-1, npts - 1]).float().median()), '| step period', int(step.median()),
          'periods', [int(x) for x in step.tolist()])
k.set_blstm_backend(0)
hs, acts, cs = k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2)
for name, fn, npts in (('fwd', lambda: k.blstm_fwd(xproj, wf, wr, lens, out, B * 4 * H, 4 * H, 2), 9),
                       ('bwd', lambda: k.blstm_bwd(torch.randn_like(out), B * 4 * H, 4 * H, 2, acts, cs, wf, wr, lens, torch.bfloat16), 8)):
    buf.zero_(); fn(); torch.cuda.synchronize()
    tl = buf.cpu().view(8, 16)[:, :npts]
    d = (tl[:, 1:] - tl[:, :-1]).float()
    step = (tl[1:, 0] - tl[:-1, 0]).float()
    full = buf.cpu().view(8, 16)
    if name == 'fwd':
        print('fwd issuer: wait_h', int((full[:,1]-full[:,0]).median()), 'fence', int((full[:,10]-full[:,1]).median()), 'mma_issue', int((full[:,11]-full[:,10]).median()), 'commit', int((full[:,2]-full[:,11]).median()), 'mma_done_wait', int((full[:,3]-full[:,2]).median()),
              '| epilogue: ld', int((full[:,4]-full[:,9]).median()), 'act', int((full[:,5]-full[:,4]).median()), 'bar', int((full[:,6]-full[:,5]).median()), 'cell+send', int((full[:,7]-full[:,6]).median()), 'stores', int((full[:,8]-full[:,7]).median()),
              '| mma_done->epilogue_start', int((full[:,9]-full[:,3]).median()), 'send->next_h_ready', int((full[1:,1]-full[:-1,7]).median()))
    print(name, 'per-step periods', [int(x) for x in step.tolist()])
    print(name, 'raw rows (relative to first point of step 64):')
    for r in full.tolist():
        print('   ', [int(x - full[0, 0].item()) if x else 0 for x in r[:12]])
    print(name, 'cycles between points (median over steps 64..71):', [int(x) for x in d.median(0).values.tolist()],
          'step period', int(step.median()))
k.lib.b200st_debug_timeline(None)
