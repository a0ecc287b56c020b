"""Where does the REPLAYED whole-step CUDA graph spend its time?  One-thread stamp kernels (b200st_debug_stamp: %globaltimer)
are captured at the section boundaries of the step -- forward calls are wrapped, backward boundaries are tensor gradient hooks --
and read back after each replay.  Unlike scripts/section_times.py (each section captured alone) this is the real graph with its
side branches.    python scripts/graph_timeline.py [replays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from b200st import runtime as rt
from b200st.graph import GraphedTrainStep
from b200st.kernels import K
from b200st.train_step import Trainer_ST
from modules.optim import Optimizer
from oracle import st_oracle as O

rt.set_compute_dtype('bf16')
cfg = bench.st_config()
dev = torch.device('cuda')
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feats'].to(dev)],
         'acouslen': host['acous_lens']}
opt = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)
tr = Trainer_ST(use_gpu=True, batch_size=64, optimizer=opt)
buf = torch.zeros(48, dtype=torch.int64, device=dev)
k = K()
NAMES = ['step start', 'acoustic encoder fwd done', 'LAS decoder fwd done', 'mix + Transformer fwd + loss done',
         None, None, 'Transformer bwd done (gradient of the dynamic embedding ready)', 'LAS decoder bwd done (gradient of the encoder output ready)',
         'backward chain done (autograd returned)', 'deferred weight-gradient work joined', 'optimizer done']


def stamp(i):
    k.debug_stamp(buf[i:i + 1])


def hooked(t, i):
    if torch.is_tensor(t) and t.requires_grad:
        t.register_hook(lambda g: (stamp(i), g)[1])
    return t


enc_fwd = model.las.encoder.forward
def enc_wrapped(*a, **kw):
    stamp(0)
    y = enc_fwd(*a, **kw)
    stamp(1)
    return hooked(y, 7)
model.las.encoder.forward = enc_wrapped
acous = model._encoder_acous
def acous_wrapped(*a, **kw):
    r = acous(*a, **kw)
    stamp(2)
    return (hooked(r[0], 6),) + tuple(r[1:])
model._encoder_acous = acous_wrapped
fwd_train = model.forward_train
def fwd_wrapped(*a, **kw):
    r = fwd_train(*a, **kw)
    stamp(3)
    return r
model.forward_train = fwd_wrapped
join = rt.join_deferred
def join_wrapped():
    stamp(8)
    join()
    stamp(9)
rt.join_deferred = join_wrapped
import b200st.train_step as TS
TS.rt.join_deferred = join_wrapped
ostep = opt.step
def ostep_wrapped(*a, **kw):
    r = ostep(*a, **kw)
    stamp(10)
    return r
opt.step = ostep_wrapped

# finer: every recurrence launch of the backward pass (slots 16 + 2 i, 17 + 2 i) and of the forward pass (32 + 2 i, 33 + 2 i)
cnt = {'b': 0, 'f': 0}
bb, bf_ = k.blstm_bwd, k.blstm_fwd
def blstm_bwd_wrapped(*a, **kw):
    i = cnt['b'] % 4; cnt['b'] += 1
    stamp(16 + 2 * i); r = bb(*a, **kw); stamp(17 + 2 * i)
    return r
def blstm_fwd_wrapped(*a, **kw):
    i = cnt['f'] % 4; cnt['f'] += 1
    stamp(32 + 2 * i); r = bf_(*a, **kw); stamp(33 + 2 * i)
    return r
k.blstm_bwd, k.blstm_fwd = blstm_bwd_wrapped, blstm_fwd_wrapped

g = GraphedTrainStep(model, tr, items, with_optimizer=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rows = []
for _ in range(3):
    g()
for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g(); e1.record()
    torch.cuda.synchronize()
    rows.append((buf.cpu().tolist(), e0.elapsed_time(e1)))
idx = [i for i, nme in enumerate(NAMES) if nme]
print(f'whole-step graph, {n} replays (us since the first stamp; replay time by CUDA events in ms):')
for t, ms in rows:
    print('  ' + ' '.join(f'{(t[i] - t[0]) / 1e3:8.1f}' for i in idx) + f'   | {ms:.3f} ms')
import statistics
med = [statistics.median((t[i] - t[0]) / 1e3 for t, _ in rows) for i in idx]
print('median section lengths (us):')
for a, b, m0, m1 in zip(idx[:-1], idx[1:], med[:-1], med[1:]):
    print(f'  {m1 - m0:8.1f}  until: {NAMES[b]}')

t = rows[-1][0]
print('recurrence launches inside the graph (last replay): us from the stamp in front of the launch to the stamp behind it')
print('  forward  (layers 1..4): ' + ', '.join(f'{(t[33 + 2 * i] - t[32 + 2 * i]) / 1e3:.1f}' for i in range(4)))
print('  backward (layers 4..1): ' + ', '.join(f'{(t[17 + 2 * i] - t[16 + 2 * i]) / 1e3:.1f}' for i in range(4)))
print('  backward: between a recurrence and the next one (input-gradient GEMM etc.): ' +
      ', '.join(f'{(t[16 + 2 * (i + 1)] - t[17 + 2 * i]) / 1e3:.1f}' for i in range(3)))
print(f'  gradient of the encoder output ready -> first backward recurrence launched: {(t[16] - t[7]) / 1e3:.1f};  last backward recurrence '
      f'done -> autograd returned: {(t[8] - t[23]) / 1e3:.1f}')
