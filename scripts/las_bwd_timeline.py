"""Where does ONE step of the LAS decoder's backward loop (BPTT, inside the replayed whole-step graph) spend its ~40 us?
One-thread stamp kernels (b200st_debug_stamp) are captured behind the attention backward and behind each of the three LSTM cell
backward kernels of a few middle steps -- the four fixed points of a step's dependent chain:
    attention bwd -> cell 2 bwd -> GEMM dG2 W_ih2 -> cell 1 bwd -> GEMM dG1 W_ih1 -> cell 0 bwd -> GEMM dG0 W_ih0 -> next step's attention bwd
The stamps sit on the chain themselves (~2 us each), so the absolute step is a little longer than in the unstamped graph.
    python scripts/las_bwd_timeline.py [replays]"""
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
FIRST, N = 10, 6                       # loop iterations FIRST .. FIRST + N - 1 of the 31 (counted in backward order)
buf = torch.zeros(4 * N + 4, dtype=torch.int64, device=dev)
k = K()
cnt = {'att': 0, 'cell': 0}
att, cell = k.las_attn_bwd, k.lstm_cell_bwd


def att_wrapped(*a, **kw):
    r = att(*a, **kw)
    i = cnt['att'] % 31 - FIRST
    if 0 <= i <= N:                    # one more than N: closes the last step
        k.debug_stamp(buf[4 * i:4 * i + 1]) if i < N else k.debug_stamp(buf[4 * N:4 * N + 1])
    cnt['att'] += 1
    return r


def cell_wrapped(*a, **kw):
    r = cell(*a, **kw)
    c = cnt['cell'] % 93
    i, j = c // 3 - FIRST, c % 3
    if 0 <= i < N:
        k.debug_stamp(buf[4 * i + 1 + j:4 * i + 2 + j])
    cnt['cell'] += 1
    return r


k.las_attn_bwd, k.lstm_cell_bwd = att_wrapped, cell_wrapped
g = GraphedTrainStep(model, tr, items, with_optimizer=True)
replays = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rows = []
for _ in range(replays):
    g()
    torch.cuda.synchronize()
    rows.append(buf.tolist())
t = torch.tensor(rows[1:], dtype=torch.float64)          # drop the first replay
seg = []
for i in range(N):
    b = t[:, 4 * i:4 * i + 4]
    nxt = t[:, 4 * (i + 1)]
    seg.append(torch.stack([b[:, 1] - b[:, 0], b[:, 2] - b[:, 1], b[:, 3] - b[:, 2], nxt - b[:, 3], nxt - b[:, 0]], 1))
seg = torch.stack(seg, 1).reshape(-1, 5)                 # [replays * N, 5] ns
med = seg.median(0).values / 1000.0
print('LAS decoder backward, one loop iteration inside the replayed graph, median us over %d iterations x %d replays:' % (N, len(rows) - 1))
print('  attention bwd done -> cell 2 bwd done        %6.2f   (join with dcv W_fb on the side stream + cell kernel)' % med[0])
print('  cell 2 bwd done    -> cell 1 bwd done        %6.2f   (GEMM dG2 W_ih2 + cell kernel)' % med[1])
print('  cell 1 bwd done    -> cell 0 bwd done        %6.2f   (GEMM dG1 W_ih1 + cell kernel)' % med[2])
print('  cell 0 bwd done    -> next attention bwd done %6.2f   (GEMM dG0 W_ih0[:, E:] into dcv + attention backward)' % med[3])
print('  whole iteration                              %6.2f' % med[4])
