"""One replay of the whole-step CUDA graph (the thing bench.py times) after warm-up, for an ncu launch list of the GRAPH's
kernel nodes:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/one_step_graph.py
(ncu profiles the kernel nodes of a replayed graph one by one).  One `stamp_kernel` launch (b200st_debug_stamp) marks the start of the measured
replay; scripts/summarize_graph_launches.py aggregates what follows it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from b200st import runtime
from b200st.graph import GraphedTrainStep
from b200st.train_step import Trainer_ST
from modules.optim import Optimizer
from oracle import st_oracle as O
runtime.set_compute_dtype(sys.argv[1] if len(sys.argv) > 1 else 'bf16')
cfg = bench.st_config()
dev = torch.device('cuda')
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feats'].to(dev)],
         'acouslen': host['acous_lens']}
opt = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)
tr = Trainer_ST(use_gpu=True, batch_size=64, optimizer=opt)
g = GraphedTrainStep(model, tr, items, with_optimizer=True)
for _ in range(2):
    g()
torch.cuda.synchronize()
print('MARK replay begins', flush=True)
from b200st.kernels import K
marker = torch.zeros(1, dtype=torch.int64, device=dev)
K().debug_stamp(marker)
loss = g()
torch.cuda.synchronize()
print('loss', float(loss) if loss is not None else None)
