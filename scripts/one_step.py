"""One eager cfg-3 training step after N warm-up steps (for ncu launch lists): python scripts/one_step.py [warm] [dtype]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from b200st import runtime
from oracle import st_oracle as O
from b200st.train_step import Trainer_ST
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 1
runtime.set_compute_dtype(sys.argv[2] if len(sys.argv) > 2 else 'bf16')
cfg = bench.st_config()
dev = torch.device('cuda')
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feat' + 's'].to(dev)],
         'acouslen': host['acous_lens']}
from modules.optim import Optimizer
opt = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)     # trainer_base.py:422-426
tr = Trainer_ST(use_gpu=True, batch_size=64, optimizer=opt)
for i in range(warm + 1):
    if i == warm:
        torch.cuda.synchronize(); print('MARK step begins', flush=True)
        marker = torch.zeros(7, device=dev)            # a recognisable 7-element fill marks the step start in the trace
    loss = tr._train_batch(model, items)
    model.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print('loss', loss)
