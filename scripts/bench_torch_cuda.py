"""The 'reference GPU path' (SURVEY.md 8d): the reference's algorithm executed by STOCK PyTorch CUDA kernels (cuDNN LSTM,
cuBLAS, ATen softmax / LayerNorm / autograd) on the same B200, for context next to bench.py's numbers.  It runs the oracle
restatement (the same torch primitives at the same call sites as the reference) with parameters and inputs on cuda:0.
Measurement script only: nothing here is part of the product path.
    python scripts/bench_torch_cuda.py [--batch 64] [--frames 1000] [--steps 3] [--tf32] [--autocast]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from oracle import st_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--frames', type=int, default=1000)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--tf32', action='store_true')
ap.add_argument('--autocast', action='store_true', help='torch.autocast(bfloat16)')
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = args.tf32
torch.backends.cudnn.allow_tf32 = args.tf32
dev = torch.device('cuda', 0)
cfg = bench.st_config()
P = {k: v.to(dev).requires_grad_(True) for k, v in O.init_params(cfg, seed=333).items()}
data = O.synthetic_batch(cfg, args.batch, args.frames, seed=333)
_pps = torch.nn.utils.rnn.pack_padded_sequence            # wants its lengths on the host
torch.nn.utils.rnn.pack_padded_sequence = lambda x, lens, **kw: _pps(x, lens.cpu(), **kw)
_lstm = torch._VF.lstm                                    # cuDNN's backward needs the training-mode forward
torch._VF.lstm = lambda *a: _lstm(*a[:7], True, *a[8:])
torch.set_default_device(dev)            # the oracle builds its masks / index tensors on the default device
src, tgt, feats = data['src'].to(dev), data['tgt'].to(dev), data['acous_feats'].to(dev)
adam = torch.optim.Adam(list(P.values()), lr=1e-5)


def step():
    for v in P.values():
        v.grad = None
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=args.autocast):
        loss, _ = O.train_step_st(P, cfg, src, tgt, feats, data['acous_lens'])
    loss.backward()
    torch.nn.utils.clip_grad_norm_([v for v in P.values() if v.grad is not None], 1.0)
    adam.step()
    return loss


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.steps):
    loss = step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / args.steps * 1e3
print(json.dumps({'impl': 'stock PyTorch CUDA (oracle restatement on cuda:0)', 'metric': bench.METRIC, 'value': args.batch / (ms / 1e3),
                  'unit': bench.UNIT, 'ms_per_step': ms, 'batch': args.batch, 'frames': args.frames, 'tf32': args.tf32,
                  'autocast_bf16': args.autocast, 'loss': float(loss), 'torch': torch.__version__}))
