"""2-GPU NCCL check of the data-parallel step (run under torchrun, one rank per GPU):
   DP(2 ranks x half batch, GradAllReducer) == one process with minibatch_partition=2 (the reference's serial
   accumulation, trainer_st.py:225-290), fp32 compute, eager AND through the whole-step CUDA graph.
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 scripts/check_dp_nccl.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
from b200st import runtime
from b200st.dp import GradAllReducer
from b200st.graph import GraphedTrainStep
from b200st.train_step import Trainer_ST
from oracle import st_oracle as O
from helpers import build_model

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl', rank=rank, world_size=world)
dev = torch.device('cuda')
runtime.set_compute_dtype('fp32')
cfg = O.STConfig(enc_vocab_size=300, dec_vocab_size=300, enc_embedding_size=40, dec_embedding_size=40,
                 max_seq_len_src=10, max_seq_len_tgt=13, num_heads=4, dim_model=64, dim_feedforward=128,
                 enc_layers=2, dec_layers=2, acous_dim=24, acous_hidden_size=32)
torch.manual_seed(5)
P = O.init_params(cfg, seed=11)
B = 8
data = O.synthetic_batch(cfg, B, 96, seed=21)
lens = [int(n) for n in data['acous_lens']]
lens[0] = lens[B // 2] = max(lens)                   # both halves pad to the same feature length


def items(sl):
    return {'srcid': [data['src'][sl].to(dev)], 'tgtid': [data['tgt'][sl].to(dev)],
            'acous_feat': [data['acous_feats'][sl].to(dev)], 'acouslen': lens[sl]}


def grads(m):
    return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}


# serial reference semantics (every rank computes it; compared on rank 0)
m = build_model(cfg, P, device='cuda'); m.train()
Trainer_ST(use_gpu=True, batch_size=B, minibatch_partition=2)._train_batch_device(m, items(slice(0, B)))
ref = grads(m)

half = B // world
sl = slice(rank * half, (rank + 1) * half)
worst = {}
for mode in ('eager', 'graph'):
    m2 = build_model(cfg, P, device='cuda'); m2.train()
    red = GradAllReducer(m2, bucket_bytes=64 << 10)
    tr = Trainer_ST(use_gpu=True, batch_size=half, minibatch_partition=1, reducer=red)
    if mode == 'eager':
        tr._train_batch_device(m2, items(sl))
    else:
        g = GraphedTrainStep(m2, tr, items(sl))
        m2.zero_grad(set_to_none=False)
        g(items(sl))
    torch.cuda.synchronize()
    got = grads(m2)
    assert set(got) == set(ref), set(got) ^ set(ref)
    worst[mode] = max(float((got[n] - ref[n]).norm()) / (float(ref[n].norm()) + 1e-12) for n in ref)
    red.remove()
dist.barrier()
if rank == 0:
    print('DP2 vs minibatch_partition=2, worst relative gradient difference:', worst, flush=True)
    assert all(v < 1e-4 for v in worst.values()), worst
    print('OK', flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
