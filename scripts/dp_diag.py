"""Where does the data-parallel step lose time?  Run under torchrun (one rank per GPU):
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 scripts/dp_diag.py
Prints (rank 0): raw NCCL all-reduce time for the gradient volume (one flat buffer / 32 MB buckets / per-parameter
coalesced groups), then the whole-step CUDA-graph time with and without the gradient exchange, for several bucket
sizes.  Times are CUDA events, max over ranks."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
local = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# ---- 1. raw collectives on the gradient volume (69.5 M fp32 = 278 MB)
n = 69_545_104
flat = torch.zeros(n, device=dev)
ms = timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
say(f'[raw] one flat all-reduce of {n * 4 / 1e6:.0f} MB: {ms:.3f} ms  (algbw {n * 4 / ms / 1e6:.1f} GB/s)')
for mb in (8, 32, 64):
    chunks = list(flat.split(mb * (1 << 20) // 4))
    ms = timed(lambda: [dist.all_reduce(c, op=dist.ReduceOp.AVG) for c in chunks])
    say(f'[raw] {len(chunks)} x {mb} MB all-reduces: {ms:.3f} ms')

from b200st import runtime
from b200st.dp import GradAllReducer
from b200st.graph import GraphedTrainStep
from b200st.train_step import Trainer_ST
from oracle import st_oracle as O
import bench

runtime.set_compute_dtype('bf16')
cfg = bench.st_config()
model = bench.build_model(cfg, dev)
params = [p for p in model.parameters() if p.requires_grad]
gl = [torch.zeros_like(p) for p in params]


def coalesced():
    with dist._coalescing_manager(None, device=dev, async_ops=True) as cm:
        for g in gl:
            dist.all_reduce(g, op=dist.ReduceOp.AVG)
    cm.wait()


ms = timed(coalesced)
say(f'[raw] ONE coalesced group over {len(gl)} parameter tensors: {ms:.3f} ms')

host = O.synthetic_batch(cfg, 64, 1000, seed=333 + rank)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feats'].to(dev)],
         'acouslen': host['acous_lens']}


def graph_ms(reducer):
    tr = Trainer_ST(use_gpu=True, batch_size=64, minibatch_partition=1, reducer=reducer)
    g = GraphedTrainStep(model, tr, items)
    ms = timed(lambda: g(), steps=8, warm=3)
    del g
    model.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    return ms


say(f'[step] graph, no gradient exchange: {graph_ms(None):.3f} ms')
for mb in (32, 8, 128, 1024):
    red = GradAllReducer(model, bucket_bytes=mb << 20)
    say(f'[step] graph, GradAllReducer bucket {mb} MB: {graph_ms(red):.3f} ms')
    red.remove()
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
