"""Per-bucket timeline of the data-parallel step: where the gradient all-reduce sits relative to backward
and the optimizer.  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29573 scripts/dp_timeline.py
Eager step with CUDA events (a graph replay cannot be instrumented), then the graph-replayed step with / without exchange."""
import os, sys
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
from b200st import runtime
from b200st import runtime as rt
from b200st.dp import GradAllReducer
from b200st.graph import GraphedTrainStep
from b200st.train_step import Trainer_ST
from modules.optim import Optimizer
from oracle import st_oracle as O
import bench

say = lambda *a: print(*a, flush=True) if rank == 0 else None
runtime.set_compute_dtype('bf16')
cfg = bench.st_config()
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333 + rank)
items = {'srcid': [host['src'].to(dev)], 'tgtid': [host['tgt'].to(dev)], 'acous_feat': [host['acous_feats'].to(dev)],
         'acouslen': host['acous_lens']}
bucket_mb = int(os.environ.get('BUCKET_MB', '32'))
red = GradAllReducer(model, bucket_bytes=bucket_mb << 20)
opt = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)
tr = Trainer_ST(use_gpu=True, batch_size=64, reducer=red, optimizer=opt)
for _ in range(3):
    tr._train_batch(model, items)
torch.cuda.synchronize(); dist.barrier()
# ---- eager timeline
red.trace = []
E = lambda: torch.cuda.Event(enable_timing=True)
e_start, e_bwd_done, e_red_done, e_opt_done = E(), E(), E(), E()
orig_finish = red.finish
def finish():
    e_bwd_done.record()
    orig_finish()
    e_red_done.record()
red.finish = finish
e_start.record()
tr._train_batch_device(model, items)
opt.step()
e_opt_done.record()
torch.cuda.synchronize()
model.zero_grad()
say(f'[eager, N={world}, bucket {bucket_mb} MB] backward done (weight-gradient GEMMs joined) at {e_start.elapsed_time(e_bwd_done):.3f} ms; '
    f'last all-reduce waited for at {e_start.elapsed_time(e_red_done):.3f} ms; optimizer done at {e_start.elapsed_time(e_opt_done):.3f} ms')
for i, (a, b, nbytes, nt) in enumerate(red.trace):
    say(f'   bucket {i}: {nbytes / 1e6:7.1f} MB in {nt:3d} tensors, all-reduce {e_start.elapsed_time(a):7.3f} -> {e_start.elapsed_time(b):7.3f} ms '
        f'({a.elapsed_time(b) * 1e3:6.0f} us, {nbytes / max(a.elapsed_time(b), 1e-6) / 1e6:6.1f} GB/s algbw)')
red.trace = None
red.finish = orig_finish

def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = E(), E()
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

for name, reducer, with_opt in (('no exchange, no optimizer', None, False), ('no exchange, optimizer', None, True),
                                ('exchange, no optimizer', red, False), ('exchange + optimizer (the bench step)', red, True)):
    model.zero_grad(set_to_none=True)
    red.arm(reducer is not None)          # the hooks live on the model: disarm them for the no-exchange variants
    t2 = Trainer_ST(use_gpu=True, batch_size=64, reducer=reducer, optimizer=opt)
    g = GraphedTrainStep(model, t2, items, with_optimizer=with_opt)
    say(f'[graph] {name}: {timed(lambda: g()):.3f} ms')
    del g
torch.cuda.synchronize(); dist.barrier()
os._exit(0)
