#!/usr/bin/env python
"""Benchmark of the joint speech-translation training hot path (BASELINE.json metric: joint-ST train utt/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

One "step" = Seq2seq.forward_train(mode='ST') + masked NLL (trainer_st.py:268-288) + backward (+ gradient
mean-all-reduce when N > 1), on BASELINE.json configs[2]: synthetic 80-dim fbank, 1000 frames (padded to
1008 by the reference's rule), per-GPU batch 64, V=10k, d=512, 8 heads, 6+6 layers, E=200, H=256,
max_seq_len_src=32, target length 50.  Weak scaling: per-GPU batch is fixed, N=8 is the global-batch-512
config.  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = 'joint_st_train_utt_per_s'
UNIT = 'utt/s'


def st_config():
    from oracle.st_oracle import STConfig
    return STConfig(enc_vocab_size=10000, dec_vocab_size=10000, enc_embedding_size=200,
                    dec_embedding_size=200, max_seq_len_src=32, max_seq_len_tgt=50, num_heads=8,
                    dim_model=512, dim_feedforward=1024, enc_layers=6, dec_layers=6, acous_dim=80,
                    acous_hidden_size=256)


def workload(args):
    return {'workload': 'BASELINE.json configs[2]: joint ST embedding passing, pyramidal BLSTM enc -> dynamic+'
                        'static embedding mix -> TFEnc/TFDec 6+6, fwd+bwd',
            'per_gpu_batch': args.batch, 'global_batch': args.batch * args.gpus, 'frames': args.frames,
            'frames_padded': args.frames + 8 - args.frames % 8, 'acous_dim': 80, 'vocab': 10000,
            'dim_model': 512, 'heads': 8, 'layers': '6+6', 'max_seq_len_src': 32, 'tgt_len': 50,
            'step': 'Trainer_ST._train_batch: forward_train(ST) + masked NLL + backward' +
                    (' + grad all-reduce' if args.gpus > 1 else '') + ' + grad-norm clip(1.0) + Adam',
            'parallelism': f'dp{args.gpus}', 'dropout': 0.0}


# ------------------------------------------------------------------------------------------------
# algorithmic work: forward FLOPs per batch, de-duplicated key projection
# ------------------------------------------------------------------------------------------------
def algorithmic_flops(B, T_pad, F=80, H=256, D=512, FF=1024, V=10000, E=200, S=31, L=50, n_enc=6, n_dec=6):
    blstm = 0
    for layer in range(4):
        t = T_pad // (2 ** layer)
        i = F if layer == 0 else 4 * H
        blstm += 2 * B * t * (i + H) * 4 * H * 2
    tk = T_pad // 8
    las_step = 2 * B * ((E + 2 * D) * 4 * D + 2 * (2 * D) * 4 * D + 2 * tk * D + (2 * H + D) * D + D * V)
    las = S * las_step + 2 * B * tk * 2 * H * D          # key projection once
    mix = 2 * B * S * (E + D) * D

    def mha(lq, lk):
        return 2 * B * D * D * (2 * lq + 2 * lk) + 4 * B * lq * lk * D
    ffn = lambda l: 4 * B * l * D * FF
    tfenc = n_enc * (mha(S, S) + ffn(S))
    tfdec = n_dec * (mha(L, L) + mha(L, S) + ffn(L)) + 2 * B * L * E * D
    out = 2 * B * L * D * V
    fwd = blstm + las + mix + tfenc + tfdec + out
    return {'fwd': fwd, 'fwd_bwd': 3 * fwd, 'blstm_recurrent_fwd': sum(
        2 * B * (T_pad // 2 ** l) * H * 4 * H * 2 for l in range(4))}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# per-kernel-family timing with CUDA events on the launching stream
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    """Wraps methods of the kernel backend; records a CUDA event pair around each call of the chosen
    families.  Events are recorded on the current stream = the stream the kernels are launched on."""

    def __init__(self, backend, names):
        self.backend, self.names = backend, list(names)
        self.records = {n: [] for n in names}
        self.meta = {n: [] for n in names}
        self.big = {}              # (family, shape key) -> (bound method, args, kwargs, flops) of the >= 5e10 FLOP GEMMs
        self._orig = {}

    def __enter__(self):
        for n in self.names:
            orig = getattr(self.backend, n)
            self._orig[n] = orig

            def wrapped(*a, __orig=orig, __n=n, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = __orig(*a, **kw)
                e1.record()
                self.records[__n].append((e0, e1))
                if __n == 'gemm':
                    x, y = a[0], a[1]
                    ta, tb = kw.get('trans_a', False), kw.get('trans_b', False)
                    x2 = x[0] if x.dim() == 3 else x
                    y2 = y[0] if y.dim() == 3 else y
                    m, k = (x2.size(1), x2.size(0)) if ta else (x2.size(0), x2.size(1))
                    n_ = y2.size(0) if tb else y2.size(1)
                    self.meta[__n].append(2 * m * n_ * k * (x.size(0) if x.dim() == 3 else 1))
                    if self.meta[__n][-1] >= 5e10:
                        self.big.setdefault((__n, m, n_, k, ta, tb), (__orig, a, dict(kw), self.meta[__n][-1]))
                elif __n == 'gemm2':
                    self.meta[__n].append(2 * a[0].size(0) * (a[1].size(0) if kw.get('trans_b') else a[1].size(1)) *
                                          (a[0].size(1) + a[2].size(1)))
                    if self.meta[__n][-1] >= 5e10:
                        self.big.setdefault((__n, a[0].size(0), a[0].size(1) + a[2].size(1)),
                                            (__orig, a, dict(kw), self.meta[__n][-1]))
                elif __n in ('blstm_fwd',):
                    _, T, B, H4 = a[0].shape
                    self.meta[__n].append(2 * 2 * T * B * (H4 // 4) * H4)
                elif __n in ('blstm_bwd',):
                    _, T, B, H = a[5].shape
                    self.meta[__n].append(2 * 2 * T * B * H * 4 * H)
                return r
            setattr(self.backend, n, wrapped)
        return self

    def __exit__(self, *exc):
        for n, orig in self._orig.items():
            try:
                delattr(self.backend, n)          # restore the class method
            except AttributeError:
                setattr(self.backend, n, orig)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for n, evs in self.records.items():
            ms = [a.elapsed_time(b) for a, b in evs]
            out[n] = {'calls': len(ms), 'ms': sum(ms), 'flops': float(sum(self.meta[n])) if self.meta[n] else None}
        return out


ALL_FAMILIES = ['gemm', 'gemm2', 'layernorm_fwd', 'layernorm_bwd', 'mha_fwd', 'mha_bwd', 'lstm_cell_fwd', 'lstm_cell_bwd',
                'blstm_fwd', 'blstm_bwd', 'las_attn_fwd', 'las_attn_bwd', 'argmax_rows', 'las_update_lengths',
                'embedding_fwd', 'embedding_bwd', 'mix_gather_concat', 'log_softmax_fwd', 'log_softmax_bwd',
                'masked_nll_fwd', 'masked_nll_bwd', 'add', 'add_posenc', 'transpose01', 'cast', 'colsum',
                'relu_bwd', 'token_mask', 'length_mask', 'clip_adam_step']


# ------------------------------------------------------------------------------------------------
# the CPU arm: the reference's algorithm (oracle port, plain PyTorch fp32) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step(P, cfg, data, adam=None):
    from oracle import st_oracle as O
    for v in P.values():
        v.grad = None
    loss, _ = O.train_step_st(P, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
    loss.backward()
    if adam is not None:                     # Optimizer.step(): modules/optim.py:31-36
        torch.nn.utils.clip_grad_norm_([v for v in P.values() if v.grad is not None], 1.0)
        adam.step()
    return float(loss.detach())


def cpu_baseline(args, budget_s=25.0, steps=1, warmup=0):
    """Times `steps` oracle steps on a bounded sample of the workload (same shapes, smaller batch).
    The batch is chosen from a one-utterance probe so that the whole call stays within ~budget_s seconds."""
    from oracle import st_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = st_config()
    P = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=333).items()}
    adam = torch.optim.Adam(list(P.values()), lr=1e-5)
    probe = O.synthetic_batch(cfg, 1, args.frames, seed=1)
    t0 = time.perf_counter(); cpu_step(P, cfg, probe, adam); t_probe = time.perf_counter() - t0   # also warms the threads
    total_steps = steps + warmup
    # cost model: t(b) ~ t_probe * (0.5 + 0.5 * b)  (the 1890 serial LSTM steps have a large batch-independent part)
    b = 1
    while b < min(args.batch, 8) and t_probe * (0.5 + 0.5 * (b * 2)) * total_steps <= budget_s - t_probe:
        b *= 2
    if t_probe * total_steps > budget_s:        # even batch 1 blows the budget: the probe IS the measurement
        dt, b, steps_done = t_probe, 1, 1
    else:
        data = O.synthetic_batch(cfg, b, args.frames, seed=333)
        for _ in range(warmup):
            cpu_step(P, cfg, data, adam)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_step(P, cfg, data, adam)
        dt = (time.perf_counter() - t0) / steps
        steps_done = steps
    return {'value': b / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'configs[2] shapes ({args.frames} frames, V=10k, 6+6 layers), batch {b} of {args.batch}, '
                      f'{steps_done} timed step(s) of fwd+bwd+clip+Adam, fp32, torch.set_num_threads({cores}); the reference is '
                      f'pure Python/PyTorch and cannot travel to the GPU box, so its algorithm is timed through '
                      f'oracle/st_oracle.py (same torch primitives at the same call sites)',
            'ms_per_step': dt * 1e3, 'batch': b}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb = cpu_baseline(args, budget_s=120.0, steps=max(1, args.steps), warmup=max(0, args.warmup))
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload(args), 'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
def _profile_traffic(frames_padded, batch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the recurrence kernels from the committed
    `ncu --set full` capture (profiles/r01_blstm_ncu.json: one fwd + one bwd launch at T=1008, B=64), averaged over
    fwd+bwd and scaled to the AVERAGE launch of this run (the four pyramid layers run T, T/2, T/4, T/8 steps; the
    kernels' traffic is linear in T*B), like `achieved`; None if not captured."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_blstm_ncu.json')))
        ref = d['blstm_fwd_tc']
        mean_t = sum(frames_padded // 2 ** l for l in range(4)) / 4.0
        return d['dram_bytes_per_launch_avg'] * (mean_t * batch) / (ref['T'] * ref['B'])
    except Exception:
        return None


def build_model(cfg, device):
    from models.Seq2seq import Seq2seq
    torch.manual_seed(333)
    m = Seq2seq(cfg.enc_vocab_size, cfg.dec_vocab_size, share_embedder=False,
                enc_embedding_size=cfg.enc_embedding_size, dec_embedding_size=cfg.dec_embedding_size,
                max_seq_len_src=cfg.max_seq_len_src, max_seq_len_tgt=cfg.max_seq_len_tgt,
                num_heads=cfg.num_heads, dim_model=cfg.dim_model, dim_feedforward=cfg.dim_feedforward,
                enc_layers=cfg.enc_layers, dec_layers=cfg.dec_layers, embedding_dropout=0.0, dropout=0.0,
                acous_dim=cfg.acous_dim, acous_hidden_size=cfg.acous_hidden_size, mode='ST', load_mode='null')
    for mod in m.modules():
        if type(mod).__name__ == 'ScaledDotProductAttention':
            mod.dropout.p = 0.0
    return m.to(device).train()


def run_b200(args):
    import torch.distributed as dist
    from oracle import st_oracle as O
    from b200st import runtime
    from b200st.dp import GradAllReducer
    from b200st.kernels import K
    from b200st.train_step import Trainer_ST

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)'
    runtime.set_compute_dtype(args.dtype)
    cfg = st_config()
    model = build_model(cfg, device)
    reducer = GradAllReducer(model) if world > 1 else None
    from modules.optim import Optimizer
    # the reference's optimizer (trainer_base.py:422-426) at its warm-up starting rate (learning_rate_init)
    optimizer = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)
    trainer = Trainer_ST(use_gpu=True, batch_size=args.batch, minibatch_partition=1, reducer=reducer,
                         optimizer=optimizer)

    host = O.synthetic_batch(cfg, args.batch, args.frames, seed=333 + rank)
    pin = lambda t: t.pin_memory()
    batch_items = {'srcid': [pin(host['src'])], 'tgtid': [pin(host['tgt'])],
                   'acous_feat': [pin(host['acous_feats'])], 'acouslen': host['acous_lens'],
                   'srclen': [cfg.max_seq_len_src] * args.batch, 'tgtlen': [cfg.max_seq_len_tgt] * args.batch}
    h2d = sum(batch_items[k][0].numel() * batch_items[k][0].element_size() for k in ('srcid', 'tgtid', 'acous_feat')) + 4 * args.batch
    dev_items = dict(batch_items)
    for k in ('srcid', 'tgtid', 'acous_feat'):
        dev_items[k] = [batch_items[k][0].to(device)]

    def eager_step(items):
        out = trainer._train_batch(model, items)
        model.zero_grad(set_to_none=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    for _ in range(max(args.warmup, 3)):
        eager_step(dev_items)
    # ---- profiling pass (eager, untimed): every kernel family with CUDA events around each launch
    n_before = K().launch_count()
    with KernelTimer(K(), ALL_FAMILIES) as kt:
        eager_step(dev_items)
    launches_per_step = K().launch_count() - n_before
    fam = kt.summary()
    # Dominant kernel = the persistent BLSTM recurrence pair (largest single-kernel share of the step in the ncu launch
    # list, profiles/r01_launches_summary.txt; the eager per-family sums above overstate the many tiny GEMM launches
    # because each bracket then also contains host launch gaps).
    roof_names = ['blstm_fwd', 'blstm_bwd']
    roof = {n: fam[n] for n in roof_names}
    # secondary roofline: the tensor-bound GEMMs proper (BLSTM input projections and their input-gradient GEMMs =
    # gemm_tc_pair_kernel, cta_group::2): each distinct call of the step re-issued back to back on its own operands
    b_ms = b_fl = 0.0
    b_n, b_shapes = 0, []
    for key, (fn, a, kw, fl) in kt.big.items():
        for _ in range(3):
            fn(*a, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(10_000_000)
        e0.record()
        for _ in range(20):
            fn(*a, **kw)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20
        b_ms += t; b_fl += fl; b_n += 1
        b_shapes.append({'call': list(key), 'us': round(t * 1e3, 2), 'tflops': round(fl / t / 1e9, 1)})
    big_gemm = (b_ms, b_fl, b_n, b_shapes)
    del kt
    ms_eager = timed(lambda: eager_step(dev_items), args.steps)

    # ---- the measured step: one CUDA graph of forward_train + loss + backward (+ all-reduce)
    graphed = None
    ms_fb = None
    if not args.no_graph:
        try:
            from b200st.graph import GraphedTrainStep
            g_fb = GraphedTrainStep(model, trainer, dev_items)           # forward + backward only (north-star hot path)
            for _ in range(max(args.warmup, 3)):
                g_fb()
            ms_fb = timed(lambda: g_fb(), args.steps)
            del g_fb
            model.zero_grad(set_to_none=True)
            graphed = GraphedTrainStep(model, trainer, dev_items, with_optimizer=True)
        except Exception as ex:            # e.g. a collective that cannot be captured: stay eager, say so
            print(f'[bench] CUDA graph capture failed ({type(ex).__name__}: {ex}); timing the eager step', file=sys.stderr)
            graphed = None
            model.zero_grad(set_to_none=True)
    probe_param = model.out_tgt.weight
    probe_before = probe_param.detach().clone()
    sampler = ClockSampler(local)
    sampler.start()
    if graphed is not None:
        for _ in range(max(args.warmup, 3)):
            graphed()
        ms_dev = timed(lambda: graphed(), args.steps)                    # inputs resident in HBM

        # end to end through the public API: every step copies a batch from pinned host memory (H2D, on a copy stream,
        # double-buffered so that the transfer of batch i+1 hides under step i) and reads the loss back (D2H)
        graphed.prefetch(batch_items)

        def e2e_step():
            loss_dev = graphed.step_prefetched()                         # swap the staged batch in, replay the step
            graphed.prefetch(batch_items)                                # H2D of the next batch from pinned host memory
            return float(loss_dev)                                       # D2H read of the loss
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
    else:
        ms_dev = ms_eager
        ms_e2e = timed(lambda: eager_step(batch_items), args.steps)
    launches = launches_per_step * args.steps
    # nvidia-smi takes a moment to start and reports every 100 ms while K steps last ~70 ms: keep replaying the same step
    # (untimed) until the sampler has seen the GPU under this load for a few periods
    # (same count on every rank -- ms_dev is the all-reduced maximum -- because the step contains collectives)
    for _ in range(int(min(300, max(20, 1200.0 / max(ms_dev, 1e-3))))):
        if graphed is not None:
            graphed()
        else:
            eager_step(dev_items)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    # the timed replays really stepped the weights (clip + Adam is inside the captured step)
    optimizer_applied = bool((probe_param.detach() != probe_before).any())
    opt_steps = float(optimizer._fused._step) if optimizer._fused is not None else 0.0

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops_sustained)' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        # secondary roofline: the tensor-bound GEMMs proper -- every `gemm` / `gemm2` call of >= 5e10 FLOP in the eager
        # pass (the BLSTM input projections and their input-gradient GEMMs: gemm_tc_pair_kernel, cta_group::2)
        big_ms, big_fl, big_n, big_shapes = big_gemm
        r_calls = sum(roof[n]['calls'] for n in roof_names)
        r_ms = sum(roof[n]['ms'] for n in roof_names)
        r_flops = sum(roof[n]['flops'] or 0 for n in roof_names)
        achieved = r_flops / max(r_ms, 1e-9) / 1e9          # TFLOP/s
        t_pad = args.frames + 8 - args.frames % 8
        alg = algorithmic_flops(args.batch, t_pad)
        total_units = args.batch * world
        line = {
            'metric': METRIC, 'value': total_units / (ms_dev / 1e3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.dtype == 'bf16' else 'f32', 'data': 'synthetic',
            'config': dict(workload(args), l2='per-step working set (saved activations > 1 GB) exceeds the 126 MB L2; no flush needed',
                           step_share_ms={k: round(v['ms'], 3) for k, v in fam.items() if v['ms'] > 0.01},
                           algorithmic_tflop_per_step=alg['fwd_bwd'] / 1e12,
                           whole_step_tflops=alg['fwd_bwd'] / 1e12 / (ms_dev / 1e3)),
            'e2e': {'value': total_units / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e,
                    'how': 'GraphedTrainStep.prefetch() + step_prefetched(): per step one H2D copy of fbank + ids + lengths '
                           'from pinned host memory (copy stream, double-buffered: batch i+1 transfers while step i runs) '
                           'and one D2H read of the loss'},
            'gpu_launches': int(launches), 'gpu_launches_per_step': int(launches_per_step),
            'execution': ('one CUDA graph replay per step' if graphed is not None else 'eager launches'),
            'fwd_bwd_only': (None if ms_fb is None else {
                'ms_per_step': ms_fb, 'value': total_units / (ms_fb / 1e3), 'unit': UNIT,
                'what': 'the same graph without the optimizer step (forward + loss + backward' +
                        (' + grad all-reduce)' if world > 1 else ')')}),
            'eager_ms_per_step': ms_eager,
            'optimizer': {'kind': 'fused grad-norm clip (1.0) + Adam (lr 1e-5), inside the timed step',
                          'weights_changed_during_timed_region': optimizer_applied, 'adam_step_count': opt_steps},
            'clocks': clocks,
            'roofline': {'kernel': '+'.join(roof_names), 'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf,
                         'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': _profile_traffic(args.frames + 8 - args.frames % 8, args.batch), 'peak_source': peak_src,
                         'us_per_time_step': 1e3 * r_ms / (2 * sum((args.frames + 8 - args.frames % 8) // 2 ** l for l in range(4))),
                         'launches': r_calls, 'avg_launch_ms': r_ms / max(r_calls, 1),
                         'timed': 'CUDA events around every launch of this kernel family on its stream, eager pass of the '
                                  'same step in this run (the timed region replays the step as one CUDA graph)',
                         'note': 'recurrent-GEMM FLOPs 2*2dirs*T*B*H*4H per launch; this kernel is bound by the '
                                 'serial time-step chain (latency), not by tensor throughput'},
        }
        if big_n:
            peak_b = peaks.get('bf16_tflops', 1650.0)
            line['roofline_gemm'] = {
                'kernel': 'gemm_tc_pair_kernel (tcgen05 cta_group::2, 256x256 tile per SM pair): every GEMM of the step with '
                          '>= 5e10 FLOP -- BLSTM input projections, their input-gradient GEMMs and the largest weight '
                          'gradient (split-K over the SM pairs)',
                'bound': 'tensor', 'achieved': big_fl / max(big_ms, 1e-9) / 1e9, 'peak': peak_b, 'unit': 'TFLOP/s',
                'frac': big_fl / max(big_ms, 1e-9) / 1e9 / peak_b, 'launches': big_n,
                'avg_launch_ms': big_ms / big_n, 'shapes': big_shapes, 'peak_source': 'measured (MEASURED_PEAKS.json bf16_tflops, burst: '
                'each launch is bracketed alone)' if peaks else 'fallback', 'traffic': None,
                'timed': 'every distinct >= 5e10-FLOP GEMM call of the step re-issued 20x back to back on its own operands, CUDA '
                         'events around the batch (an eager bracket would include host launch gaps)'}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args)
            line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
        emit(line)
    if world > 1:
        # A process group whose collectives were captured into a CUDA graph can block in teardown; results are
        # already printed, so synchronise, drop the graph and leave without running NCCL's destructors.
        del graphed
        torch.cuda.synchronize()
        dist.barrier()
        sys.stderr.flush()
        os._exit(0)


_REAL_STDOUT = None


def _capture_stdout():
    """Only the final JSON line may reach stdout (NCCL / libraries print banners there): route fd 1 to stderr for the
    duration of the run and keep a handle on the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--dtype', default=os.environ.get('B200ST_BENCH_DTYPE', 'bf16'), choices=['bf16', 'fp32'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--frames', type=int, default=1000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
