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
        self.calls = {}            # family -> [(bound method, args, kwargs)] of the roofline kernel's launches, for re-timing
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
                    self.calls.setdefault(__n, []).append((__orig, a, dict(kw)))
                elif __n in ('blstm_bwd',):
                    _, T, B, H = a[5].shape
                    self.meta[__n].append(2 * 2 * T * B * H * 4 * H)
                    self.calls.setdefault(__n, []).append((__orig, a, dict(kw)))
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


ALL_FAMILIES = ['gemm', 'gemm2', 'gemm_ln', 'gemm_lnbwd', 'las_decoder_fwd', 'las_stack_grad', 'layernorm_fwd', 'layernorm_bwd', 'mha_fwd', 'mha_bwd', 'lstm_cell_fwd', 'lstm_cell_bwd',
                'blstm_fwd', 'blstm_bwd', 'las_attn_fwd', 'las_attn_bwd', 'argmax_rows', 'las_update_lengths',
                'embedding_fwd', 'embedding_bwd', 'mix_gather_concat', 'log_softmax_fwd', 'log_softmax_bwd',
                'masked_nll_fwd', 'masked_nll_bwd', 'add', 'add_posenc', 'transpose01', 'cast', 'colsum',
                'relu_bwd', 'token_mask', 'length_mask', 'clip_adam_step']


# ------------------------------------------------------------------------------------------------
# the CPU arm: the reference's algorithm (oracle port, plain PyTorch fp32) on the host cores
# ------------------------------------------------------------------------------------------------
REF_STAGED = os.path.join(ROOT, 'oracle', '_ref')


def _reference_root():
    """The UNMODIFIED reference tree: staged by oracle/build_ref.sh (travels to the GPU box), else the build container's."""
    for cand in (REF_STAGED, '/root/reference'):
        if os.path.isfile(os.path.join(cand, 'trainer', 'trainer_st.py')):
            return cand
    return None


def _reference_step_fn(ref_root, cfg, batch, frames):
    """step() = the reference's OWN `Trainer_ST._train_batch` (trainer/trainer_st.py:211-299: forward_train('ST') + NLLLoss
    + backward + Optimizer.step() = clip_grad_norm_ + Adam) on the reference's own `models.Seq2seq.Seq2seq`, CPU, fp32.
    Import shims of SURVEY.md 8c only (absent third-party packages, the hard-coded .npy); none touches arithmetic."""
    import tempfile
    import types
    import numpy as np
    from oracle import st_oracle as O                      # synthetic_batch only: the same seeded inputs as the GPU arm
    sys.dont_write_bytecode = True
    for p in (PKG, os.path.join(ROOT, 'tests')):           # this repo's models/ modules/ must NOT shadow the reference's
        while p in sys.path:
            sys.path.remove(p)
    assert 'models' not in sys.modules and 'modules' not in sys.modules
    sys.path.insert(0, ref_root)
    for name in ['bpemb', 'matplotlib', 'matplotlib.pyplot', 'torchtext']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['bpemb'].BPEmb = object
    sys.modules['matplotlib'].use = lambda *a, **k: None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    np_load = np.load
    np.load = lambda p, *a, **k: (np.zeros(cfg.dim_model, np.float32) if str(p).endswith('dyn_emb_ave.npy')
                                  else np_load(p, *a, **k))
    from models.Seq2seq import Seq2seq
    from modules.optim import Optimizer
    from trainer.trainer_st import Trainer_ST
    assert os.path.realpath(sys.modules['models.Seq2seq'].__file__).startswith(os.path.realpath(ref_root))
    torch.manual_seed(333)
    model = Seq2seq(cfg.enc_vocab_size, cfg.dec_vocab_size, share_embedder=False,
                    enc_embedding_size=cfg.enc_embedding_size, dec_embedding_size=cfg.dec_embedding_size,
                    max_seq_len_src=cfg.max_seq_len_src, max_seq_len_tgt=cfg.max_seq_len_tgt, num_heads=cfg.num_heads,
                    dim_model=cfg.dim_model, dim_feedforward=cfg.dim_feedforward, enc_layers=cfg.enc_layers,
                    dec_layers=cfg.dec_layers, embedding_dropout=0.0, dropout=0.0, acous_dim=cfg.acous_dim,
                    acous_hidden_size=cfg.acous_hidden_size, mode='ST', load_mode='null')
    for mod in model.modules():                            # the hidden attention dropout (layers.py:207), like the GPU arm
        if type(mod).__name__ == 'ScaledDotProductAttention':
            mod.dropout.p = 0.0
    model.train()
    t = Trainer_ST(expt_dir=tempfile.mkdtemp(prefix='b200st_ref_'), load_dir=None, load_mode='null', batch_size=batch,
                   use_gpu=False, learning_rate=1e-5, learning_rate_init=1e-5, lr_warmup_steps=0, max_grad_norm=1.0,
                   loss_coeff={'nll_asr': 1.0, 'nll_mt': 1.0, 'nll_st': 1.0}, minibatch_partition=1)
    t.optimizer = Optimizer(torch.optim.Adam(model.parameters(), lr=1e-5), max_grad_norm=1.0)   # trainer_base.py:420-426

    def make_items(b, seed):
        d = O.synthetic_batch(cfg, b, frames, seed=seed)
        return {'srcid': [d['src']], 'srclen': [cfg.max_seq_len_src] * b, 'tgtid': [d['tgt']],
                'tgtlen': [cfg.max_seq_len_tgt] * b, 'acous_feat': [d['acous_feats']],
                'acouslen': [torch.tensor([n]) for n in d['acous_lens']]}

    def step(items):
        t.minibatch_size = items['srcid'][0].size(0)
        return float(t._train_batch(model, items, None, 0, 1)['nll_loss_de'])
    return step, make_items


def _port_step_fn(cfg, frames):
    """Fallback when no reference tree is at hand: the oracle restatement (same torch primitives at the same call sites)."""
    from oracle import st_oracle as O
    P = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=333).items()}
    adam = torch.optim.Adam(list(P.values()), lr=1e-5)

    def step(data):
        for v in P.values():
            v.grad = None
        loss, _ = O.train_step_st(P, cfg, data['src'], data['tgt'], data['acous_feats'], data['acous_lens'])
        loss.backward()
        torch.nn.utils.clip_grad_norm_([v for v in P.values() if v.grad is not None], 1.0)   # modules/optim.py:31-36
        adam.step()
        return float(loss.detach())
    return step, (lambda b, seed: O.synthetic_batch(cfg, b, frames, seed=seed))


def cpu_reference(args, batch, steps, warmup, budget_s):
    """Times the reference's CPU implementation of the step at `batch` utterances per step, all host cores.  As many of
    the requested warm-up + timed steps as fit into ~budget_s seconds are run (at least one timed step; a one-utterance
    probe warms the thread pool and sizes the plan)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = st_config()
    ref_root = _reference_root()
    if ref_root is not None:
        step, make = _reference_step_fn(ref_root, cfg, batch, args.frames)
        kind, what = 'reference', (f"the UNMODIFIED reference ({'oracle/_ref' if ref_root == REF_STAGED else ref_root}): its own "
                                   f"Trainer_ST._train_batch on its own models.Seq2seq")
    else:
        step, make = _port_step_fn(cfg, args.frames)
        kind, what = 'port', 'oracle/st_oracle.py (no reference tree staged: run oracle/build_ref.sh in the build container)'
    t0 = time.perf_counter(); step(make(1, 1)); t_probe = time.perf_counter() - t0
    items = make(batch, 333)
    t0 = time.perf_counter(); step(items); t_first = time.perf_counter() - t0       # doubles as warm-up when one is requested
    remaining = budget_s - t_probe - t_first
    if warmup >= 1 and remaining >= t_first:       # the first full step was the warm-up; time what still fits
        n = int(max(1, min(steps, remaining // max(t_first, 1e-9))))
        t0 = time.perf_counter()
        for _ in range(n):
            step(items)
        dt, timed, warmed = (time.perf_counter() - t0) / n, n, 1
    else:                                          # budget allows ONE full-size step: it is the measurement (cold)
        dt, timed, warmed = t_first, 1, 0
    return {'value': batch / dt, 'unit': UNIT, 'cores': cores, 'kind': kind, 'ms_per_step': dt * 1e3, 'batch': batch,
            'timed_steps': timed, 'warmup_steps': warmed,
            'sample': f'configs[2] shapes ({args.frames} frames, V=10k, 6+6 layers), batch {batch} of {args.batch} per step, '
                      f'{timed} timed step(s) after {warmed} warm-up step(s) (requested {steps}+{warmup}; bounded to ~{budget_s:.0f} s '
                      f'of CPU work), fwd + loss + bwd + clip_grad_norm_ + Adam, fp32, torch.set_num_threads({cores}); {what}'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cb = cpu_reference(args, args.batch, max(1, args.steps), max(0, args.warmup), args.ref_budget)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': cb['timed_steps'], 'warmup': cb['warmup_steps'], 'requested': {'steps': args.steps, 'warmup': args.warmup},
            'ms_per_step': cb['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload(args), 'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


def cpu_baseline_subprocess(args, batch=16, budget_s=30.0):
    """The `cpu_baseline` object of the GPU arm's line: the reference arm at a bounded batch in a fresh process (the
    reference's `models` / `modules` packages cannot share a process with this repo's)."""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--batch', str(batch), '--frames', str(args.frames),
           '--steps', '1', '--warmup', '0', '--ref-budget', str(budget_s)]
    env = dict(os.environ)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT'):
        env.pop(k, None)
    env['CUDA_VISIBLE_DEVICES'] = ''
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    cb = line['cpu_baseline']
    cb['sample'] = cb['sample'].replace(f'batch {batch} of {batch}', f'batch {batch} of {args.batch}')
    return cb


def stock_torch_cuda(args):
    """The reference's algorithm on STOCK PyTorch CUDA kernels (cuDNN LSTM, cuBLAS, ATen) on this GPU, same shapes, fp32
    with TF32 off: scripts/bench_torch_cuda.py in a fresh process.  Context for the GPU-vs-GPU comparison; None on failure."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'bench_torch_cuda.py'), '--batch', str(args.batch),
                            '--frames', str(args.frames), '--steps', '3'], capture_output=True, text=True, timeout=600)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return {'value': d['value'], 'unit': UNIT, 'ms_per_step': d['ms_per_step'], 'what': d['impl'] + ', fp32 (TF32 off), eager, '
                'fwd + loss + bwd + clip_grad_norm_ + Adam, 3 timed steps after 2 warm-ups, wall clock around a synchronize'}
    except Exception as ex:
        return {'value': None, 'error': f'{type(ex).__name__}: {ex}'[:200]}


# ------------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------------
def hbm_kernels(device, B, cfg, reps=10):
    """Achieved HBM GB/s of the bandwidth-bound kernels north_star names (loss, LayerNorm, the mix), each launched ALONE at
    its configs[2] shape with the L2 flushed before every launch (a 512 MB memset), CUDA events around the launch, median
    of `reps`.  `achieved` = algorithmic bytes / that time; the small kernels (a few MB) are launch-latency bound, which
    is exactly what the figure shows."""
    from b200st.kernels import K
    k = K()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = peaks.get('hbm_gbs', 6550.0)
    dt = torch.bfloat16 if runtime_dtype() == 'bf16' else torch.float32
    es = 2 if dt == torch.bfloat16 else 4
    L, S, D, V, E = cfg.max_seq_len_tgt, cfg.max_seq_len_src - 1, cfg.dim_model, cfg.dec_vocab_size, cfg.enc_embedding_size
    rows = B * L
    g = torch.Generator(device='cpu').manual_seed(1)
    logits = torch.randn(rows, V, generator=g).to(device, dt)
    target = torch.randint(5, V, (rows,), generator=g).to(device)
    mask = torch.ones(rows, dtype=torch.uint8, device=device)
    scale = torch.full((1,), 1.0 / rows, device=device)
    x = torch.randn(rows, D, generator=g).to(device, dt)
    dy = torch.randn(rows, D, generator=g).to(device, dt)
    gamma, beta = torch.ones(D, device=device), torch.zeros(D, device=device)
    _, mean, rstd = k.layernorm_fwd(x, gamma, beta, 1e-6)
    ids = torch.randint(5, V, (B * S,), generator=g).to(device)
    table = torch.randn(V, E, generator=g).to(device)
    dyn = torch.randn(B * S, D, generator=g).to(device, dt)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    cases = {
        'loss': ('softmax_nll_fused_vec_kernel (fused softmax + masked NLL + dlogits)', 2 * rows * V * es,
                 lambda: k.softmax_nll_fused(logits, target, mask, scale)),
        'layernorm_fwd': ('layernorm_fwd_kernel', 2 * rows * D * es, lambda: k.layernorm_fwd(x, gamma, beta, 1e-6)),
        'layernorm_bwd': ('layernorm_bwd_reg_kernel (dx + per-CTA dgamma/dbeta partial sums)', 3 * rows * D * es,
                          lambda: k.layernorm_bwd_partial(dy, x, gamma, mean, rstd)),
        'mix': ('mix_gather_concat_kernel (static rows gathered next to the dynamic embedding)',
                B * S * (E * 4 + D * es + (E + D) * es), lambda: k.mix_gather_concat(ids, table, dyn)),
    }
    out = {}
    for name, (kern, nbytes, fn) in cases.items():
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        out[name] = {'kernel': kern, 'bound': 'hbm', 'achieved': nbytes / ms / 1e6, 'peak': peak, 'unit': 'GB/s',
                     'frac': nbytes / ms / 1e6 / peak, 'algorithmic_bytes': nbytes, 'us': round(ms * 1e3, 2)}
    out['how'] = ('each kernel launched alone at its configs[2] shape, L2 flushed (512 MB memset) before every launch, CUDA '
                  f'events around the launch, median of {reps}; peak = MEASURED_PEAKS.json hbm_gbs')
    return out


def runtime_dtype():
    from b200st import runtime
    return 'bf16' if runtime.compute_dtype() == torch.bfloat16 else 'fp32'


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
    roof = {n: dict(fam[n]) for n in roof_names}
    # The bracket of the eager pass also contains whatever the host did between the two event records -- for the recurrence
    # wrappers that is the allocation of ~0.7 GB of saved state, and a cudaMalloc there shows up as tens of ms of "kernel
    # time" in some runs.  Every recorded launch is therefore re-issued on its own operands three times (the allocator has
    # the blocks cached from the second time on) and the MINIMUM bracket is what enters `roofline`.
    for n in roof_names:
        total = 0.0
        for fn, a, kw in kt.calls.get(n, []):
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out_ = fn(*a, **kw)
                e1.record()
                torch.cuda.synchronize()
                del out_
                t = e0.elapsed_time(e1)
                best = t if best is None else min(best, t)
            total += best
        if kt.calls.get(n):
            roof[n]['ms'] = total
            fam[n] = dict(fam[n], ms=total)
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
    hbm = hbm_kernels(device, args.batch, cfg) if rank == 0 else None
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
            import traceback
            traceback.print_exc()
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
        # roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of the recurrence kernels from ONE `ncu --set full` capture
        # (ncu cannot run inside the bench): the committed round-2 capture at T = 1008 (profiles/r02_blstm_ncu.json), scaled by
        # T to the four pyramid layers and averaged per launch exactly like `achieved` (forward + backward launches).
        traffic, traffic_note = None, 'no ncu capture found under profiles/'
        try:
            with open(os.path.join(ROOT, 'profiles', 'r02_blstm_ncu.json')) as fh:
                cap = json.load(fh)
            t_pad = args.frames + 8 - args.frames % 8
            if cap['blstm_fwd_tc']['T'] == t_pad and cap['blstm_fwd_tc']['B'] == args.batch:
                t_mean = sum(t_pad // 2 ** l for l in range(4)) / 4.0
                traffic = cap['dram_bytes_per_launch_T1008_avg'] * t_mean / t_pad
                traffic_note = ('DRAM bytes per launch averaged over the 8 recurrence launches of a step (4 layers x fwd/bwd): ncu --set full '
                                'capture of the T = %d launches (fwd %.0f MB, bwd %.0f MB; profiles/r02_blstm_ncu.json, r02_ncu_summary.txt) '
                                'scaled by T per layer; the algorithmic bytes of the same launches are 990 / 957 MB, i.e. no wasted re-reads'
                                % (t_pad, (cap['blstm_fwd_tc']['dram_bytes_read'] + cap['blstm_fwd_tc']['dram_bytes_write']) / 1e6,
                                   (cap['blstm_bwd_tc']['dram_bytes_read'] + cap['blstm_bwd_tc']['dram_bytes_write']) / 1e6))
            else:
                traffic_note = 'the committed ncu capture is for another shape'
        except (OSError, KeyError, ValueError):
            pass
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
                         'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_src,
                         'us_per_time_step': 1e3 * r_ms / (2 * sum((args.frames + 8 - args.frames % 8) // 2 ** l for l in range(4))),
                         'launches': r_calls, 'avg_launch_ms': r_ms / max(r_calls, 1),
                         'timed': 'every launch of this kernel family in an eager pass of the same step is re-issued alone on its '
                                  'own operands three times, CUDA events around the call on its stream, minimum taken (a first '
                                  'bracket can contain the allocation of the saved state); the timed region itself replays the '
                                  'step as one CUDA graph',
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
        line['hbm'] = hbm
        if world == 1 and not args.no_cpu_baseline:
            del graphed
            torch.cuda.empty_cache()
            line['stock_torch_cuda'] = stock_torch_cuda(args)
            line['cpu_baseline'] = cpu_baseline_subprocess(args)
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
    ap.add_argument('--ref-budget', type=float, default=150.0, help='--impl reference: seconds of CPU work to spend')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of a CUDA graph replay')
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
