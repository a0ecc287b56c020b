"""Where does the graphed cfg-3 training step spend its time?  Each section (acoustic encoder, LAS decoder,
mix + Transformer + loss) is captured into its own CUDA graph (fwd + bwd) and replayed; so is the whole step.
usage: python scripts/section_times.py [dtype]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from b200st import runtime, functional as BF
from oracle import st_oracle as O
from b200st.train_step import Trainer_ST
from models.Enc import padded_lengths

runtime.set_compute_dtype(sys.argv[1] if len(sys.argv) > 1 else 'bf16')
cfg = bench.st_config()
dev = torch.device('cuda')
model = bench.build_model(cfg, dev)
host = O.synthetic_batch(cfg, 64, 1000, seed=333)
src, tgt, feats = host['src'].to(dev), host['tgt'].to(dev), host['acous_feats'].to(dev)
lens_dev = torch.as_tensor([int(n) for n in host['acous_lens']], dtype=torch.int32, device=dev)
items = {'srcid': [src], 'tgtid': [tgt], 'acous_feat': [feats], 'acouslen': lens_dev}
tr = Trainer_ST(use_gpu=True, batch_size=64)


def timed_graph(name, fn, reps=10):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            model.zero_grad(set_to_none=True); fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    model.zero_grad(set_to_none=True); runtime.clear_cache()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
        runtime.join_deferred()        # weight-gradient GEMMs forked during backward
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:34s} {e0.elapsed_time(e1) / reps:8.3f} ms', flush=True)


# ---- section inputs (detached leaves)
with torch.no_grad():
    ln, _ = padded_lengths(lens_dev, feats.size(0), feats.size(1), dev)
    acous_out = model.las.encoder(feats, acous_lens=lens_dev, is_training=False, lens_dev=ln).detach()
    klens = ln // 8
    emb_dyn, _, _, lengths = model.las.decoder.forward_device(acous_out, klens, need_logps=False)
    emb_dyn = emb_dyn.detach()


def sec_enc():
    y = model.las.encoder(feats, acous_lens=lens_dev, is_training=False, lens_dev=ln)
    y.backward(torch.ones_like(y))


def sec_enc_fwd():
    with torch.no_grad():
        model.las.encoder(feats, acous_lens=lens_dev, is_training=False, lens_dev=ln)


def sec_dec():
    a = acous_out.clone().requires_grad_(True)
    e, _, _, _ = model.las.decoder.forward_device(a, klens, need_logps=False)
    e.backward(torch.ones_like(e))


def sec_dec_fwd():
    with torch.no_grad():
        model.las.decoder.forward_device(acous_out, klens, need_logps=False)


def sec_tf():
    d = emb_dyn.clone().requires_grad_(True)
    tgt_mask, emb_tgt = model._get_tgt_emb(tgt, dev)
    src_trim = model._pre_proc_src(src, dev)
    _, emb_src, _ = model._get_src_emb(src_trim, d, dev)
    mask = model._length_mask(lengths, emb_src.size(1))
    enc = model._encoder_en(emb_src, src_mask=mask)
    _, _, logps, _, _ = model._decoder_de(emb_tgt, enc, tgt_mask=tgt_mask, src_mask=mask)
    (logps.float().sum() * 1e-6).backward()


timed_graph('whole step', lambda: tr._train_batch_device(model, items))
timed_graph('acoustic encoder fwd+bwd', sec_enc)
timed_graph('acoustic encoder fwd', sec_enc_fwd)
timed_graph('LAS decoder fwd+bwd', sec_dec)
timed_graph('LAS decoder fwd', sec_dec_fwd)
timed_graph('mix + TF enc/dec + out fwd+bwd', sec_tf)
