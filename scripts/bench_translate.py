"""BASELINE.json configs[4]: translate.py ST inference on the joint model -- greedy and beam-5 decode, synthetic
1000-frame utterances, batch 128, utterances/s on one B200 (Seq2seq.forward_translate, Seq2seq.py:641-796).
    python scripts/bench_translate.py [--batch 128] [--frames 1000] [--max-len 50] [--dtype bf16] [--reps 3]
Prints one JSON line per beam width.  Random-init weights never emit EOS reliably, so every utterance decodes the full
--max-len tokens (the worst case for the decoder loop)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import bench
from b200st import runtime
from oracle import st_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=128)
ap.add_argument('--frames', type=int, default=1000)
ap.add_argument('--max-len', type=int, default=50)
ap.add_argument('--dtype', default='bf16')
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--beams', default='1,5')
ap.add_argument('--no-cache', action='store_true', help="the reference's recompute-every-step decoder loop")
ap.add_argument('--las-max-len', type=int, default=0,
                help='LAS decoder steps: 0 = the model default (max_seq_len_src = 32 -> 31 steps), 150 = what the reference CLI '
                     'sets before translating (translate.py:70-77: model.las.decoder.max_seq_len = 150)')
args = ap.parse_args()

dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
runtime.set_compute_dtype(args.dtype)
cfg = bench.st_config()
model = bench.build_model(cfg, dev).eval()
if args.no_cache and hasattr(model, 'decode_cache'):
    model.decode_cache = False
if args.las_max_len:
    model.las.decoder.max_seq_len = args.las_max_len          # translate.py:73
    model.enc_src.expand_time(args.las_max_len)                # translate.py:75-77 (sinusoid tables long enough)
    model.dec_tgt.expand_time(max(args.las_max_len, args.max_len))
data = O.synthetic_batch(cfg, args.batch, args.frames, seed=5)
feats = data['acous_feats'].to(dev)
lens = data['acous_lens']
for k in (int(b) for b in args.beams.split(',')):
    def run():
        return model.forward_translate(acous_feats=feats, acous_lens=lens, beam_width=k, penalty_factor=1,
                                       use_gpu=True, max_seq_len=args.max_len, mode='ST')
    for _ in range(2):            # first call captures the graphs
        out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(json.dumps({'metric': 'st_translate_utt_per_s', 'beam_width': k, 'value': args.batch / (ms / 1e3), 'unit': 'utt/s',
                      'ms_per_batch': ms, 'batch': args.batch, 'frames': args.frames, 'max_seq_len': args.max_len,
                      'las_decoder_steps': int(model.las.decoder.max_seq_len) - 1, 'dtype': args.dtype, 'decoder': 'recompute' if args.no_cache else 'kv-cache',
                      'out_shape': list(out.shape), 'checksum': int(out.sum())}), flush=True)
