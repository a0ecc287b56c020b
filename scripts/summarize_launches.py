"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel for ONE training step.
usage: python scripts/summarize_launches.py gpurun_out/launches_r1.csv [n_steps_in_file]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = rows[0], rows[1:]
iN, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
# the measured step = everything after the 4th-from-last... : steps are delimited by the blstm_fwd kernels (4 per step)
idx_fwd = [i for i, r in enumerate(data) if 'blstm_fwd' in r[iN]]
n_layers = 4
start = idx_fwd[-n_layers] if len(idx_fwd) >= n_layers else 0
# walk back to the start of that step: the first kernels of a step precede the first blstm_fwd by the embedding/mask/
# transpose/projection launches; use the end of the previous step's last blstm_bwd + its trailing GEMMs as the boundary
idx_bwd = [i for i, r in enumerate(data) if 'blstm_bwd' in r[iN] and i < start]
if idx_bwd:
    j = idx_bwd[-1] + 1
    while j < start and ('gemm' in data[j][iN] or 'colsum' in data[j][iN] or 'Memset' in data[j][iN]):
        j += 1
    start = j
step = data[start:]
agg = collections.defaultdict(lambda: [0, 0.0])
def us(r):
    v = float(r[iV].replace(',', '')); u = r[iU]
    return v / 1000 if u.startswith('n') else (v if u.startswith('u') else v * 1000)
for r in step:
    name = re.sub(r'\(.*', '', r[iN])
    name = name[:90] if 'gemm' in name else re.sub(r'<.*', '', name)
    name = name.replace('void ', '').replace('b200st::', '')
    agg[name][0] += 1; agg[name][1] += us(r)
tot = sum(v[1] for v in agg.values())
print(f'one training step: {len(step)} kernel launches, {tot/1000:.2f} ms summed kernel time (ncu: serialised, cold cache)')
print(f'{"us":>10s} {"share":>6s} {"n":>5s} {"avg us":>9s}  kernel')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:10.1f} {100*t/tot:5.1f}% {n:5d} {t/n:9.2f}  {k}')
