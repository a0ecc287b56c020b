"""Aggregate the ncu launch list of scripts/one_step_graph.py by kernel for the ONE measured graph replay (everything after the
stamp_kernel launch that precedes it).  usage: python scripts/summarize_graph_launches.py gpurun_out/r02_graph_launches.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = rows[0], rows[1:]
iN, iV, iU, iG = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
marks = [i for i, r in enumerate(data) if 'stamp_kernel' in r[iN]]
step = data[marks[-1] + 1:] if marks else data
def us(r):
    v = float(r[iV].replace(',', '')); u = r[iU]
    return v / 1000 if u.startswith('n') else (v if u.startswith('u') else v * 1000)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    name = re.sub(r'\(.*', '', r[iN])
    name = name[:90] if 'gemm' in name else re.sub(r'<.*', '', name)
    name = name.replace('void ', '').replace('b200st::', '')
    agg[name][0] += 1; agg[name][1] += us(r)
tot = sum(v[1] for v in agg.values())
ours = sum(v[0] for k, v in agg.items() if not k.startswith('at::') and 'nccl' not in k)
print(f'one replay of the whole-step graph: {len(step)} kernel nodes ({ours} from libb200st.so), {tot/1000:.2f} ms summed kernel time '
      f'(ncu: serialised, cold cache; the replay itself takes ~10.4 ms because branches overlap)')
print(f'{"us":>10s} {"share":>6s} {"n":>5s} {"avg us":>9s}  kernel')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{t:10.1f} {100*t/tot:5.1f}% {n:5d} {t/n:9.2f}  {k}')
