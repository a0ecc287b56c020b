"""SASS evidence that the hot kernels use the Blackwell tensor-core / TMA path: per kernel family of libb200st.so, the counts of
UTCHMMA (tcgen05.mma), UTMALDG (TMA load), LDTM / STTM (tcgen05.ld / st of TMEM), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
plus a few raw lines of one member.   python scripts/sass_summary.py > profiles/r02_sass_tcgen05.txt   (no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'speech-translation-joint-embedding-passing_b200', 'b200st', 'libb200st.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCCP|UTCATOMSWS|SYNCS|HMMA)\b[\.\w]*')
fam = collections.OrderedDict()
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    mangled, body = f.split('\n', 1)
    name = subprocess.run(['c++filt', mangled.strip()], capture_output=True, text=True).stdout.strip()
    base = re.sub(r'^void ', '', name).split('(')[0].split('<')[0]
    c = collections.Counter(m.group(1) for m in pat.finditer(body))
    e = fam.setdefault(base, {'n': 0, 'c': collections.Counter(), 'lines': []})
    e['n'] += 1
    for k, v in c.items():
        e['c'][k] = max(e['c'][k], v)
    if not e['lines']:
        e['lines'] = [ln.strip() for ln in body.split('\n') if re.search(r'UTCHMMA|UTMALDG|LDTM|UTCBAR', ln)][:6]
print(f'# cuobjdump -sass {os.path.relpath(so, ROOT)}: {sum(e["n"] for e in fam.values())} kernels in {len(fam)} families')
print('# max count per instantiation of: UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA), LDTM/STTM = tcgen05.ld/st (TMEM),')
print('# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected)')
for base, e in fam.items():
    if e['c'].get('UTCHMMA', 0) + e['c'].get('UTMALDG', 0) + e['c'].get('LDTM', 0) == 0:
        continue
    print(f'\n{base}  ({e["n"]} instantiations)\n    ' + ', '.join(f'{k} {v}' for k, v in sorted(e['c'].items())))
    for ln in e['lines']:
        print('      ' + re.sub(r'\s+', ' ', ln)[:150])
others = [b for b, e in fam.items() if e['c'].get('UTCHMMA', 0) + e['c'].get('UTMALDG', 0) + e['c'].get('LDTM', 0) == 0]
print(f'\n# {len(others)} CUDA-core families (no tensor-core / TMA instructions): ' + ', '.join(sorted(others)))
