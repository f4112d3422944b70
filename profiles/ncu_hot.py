"""Instruction mix and hot blocks of one kernel instance from an ncu report (source page).
usage: python profiles/ncu_hot.py <rep> <kernel-substring> [n_top]"""
import csv
import subprocess
import sys
from collections import Counter

rep, want = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        sections.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
sec = [s for s in sections if want in s['name']][0]
hdr, data = sec['rows'][0], [r for r in sec['rows'][1:] if len(r) == len(sec['rows'][0])]
print(sec['name'])
ia, isrc, ismp, ithr = (hdr.index(x) for x in ('Instructions Executed', 'Source', '# Samples', 'Avg. Threads Executed'))
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ismp]) for r in data)
print('total warp instr', tot, 'samples', tots, 'static instr', len(data))
c, s = Counter(), Counter()
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[ia]); s[op] += int(r[ismp])
for op, n in c.most_common(20):
    print(f"{op:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*s[op]/tots:5.1f}%")
prev, start = None, 0
for i, r in enumerate(data + [None]):
    n = int(r[ia]) if r is not None else -1
    if prev is None or abs(n - prev) > 0.05 * max(n, prev, 1):
        if prev is not None and (i - start) >= 6 and prev > 0:
            ss = sum(int(x[ismp]) for x in data[start:i]); ii = sum(int(x[ia]) for x in data[start:i])
            print(f"block {start:5d}-{i:5d} len {i-start:4d} exec/instr {prev:10d} instr {100*ii/tot:5.1f}% samples {100*ss/tots:5.1f}% thr {data[start][ithr]}")
        start, prev = i, n
top = sorted(range(len(data)), key=lambda i: -int(data[i][ismp]))[:ntop]
for i in sorted(top):
    print(i, data[i][ismp], data[i][ia], data[i][isrc][:100])
if '--dump' in sys.argv:
    a, b = int(sys.argv[sys.argv.index('--dump') + 1]), int(sys.argv[sys.argv.index('--dump') + 2])
    for i in range(a, b):
        print(i, data[i][ismp].rjust(6), data[i][ia].rjust(9), data[i][isrc][:90])
