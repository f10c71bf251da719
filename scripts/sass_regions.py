#!/usr/bin/env python
"""Static SASS instruction counts per source-line region of one kernel.
usage: sass_regions.py <nvdisasm --print-line-info output> <first line no> <last line no> file.cu name:lo-hi ..."""
import re, sys, collections
lines = open(sys.argv[1]).read().split('\n')[int(sys.argv[2]):int(sys.argv[3])]
fname = sys.argv[4]
reg = []
for a in sys.argv[5:]:
    n, r = a.split(':'); lo, hi = r.split('-'); reg.append((n, int(lo), int(hi)))
cur = None; cnt = collections.Counter()
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', l) and cur: cnt[cur] += 1
print("total", sum(cnt.values()))
agg = collections.Counter()
for (f, ln), c in cnt.items():
    key = f
    if f == fname:
        key = 'other:%d' % ln
        for n, a, b in reg:
            if a <= ln <= b: key = n; break
    agg[key] += c
for k, v in agg.most_common(24): print(f"  {k:26s} {v}")
