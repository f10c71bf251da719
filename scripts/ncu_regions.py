#!/usr/bin/env python
"""Warp-instruction / sample totals of an ncu cuda,sass source export, grouped by line ranges of one file.
usage: ncu_regions.py export.csv file.cu name:lo-hi [name:lo-hi ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
fname = sys.argv[2]
regions = []
for a in sys.argv[3:]:
    n, r = a.split(':'); lo, hi = r.split('-'); regions.append((n, int(lo), int(hi)))
def num(s):
    try: return int(s)
    except Exception: return 0
cur = None; agg = {r[0]: [0, 0] for r in regions}; other = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0].isdigit() and len(r) >= 8 and r[2] == '-':
        ln, s, i = int(r[0]), num(r[4]), num(r[7])
        hit = False
        if cur == fname:
            for n, a, b in regions:
                if a <= ln <= b: agg[n][0] += s; agg[n][1] += i; hit = True; break
        if not hit:
            o = other.setdefault(cur if cur != fname else f"{cur}:{ln}", [0, 0]); o[0] += s; o[1] += i
for k, v in agg.items(): print(f"{k:14s} samples={v[0]:5d} warp-inst={v[1]:9d}")
for k, v in sorted(other.items(), key=lambda kv: -kv[1][1])[:12]: print(f"  {k:28s} samples={v[0]:5d} warp-inst={v[1]:9d}")
