#!/usr/bin/env python
"""Compact per-instruction view of an `ncu --page source --csv --print-source sass` export (first kernel instance):
runs of consecutive SASS instructions with the same executed count are folded into blocks.
usage: ncu_sass_blocks.py export.csv [min_exec] [--full]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
full = "--full" in sys.argv
inst = []
seen = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        if seen > 1: break
        continue
    if not r or r[0] == "Address": continue
    try: inst.append((r[1].strip(), int(r[5]), int(r[4]), int(r[6])))
    except Exception: pass
tot = sum(i[1] for i in inst); samp = sum(i[2] for i in inst)
print(f"{len(inst)} SASS instructions, {tot} warp-inst executed, {samp} samples")
if full:
    for k, (s, e, sm, th) in enumerate(inst):
        print(f"{k:5d} {e:7d} {sm:4d} {th/max(e,1):5.1f} {s}")
    sys.exit()
blocks = []
for k, (s, e, sm, th) in enumerate(inst):
    if blocks and blocks[-1][2] == e: b = blocks[-1]; b[1] = k; b[3] += e; b[4] += sm
    else: blocks.append([k, k, e, e, sm])
for b in blocks:
    if b[3] >= min_exec:
        print(f"[{b[0]:5d}-{b[1]:5d}] n={b[1]-b[0]+1:4d} exec/inst={b[2]:7d} warp-inst={b[3]:8d} ({100*b[3]/tot:4.1f}%) samples={b[4]:4d}  {inst[b[0]][0][:60]}")
