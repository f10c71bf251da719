#!/usr/bin/env python
"""A/B timing of the fused step: lean kernel (elg_step_fast.cu) vs the generic kernel, per env count.
usage: python scripts/step_ab.py [quick]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from step_sweep import time_cfg, lib, _lib
quick = len(sys.argv) > 1
rows = []
cfgs = ((4096, 22), (4096, 1)) if quick else ((4096, 22), (4096, 1), (1024, 1), (16384, 6), (65536, 3))
for n, rep in cfgs:
    for name, flag in (("fast", 0), ("generic", 2)):
        lib.elg_set_step_tuning(0, 0, 0, flag)
        t = time_cfg("anymal_c_rough", n, rep, _lib.PHASE_FUSED, steps=200 if n > 8192 else 400)
        rows.append((name, n, rep, t))
        print(f"{name:8s} N={n:6d} replicas={rep:2d}  {t:8.2f} us/launch  {n * 2786 / t / 1e3:8.1f} GB/s algorithmic", flush=True)
lib.elg_set_step_tuning(0, 0, 0, 0)
for case in ("anymal_c_flat", "a1_rough", "go2_rough"):
    t = time_cfg(case, 4096, 22, _lib.PHASE_FUSED)
    print(f"fast     {case} N=4096  {t:8.2f} us/launch", flush=True)
if not quick:
    for cap in (12, 16, 20, 24, 28):
        lib.elg_set_step_tuning(cap, 0, 1, 0)
        t = time_cfg("anymal_c_rough", 65536, 3, _lib.PHASE_FUSED, steps=100)
        print(f"fast cap={cap:2d} N=65536  {t:8.2f} us/launch  {65536 * 2786 / t / 1e3:8.1f} GB/s", flush=True)
    lib.elg_set_step_tuning(0, 0, 0, 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "step_ab.json"), "w"))
