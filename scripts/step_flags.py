#!/usr/bin/env python
"""Experiment switches of the lean step kernel (FastPlan.flags), timed side by side. usage: step_flags.py f1 f2 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from step_sweep import time_cfg, lib, _lib
flags = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 6, 7]
for rep in (22, 1):
    for f in flags:
        lib.elg_set_step_tuning(0, f, 0, 0)
        t = time_cfg("anymal_c_rough", 4096, rep, _lib.PHASE_FUSED, steps=400)
        print(f"flags={f:2d} N=4096 replicas={rep:2d}  {t:7.2f} us/launch", flush=True)
lib.elg_set_step_tuning(0, 0, 0, 0)
