#!/usr/bin/env python
"""PDL on / off (flag 16) for CTA shapes that can / cannot share an SM with the next launch's CTAs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from step_sweep import time_cfg, lib, _lib
for n, cap in ((4096, 0), (1776, 12), (148 * 8, 8), (148 * 16, 16), (148 * 20, 20)):
    for f in (0, 16):
        lib.elg_set_step_tuning(cap, f, 1, 0) if cap else lib.elg_set_step_tuning(0, f, 0, 0)
        for rep in (22, 1):
            t = time_cfg("anymal_c_rough", n, rep if n > 2000 else rep * 2, _lib.PHASE_FUSED, steps=400)
            print(f"N={n:5d} cap={cap:2d} flags={f:2d} replicas={rep:2d}  {t:7.2f} us/launch  {t / n * 1e3:7.3f} ns/env", flush=True)
lib.elg_set_step_tuning(0, 0, 0, 0)
