#!/usr/bin/env python
"""Many-chunk launches of the lean step kernel: register budget for 1 / 3 / 4 CTAs per SM, 8- and 12-env CTAs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from step_sweep import time_cfg, lib, _lib
for n, rep in ((65536, 3), (32768, 5), (8192, 12)):
    for cap, ctas in ((0, 1), (0, 3), (0, 4), (8, 3), (8, 4), (28, 1)):
        lib.elg_set_step_tuning(cap, 0, ctas, 0)
        t = time_cfg("anymal_c_rough", n, rep, _lib.PHASE_FUSED, steps=100)
        print(f"N={n:6d} cap={cap:2d} ctas/SM={ctas}  {t:8.2f} us/launch  {n * 2786 / t / 1e3:8.1f} GB/s  ({n * 2786 / t / 1e3 / 6650 * 100:4.1f} % of 6650)", flush=True)
lib.elg_set_step_tuning(0, 0, 0, 0)
