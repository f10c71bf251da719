#!/usr/bin/env python
"""Timing of the observation normaliser pair (graph of 48 calls over 24 storage slots); not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from extended_legged_gym_b200.utils.normalizer import EmpiricalNormalization

from extended_legged_gym_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
# modes of elg_set_normalizer_tuning: 0 default (column-parallel single launch up to 8192 rows), 1 two launches, 2 row-parallel single
# launch with a grid-wide hand-over
MODES = (0, 1, 2, 4, 8, 12, 16)      # (4 / 8 / 12 / 16: form 0 with the cluster size forced to 1 / 2 / 4 / 8)
for n, o in ((4096, 235), (4096, 48), (65536, 48), (32832, 235)):
    for training, mode in [(True, m) for m in MODES] + [(False, 0)]:
        lib.elg_set_normalizer_tuning(mode)
        norm = EmpiricalNormalization(shape=[o], until=int(1e12)).to(dev)
        norm.train(training)
        x = torch.randn(n, o, device=dev)
        slots = max(2, min(24, int(4e8 // (n * o * 4))))
        slot = torch.empty(slots, n, o, device=dev)
        gs = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(gs):
            norm.forward_into(x, slot[0])
            gs.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=gs):
                for i in range(48):
                    norm.forward_into(x, slot[i % slots])
            g.replay(); gs.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(gs)
            for _ in range(4):
                g.replay()
            e1.record(gs); gs.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 192
        lib.elg_set_normalizer_tuning(0)
        print(f"{n} x {o} training={training} mode={mode}: {us:.2f} us/call, {n * o * 8 / us / 1e3:.0f} GB/s", flush=True)
