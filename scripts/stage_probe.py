#!/usr/bin/env python
"""Host<->device block transfer probe at the e2e step's sizes: one copy-engine cudaMemcpyAsync against elg_stage_block (SM-issued
loads / TMA bulk pieces from mapped pinned memory), alone and with the opposite direction running on a second stream.
Not a benchmark.  usage: python scripts/stage_probe.py [out.json]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from extended_legged_gym_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda:0"
torch.cuda.init()
IN, OUT = 5259264, 3870720
REPS = 40


def timed(fn, stream, reps=REPS, other=None, other_stream=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
        if other is not None:
            other()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps   # us


res = {}
host_in = torch.empty(IN, dtype=torch.uint8).pin_memory()
host_in.fill_(3)
dev_in = torch.empty(IN, dtype=torch.uint8, device=dev)
dev_out = torch.ones(OUT, dtype=torch.uint8, device=dev)
host_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
s0, s1 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)


def ce_in():
    with torch.cuda.stream(s0):
        dev_in.copy_(host_in, non_blocking=True)


def ce_out():
    with torch.cuda.stream(s1):
        host_out.copy_(dev_out, non_blocking=True)


def k_in(mode, grid):
    def f():
        _lib.check(lib.elg_stage_block(dev_in.data_ptr(), host_in.data_ptr(), IN, mode, grid, s0.cuda_stream), "elg_stage_block")
    return f


def k_out(mode, grid):
    def f():
        _lib.check(lib.elg_stage_block(host_out.data_ptr(), dev_out.data_ptr(), OUT, mode, grid, s1.cuda_stream), "elg_stage_block")
    return f


def gbs(nbytes, us):
    return round(nbytes / us * 1e-3, 1)


t = timed(ce_in, s0)
res["h2d_copy_engine_alone"] = {"us": round(t, 1), "GBs": gbs(IN, t)}
t = timed(ce_out, s1)
res["d2h_copy_engine_alone"] = {"us": round(t, 1), "GBs": gbs(OUT, t)}
t = timed(ce_in, s0, other=ce_out)
res["h2d_copy_engine_with_d2h_copy_engine"] = {"us": round(t, 1), "GBs": gbs(IN, t)}
for mode, grids in ((0, (74, 148, 296, 592, 1184)), (1, (16, 37, 74, 148))):
    for g in grids:
        try:
            t = timed(k_in(mode, g), s0)
            t2 = timed(k_in(mode, g), s0, other=ce_out)
            res[f"h2d_kernel_mode{mode}_grid{g}"] = {"us": round(t, 1), "GBs": gbs(IN, t), "us_with_d2h_copy_engine": round(t2, 1)}
        except Exception as ex:  # noqa: BLE001
            res[f"h2d_kernel_mode{mode}_grid{g}"] = repr(ex)
        print(f"h2d mode {mode} grid {g}: {res[f'h2d_kernel_mode{mode}_grid{g}']}", flush=True)
# correctness of the kernel copy
dev_in.zero_()
k_in(0, 0)()
torch.cuda.synchronize()
res["h2d_kernel_mode0_correct"] = bool((dev_in == 3).all())
dev_in.zero_()
k_in(1, 0)()
torch.cuda.synchronize()
res["h2d_kernel_mode1_correct"] = bool((dev_in == 3).all())
for mode, g in ((0, 148), (0, 592), (1, 74)):
    t = timed(k_out(mode, g), s1)
    res[f"d2h_kernel_mode{mode}_grid{g}"] = {"us": round(t, 1), "GBs": gbs(OUT, t)}
host_out.zero_()
k_out(0, 0)()
torch.cuda.synchronize()
res["d2h_kernel_correct"] = bool((host_out == 1).all())
# both directions by kernels
t = timed(k_in(0, 296), s0, other=k_out(0, 148))
res["h2d_kernel_mode0_grid296_with_d2h_kernel"] = {"us": round(t, 1)}
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
