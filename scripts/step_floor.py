#!/usr/bin/env python
"""Launch floor of the fused step on this box ("if a floor is claimed, commit the measurement").

Same launch shape as elg_step_fast_kernel at 4096 envs (148 CTAs x 1024 threads, ~90 KB dynamic shared memory, PDL, CUDA graph):
  empty      griddepcontrol.launch_dependents / wait only
  roundtrip  per CTA one bulk load of the step's algorithmic input bytes and one bulk store of its output bytes, no arithmetic,
             cold L2 (rotating replicas > 2x L2) and warm (one replica)
usage: python scripts/step_floor.py  -> prints and writes gpurun_out/step_floor.json"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from extended_legged_gym_b200 import _lib

dev = "cuda:0"
lib = _lib.load()
SMS = torch.cuda.get_device_properties(0).multi_processor_count


def graph_time(fn, steps=400):
    stream = torch.cuda.Stream(device=dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(3):
            fn(i)
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            for i in range(steps):
                fn(i)
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream); g.replay(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / steps)
    ts.sort()
    return ts[len(ts) // 2]


out = {"sms": SMS}
cur = lambda: torch.cuda.current_stream(dev).cuda_stream
for pdl in (1, 0):
    t = graph_time(lambda i: _lib.check(lib.elg_probe_empty(SMS, 1024, 90 * 1024, pdl, cur())))
    out[f"empty_us_pdl{pdl}"] = t
    print(f"empty kernel, {SMS} x 1024 threads, 90 KB smem, pdl={pdl}: {t:.2f} us per launch", flush=True)
t = graph_time(lambda i: _lib.check(lib.elg_probe_empty(SMS, 128, 0, 1, cur())))
out["empty_us_small_cta"] = t
print(f"empty kernel, {SMS} x 128 threads, no smem, pdl=1: {t:.2f} us per launch", flush=True)

# the step's algorithmic bytes at 4096 anymal_c_rough envs: 688 B read + 2098 B written per env, 28 envs per CTA (rounded to 16 B)
envs_per_cta = 28
b_in = (688 * envs_per_cta + 15) // 16 * 16
b_out = (2098 * envs_per_cta + 15) // 16 * 16
for name, n_rep in (("cold", 24), ("warm", 1)):
    src = [torch.zeros(SMS * b_in, dtype=torch.uint8, device=dev) for _ in range(n_rep)]
    dst = [torch.zeros(SMS * b_out, dtype=torch.uint8, device=dev) for _ in range(n_rep)]
    # cold: pad the rotation beyond 2x L2 with a flush buffer touched by nobody else -- the replicas alone are 24 x 11.7 MB = 280 MB
    t = graph_time(lambda i: _lib.check(lib.elg_probe_roundtrip(src[i % n_rep].data_ptr(), dst[i % n_rep].data_ptr(), b_in, b_out, SMS, 1, cur())))
    out[f"roundtrip_us_{name}"] = t
    tot = SMS * (b_in + b_out)
    print(f"bulk round trip ({name} L2, {n_rep} replicas): {tot / 1e6:.2f} MB per launch, {t:.2f} us per launch = {tot / t / 1e3:.0f} GB/s", flush=True)
    del src, dst
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "step_floor.json"), "w"), indent=1)
