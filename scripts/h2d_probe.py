#!/usr/bin/env python
"""Host->device copy rate probe: torch pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), sizes of the
e2e loop's inputs.  Not a benchmark."""
import ctypes, sys, time
import numpy as np
import torch

dev = "cuda:0"
torch.cuda.init()
libcudart = ctypes.CDLL("libcudart.so")


def wc_tensor(nbytes):
    p = ctypes.c_void_p()
    rc = libcudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def rate(host, devt, reps=50):
    s = torch.cuda.current_stream()
    for _ in range(5):
        devt.copy_(host, non_blocking=True)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        devt.copy_(host, non_blocking=True)
    e1.record()
    s.synchronize()
    return host.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


for nbytes in (200 << 10, 836 << 10, 3620 << 10, 5260 << 10, 64 << 20):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    a.fill_(1)
    try:
        w = wc_tensor(nbytes)
        w.fill_(1)
        rw = rate(w, d)
    except Exception as ex:  # noqa: BLE001
        rw = repr(ex)
    back = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    s = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        back.copy_(d, non_blocking=True)
    e1.record(); s.synchronize()
    print(f"{nbytes >> 10:7d} KiB: H2D pinned {rate(a, d):6.1f} GB/s   H2D write-combined {rw if isinstance(rw, str) else round(rw, 1)} GB/s   "
          f"D2H pinned {nbytes * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s", flush=True)
