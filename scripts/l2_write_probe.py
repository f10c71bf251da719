#!/usr/bin/env python
"""Reference points for the clone kernel: how fast does the chip absorb ~12.7 MB of writes per launch (CUDA graph of 50 launches)?"""
import torch
dev = "cuda:0"
def timed(fn, n=50, reps=4):
    gs = torch.cuda.Stream(device=dev); gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gs):
        fn(); gs.synchronize()
        with torch.cuda.graph(gr, stream=gs):
            for _ in range(n): fn()
        gr.replay(); gs.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(gs)
        for _ in range(reps): gr.replay()
        e1.record(gs); gs.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)
for mb in (1.0, 4.0, 12.7, 50.0, 200.0):
    n = int(mb * 1e6 / 4)
    x = torch.randn(n, device=dev); y = torch.empty_like(x)
    t_fill = timed(lambda: y.fill_(1.0))
    t_copy = timed(lambda: y.copy_(x))
    row = torch.randn(97, device=dev)
    z = torch.empty(n // 97, 97, device=dev)
    t_bcast = timed(lambda: z.copy_(row.expand_as(z)))
    print(f"{mb:6.1f} MB: fill {t_fill:7.2f} us ({mb * 1e3 / t_fill:7.1f} GB/s)   copy {t_copy:7.2f} us ({mb * 1e3 / t_copy:7.1f} GB/s written)   broadcast-row copy {t_bcast:7.2f} us", flush=True)
