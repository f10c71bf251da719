#!/usr/bin/env python
"""In-kernel timeline of the fused step kernel (clock64 stamps of CTA 0) + an empty-kernel launch floor."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from step_sweep import make, lib, dev, _lib
names = ["entry", "tables done", "after sync0", "loads landed", "items done", "after B1", "scalar done", "rows done (w0)", "after B2",
         "head done", "stores issued", "stores drained"]
for n in (4096, 1024):
    env = make("anymal_c_rough", n, 0)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.elg_set_step_debug.argtypes = [C.c_void_p]
    lib.elg_set_step_debug(buf.data_ptr())
    p = env._params
    p.noise_mode, p.clip_observations = _lib.NOISE_PHILOX, 100.0
    for i in range(5):
        p.noise_offset = i
        _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), _lib.PHASE_FUSED, None))
        torch.cuda.synchronize()
    st = buf.cpu().tolist()
    print(f"N={n}: cycles since entry (x0.509 ns at 1965 MHz)")
    for i, nm in enumerate(names):
        print(f"  {nm:18s} {st[i] - st[0]:8d} cyc  {(st[i] - st[0]) / 1965.0:7.2f} us")
    lib.elg_set_step_debug(None)
# launch floor: a trivial torch kernel chain in a graph
x = torch.zeros(148 * 1024, device=dev)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    x.add_(1.0); s.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(400):
            x.add_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.cuda.synchronize(); e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    print("tiny elementwise kernel in a graph: %.2f us per launch" % (e0.elapsed_time(e1) * 1e3 / 400))
