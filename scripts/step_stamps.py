#!/usr/bin/env python
"""In-kernel timeline of the lean fused step kernel (clock64 stamps of CTA 0) + an empty-kernel launch floor."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from extended_legged_gym_b200 import build as _build
# the product library carries no stamp code: build the diagnostic variant and load that one
os.environ["ELG_LIB_PATH"] = _build.build(defines=("ELG_STEP_STAMPS",), out=os.path.join(_build.PKG_DIR, "libelg_b200_stamps.so"))
from step_sweep import make, lib, dev, _lib
names = {0: "entry", 1: "after griddepcontrol.wait", 2: "root_states landed (warp 0)", 3: "phase A done (warp 0)", 4: "B1 passed", 5: "row0: gathers issued (before phase A)",
         6: "row0: scan done", 7: "row0: head done", 8: "assembly warp done", 9: "B2 passed", 10: "stores read out (warp 0)", 11: "warp 0 reaches the TMA wait", 12: "loads landed (last warp, idle until then)"}
for case, n in (("anymal_c_rough", 4096), ("anymal_c_flat", 4096)):
    env = make(case, n, 0)
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    lib.elg_set_step_debug.argtypes = [C.c_void_p]
    lib.elg_set_step_debug(buf.data_ptr())
    p = env._params
    p.noise_mode, p.clip_observations = _lib.NOISE_PHILOX, 100.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(5):
        p.noise_offset = i
        flush.fill_(i)
        torch.cuda.synchronize()
        _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), _lib.PHASE_FUSED, None))
        torch.cuda.synchronize()
    st = buf.cpu().tolist()
    print(f"{case} N={n}: cycles since entry, cold L2 (1965 MHz)")
    for i in sorted(names, key=lambda k: st[k]):
        if st[i]:
            print(f"  {names[i]:32s} {st[i] - st[0]:8d} cyc  {(st[i] - st[0]) / 1965.0:7.2f} us")
    print('  phase A end per warp (us since entry):', ' '.join(f'{(st[32 + w] - st[0]) / 1965.0:.2f}' for w in range(32)))
    lib.elg_set_step_debug(None)
    del flush
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
