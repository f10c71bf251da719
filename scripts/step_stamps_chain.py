#!/usr/bin/env python
"""In-kernel timeline of the lean step kernel INSIDE the bench's regime: a CUDA graph of launches rotating over 22 state replicas
(cold data, warm code), clock64 stamps of CTA 0 of the last launch of the graph.  The isolated-launch timeline of
scripts/step_stamps.py runs after a full L2 flush, which also evicts the kernel's code -- it overstates every single-warp phase.
usage: python scripts/step_stamps_chain.py [v5]   (v5: stamps of the round-1 library scripts/ab/libelg_v5.so)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from extended_legged_gym_b200 import build as _build
if len(sys.argv) > 1 and sys.argv[1] == "v5":
    os.environ["ELG_LIB_PATH"] = os.path.join(ROOT, "scripts", "ab", "libelg_v5.so")
else:
    os.environ["ELG_LIB_PATH"] = _build.build(defines=("ELG_STEP_STAMPS",), out=os.path.join(_build.PKG_DIR, "libelg_b200_stamps.so"))
from step_sweep import make, lib, dev, _lib
names = {0: "entry", 1: "after griddepcontrol.wait", 2: "root_states landed (warp 0) [v5: all loads]", 3: "phase A done (warp 0)", 4: "B1 passed",
         5: "row0: gathers issued", 6: "row0: scan done", 7: "row0: head done", 8: "assembly done", 9: "B2 passed",
         10: "stores read out (warp 0)", 11: "warp 0 reaches the TMA wait", 12: "all loads landed (last warp)"}
n_rep = 22
for case, n in (("anymal_c_rough", 4096),):
    envs = [make(case, n, r) for r in range(n_rep)]
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    lib.elg_set_step_debug.argtypes = [C.c_void_p]
    lib.elg_set_step_debug(buf.data_ptr())
    stream = torch.cuda.Stream(device=dev)

    def enqueue(env, i):
        p = env._params
        p.noise_mode, p.noise_offset, p.clip_observations = _lib.NOISE_PHILOX, i, 100.0
        _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), _lib.PHASE_FUSED, torch.cuda.current_stream(dev).cuda_stream))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(3):
            enqueue(envs[i % n_rep], i)
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            for i in range(3 * n_rep):
                enqueue(envs[i % n_rep], i)
        for _ in range(4):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream); g.replay(); e1.record(stream)
        torch.cuda.synchronize()
    st = buf.cpu().tolist()
    print(f"{case} N={n}: graph of {3 * n_rep} launches over {n_rep} replicas: {e0.elapsed_time(e1) * 1e3 / (3 * n_rep):.2f} us per launch; CTA 0 of the last launch:")
    for i in sorted(names, key=lambda k: st[k]):
        if st[i]:
            print(f"  {names[i]:46s} {st[i] - st[0]:8d} cyc  {(st[i] - st[0]) / 1965.0:7.2f} us")
    print('  phase A end per warp (us since entry):', ' '.join(f'{(st[32 + w] - st[0]) / 1965.0:.2f}' for w in range(32)))
    lib.elg_set_step_debug(None)
