#!/usr/bin/env python
"""Small driver for profiling the non-step kernels under ncu: a few launches each of the fused depth camera (BASELINE
config 4), the API-level ray cast, the SDF query and the main -> rollout clone (config 5).  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench

if __name__ == "__main__":
    torch.cuda.set_device(0)
    out = bench.secondary_benchmarks("cuda:0", 1, 0, None, quick=True)
    from extended_legged_gym_b200 import synthetic
    from extended_legged_gym_b200.utils.mesh_sdf import MeshSDF, MeshSDFCfg
    hf = synthetic.make_height_field(seed=0)
    v, t = synthetic.heightfield_to_trimesh(hf)
    m = MeshSDF(MeshSDFCfg(vertices=torch.from_numpy(v), triangles=torch.from_numpy(t), max_distance=2.0), "cuda:0")
    g = torch.Generator().manual_seed(0)
    p = torch.stack([torch.rand(164160, generator=g) * 36 + 2, torch.rand(164160, generator=g) * 36 + 2, torch.rand(164160, generator=g) * 0.8], 1).cuda()
    for _ in range(3):
        m.query(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.query(p); e1.record(); torch.cuda.synchronize()
    out["sdf_query"] = {"points": 164160, "ms": e0.elapsed_time(e1), "Mpoints/s": 164160 / e0.elapsed_time(e1) / 1e3}
    import json
    print(json.dumps(out))
