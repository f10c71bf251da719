#!/usr/bin/env python
"""Depth-camera / ray-cast throughput only (the `secondary.depth_raycast` entry of bench.py); not a benchmark line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

torch.cuda.set_device(0)
out = bench.secondary_benchmarks("cuda:0", 1, 0, None, quick=True, only_depth=True)
print(json.dumps(out["depth_raycast"]))
