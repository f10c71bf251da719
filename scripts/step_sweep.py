#!/usr/bin/env python
"""Diagnostic sweep of the fused step kernel: time per launch vs env count, phase mask, replica count (L2 warm/cold).
usage: python scripts/step_sweep.py  -> prints one line per configuration"""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import common
from extended_legged_gym_b200 import _lib, synthetic
from extended_legged_gym_b200.envs import LeggedRobot
from extended_legged_gym_b200.sim_backend import SyntheticSim

dev = "cuda:0"
lib = _lib.load()
hf = synthetic.make_height_field(seed=0).to(dev)


def make(case, n, seed):
    cfg, spec, st = common.make_case_state(case, n, seed=seed)
    cfg.env.num_envs = n
    env = LeggedRobot(cfg, None, SyntheticSim(cfg, n, dev, spec=spec, height_samples=hf, state=st), dev, True)
    env.set_env_state(st)
    env.noise_u = None
    env._sync_native()
    return env


def time_cfg(case, n, n_rep, phase, steps=400, noise=_lib.NOISE_PHILOX):
    envs = [make(case, n, r) for r in range(n_rep)]
    stream = torch.cuda.Stream(device=dev)

    def enqueue(env, i):
        p = env._params
        p.noise_mode, p.noise_offset, p.clip_observations = noise, i, 100.0
        _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), phase,
                                             torch.cuda.current_stream(dev).cuda_stream))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(3):
            enqueue(envs[i % n_rep], i)
        stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            for i in range(steps):
                enqueue(envs[i % n_rep], i)
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream); g.replay(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / steps)
    ts.sort()
    return ts[len(ts) // 2]


if __name__ == "__main__":
    P = _lib
    rows = []
    for n, n_rep in ((4096, 1), (4096, 22), (1024, 1), (2048, 1), (8192, 12), (16384, 6), (65536, 3)):
        rows.append(("fused", n, n_rep, time_cfg("anymal_c_rough", n, n_rep, P.PHASE_FUSED)))
        print(rows[-1], flush=True)
    for name, ph in (("history", P.PHASE_HISTORY), ("derive", P.PHASE_DERIVE), ("obs", P.PHASE_OBS), ("reward", P.PHASE_REWARD),
                     ("term", P.PHASE_TERMINATION)):
        rows.append((name, 4096, 22, time_cfg("anymal_c_rough", 4096, 22, ph)))
        print(rows[-1], flush=True)
    rows.append(("fused-nonoise", 4096, 22, time_cfg("anymal_c_rough", 4096, 22, P.PHASE_FUSED, noise=_lib.NOISE_OFF)))
    print(rows[-1], flush=True)
    rows.append(("fused-flat", 4096, 22, time_cfg("anymal_c_flat", 4096, 22, P.PHASE_FUSED)))
    print(rows[-1], flush=True)
    lib.elg_set_step_tuning(0, 0, 0, 1)
    rows.append(("fused-nobulk", 4096, 22, time_cfg("anymal_c_rough", 4096, 22, P.PHASE_FUSED)))
    print(rows[-1], flush=True)
    lib.elg_set_step_tuning(0, 0, 0, 0)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "step_sweep.json"), "w"))
