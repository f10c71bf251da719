#!/usr/bin/env python
"""_sync_main_to_rollout at BASELINE config 5 (64 mains x 512 rollouts): TMA bulk-store kernel vs per-thread stores, CUDA graph of 50 syncs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, common
from extended_legged_gym_b200 import _lib, synthetic
from extended_legged_gym_b200.envs import RobotBatchRollout
from extended_legged_gym_b200.sim_backend import SyntheticSim
dev = "cuda:0"
lib = _lib.load()
hf = synthetic.make_height_field(seed=0).to(dev)
SWEEP = "--tiles" in sys.argv
for mains, rollouts, drift in (((64, 512, 0.0),) if SWEEP else ((64, 512, 0.0), (64, 512, 0.1), (512, 64, 0.0))):
    n = mains * (1 + rollouts)
    cfg, spec, st = common.make_case_state("anymal_c_rough", n, seed=3)
    cfg.env.num_envs, cfg.env.rollout_envs = mains, rollouts
    cfg.domain_rand.rollout_envs_sync_pos_drift = drift
    env = RobotBatchRollout(cfg, None, SyntheticSim(cfg, n, dev, spec=spec, height_samples=hf, state=st), dev, True)
    env.set_env_state(st)
    modes = (0, 1, 1 << 8, 2 << 8, 3 << 8, 4 << 8, 8 << 8)
    if SWEEP:      # rows per tile (bits 16+) x slices per main (bits 8-15); 2 = build the tiles only, 4 = bulk stores only
        modes = [0, 2, 4] + [(tr << 16) | (sl << 8) for tr in (32, 64, 128, 256) for sl in (1, 2, 4, 8, 16) if tr * sl <= 1024 and tr * sl >= 256]
    for no_bulk in modes:
        lib.elg_set_clone_tuning(no_bulk)
        gs = torch.cuda.Stream(device=dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(gs):
            env._sync_main_to_rollout(); gs.synchronize()
            with torch.cuda.graph(gr, stream=gs):
                for _ in range(50):
                    env._sync_main_to_rollout()
            gr.replay(); gs.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(gs)
            for _ in range(4):
                gr.replay()
            e1.record(gs); gs.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 200
        b = 388 * mains * rollouts
        print(f"{mains:5d} mains x {rollouts:4d} rollouts drift={drift}: { {0: 'TMA bulk', 1: 'per-thread', 2: 'build only', 4: 'stores only'}.get(no_bulk, 'TMA tile=%d slices=%d' % (no_bulk >> 16, (no_bulk >> 8) & 255)):26s} {us:7.2f} us  {b / us / 1e3:8.1f} GB/s written", flush=True)
    lib.elg_set_clone_tuning(0)
    del env
