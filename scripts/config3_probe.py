#!/usr/bin/env python
"""Six steps of BASELINE configs[2] (a1 rough, 65 536 envs, in-kernel reset path) launched eagerly: run under
`ncu --metrics gpu__time_duration.sum -k regex:elg_` for a per-kernel launch list at this size.  Not a benchmark."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from extended_legged_gym_b200 import synthetic  # noqa: E402
from extended_legged_gym_b200.envs import LeggedRobot  # noqa: E402
from extended_legged_gym_b200.sim_backend import SyntheticSim  # noqa: E402

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
case = sys.argv[2] if len(sys.argv) > 2 else "a1_rough"
hf = synthetic.make_height_field(seed=0).to(dev)
cfg, spec, st = common.make_case_state(case, n, seed=1)
cfg.env.num_envs = n
cfg.domain_rand.push_robots = False
env = LeggedRobot(cfg, None, SyntheticSim(cfg, n, dev, spec=spec, height_samples=hf, state=st), dev, True)
env.set_env_state(st)
env.noise_u = None
env._obs_clip_for_step = 100.0
for i in range(6):
    env.torques = env._compute_torques(env.actions).view(env.torques.shape)
    env.post_physics_step()
torch.cuda.synchronize()
print("resets in the last step:", int(env.reset_buf.sum()))
