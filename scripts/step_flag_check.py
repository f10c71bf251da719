#!/usr/bin/env python
"""Outputs of the lean step kernel under an experiment flag must equal the default path bit for bit. usage: step_flag_check.py flag ..."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from step_sweep import make, lib, dev, _lib
outs = {}
for f in [0] + [int(a) for a in sys.argv[1:]]:
    env = make("anymal_c_rough", 4096, 0)
    p = env._params; p.noise_mode, p.clip_observations = _lib.NOISE_PHILOX, 100.0
    lib.elg_set_step_tuning(0, f, 0, 0)
    _lib.check(lib.elg_post_physics_step(C.byref(env._dims), C.byref(p), C.byref(env._bufs), _lib.PHASE_FUSED, None))
    torch.cuda.synchronize()
    outs[f] = [t.clone() for t in (env.obs_buf, env.rew_buf, env.measured_heights, env._episode_sums_all, env.last_actions, env.feet_air_time,
                                   env._reset_bool, env.base_lin_acc, env.foot_positions, env.episode_length_buf)]
    if f:
        print(f"flag {f} == default:", all(torch.equal(a, b) for a, b in zip(outs[0], outs[f])))
lib.elg_set_step_tuning(0, 0, 0, 0)
