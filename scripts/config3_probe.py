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

# in-graph (warm instruction cache, programmatic dependent launch) time of every kernel of the step at this size
def graph_time(fn, reps=20):
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


from extended_legged_gym_b200 import _lib  # noqa: E402
flags = env._reset_bool.clone()
print("in-graph us per launch at", n, "envs:")
print("  torques ", round(graph_time(lambda: env._compute_torques(env.actions)), 2))
print("  resample", round(graph_time(env._launch_resample), 2))
print("  step    ", round(graph_time(lambda: env._launch(_lib.PHASE_FUSED, 100.0)), 2))
env._reset_bool.copy_(flags)
nres = int(env._reset_bool.sum())


def reset_only():
    rp, b = env._reset_native_synced()
    _lib.check(env._lib.elg_reset_envs(__import__("ctypes").byref(env._dims), __import__("ctypes").byref(rp), __import__("ctypes").byref(env._params),
                                       __import__("ctypes").byref(b), torch.cuda.current_stream(dev).cuda_stream), "elg_reset_envs")


print(f"  reset    {round(graph_time(reset_only), 2)}  ({nres} flagged envs)")
env._reset_bool.zero_()
env._reset_bool[::170] = True          # the same number of resets, but never two in one warp's 32 envs
print(f"  reset    {round(graph_time(reset_only), 2)}  ({int(env._reset_bool.sum())} flagged envs, at most one per 32-env group)")
env._reset_bool.zero_()
env._reset_bool[:96:32] = True
print(f"  reset    {round(graph_time(reset_only), 2)}  (3 flagged envs in 3 groups)")
env._reset_bool.zero_()
env._reset_bool[:3] = True
print(f"  reset    {round(graph_time(reset_only), 2)}  (3 flagged envs in one group)")
env._reset_bool.zero_()
print(f"  reset    {round(graph_time(reset_only), 2)}  (no env flagged)")
