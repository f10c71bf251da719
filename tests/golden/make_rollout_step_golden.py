"""Generate tests/golden/rollout_step.npz FROM THE UNMODIFIED REFERENCE (container only): the main / rollout variant of
the per-step path -- RobotBatchRollout.post_physics_step with its own _post_physics_step_callback, check_termination,
reset_idx, _reset_root_states, _push_robots, compute_reward / compute_observations and reward mixin
(envs/batch_rollout/robot_batch_rollout.py:718-1014, :1306-1413, robot_batch_rollout_rew_mixin.py) -- on synthetic state.

The reference class cannot run its __init__ (it creates the simulator), so the env is built like the other fixtures
(oracle/ref_harness.make_reference_env runs the base class's own _parse_cfg / _init_buffers on the synthetic tensors), the
object's class is then switched to RobotBatchRollout, and the rollout class's own _parse_cfg / _init_env_indices /
_prepare_reward_function run on it.  Every per-step method that runs afterwards is the rollout class's, unmodified.

    python tests/golden/make_rollout_step_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_harness as rh  # noqa: E402
import common  # noqa: E402
import make_golden as mg  # noqa: E402

EXTRA = ("env_origins", "terrain_levels")
CASES = {   # tag: (case, mains, rollouts, seed, steps, step counter before the first step)
    "a": ("anymal_c_rough", 8, 7, 0, 3, 0),
    "b": ("go2_all_terms_heading", 6, 5, 1, 3, 748),      # the second step is a push step (push_interval = 750)
}


def reference_rollout_env(case, num_main, rollouts, spec, st, hf):
    rh.install()
    from legged_gym.envs.batch_rollout.robot_batch_rollout import RobotBatchRollout as Ref
    env = rh.make_reference_env(mg.reference_cfg_for(case), spec, st, hf)
    keep = {k: getattr(env, k).clone() for k in ("commands",)}
    env.__class__ = Ref
    env.num_main_envs, env.num_rollout_per_main = num_main, rollouts
    env.total_num_envs = num_main * (1 + rollouts)
    env._parse_cfg(env.cfg)                 # (:1644-1674) episode length in whole steps, integer push interval, scales
    env._init_env_indices()                 # (:119-164)
    env._prepare_reward_function()          # (:1676-1703) binds the ROLLOUT reward mixin, episode sums over all rows
    env.commands[:] = keep["commands"]
    return env


def snapshot(env):
    snap = common.snapshot(env)
    for k in EXTRA:
        if hasattr(env, k):
            snap[k] = getattr(env, k).clone()
    names = list(env.episode_sums.keys())
    snap["episode_sums"] = torch.stack([env.episode_sums[k] for k in names])
    return snap, names


def run_reference(case, num_main, rollouts, seed, steps, counter0):
    n = num_main * (1 + rollouts)
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=True)
    inputs = {k: v.clone() for k, v in st.items()}
    hf = mg.height_field()
    env = reference_rollout_env(case, num_main, rollouts, spec, st, hf)
    env.common_step_counter = counter0
    g = torch.Generator().manual_seed(2000 + seed)
    out = {}
    for s in range(steps):
        noise_u = torch.rand(n, env.num_obs, generator=g)
        torch.manual_seed(5000 + 17 * s + seed)
        orig = torch.rand_like
        torch.rand_like = lambda t, *a, **k: noise_u.clone() if t.shape == noise_u.shape else orig(t, *a, **k)
        try:
            env.torques = env._compute_torques(env.actions).view(env.torques.shape)
            env.post_physics_step()
        finally:
            torch.rand_like = orig
        snap, names = snapshot(env)
        for k, v in snap.items():
            out[f"s{s}__{k}"] = v.numpy()
        out[f"s{s}__noise_u"] = noise_u.numpy()
        for k, v in env.extras.get("episode", {}).items():
            out[f"s{s}__extras__{k}"] = np.asarray(float(v))
    out["meta__sum_names"] = np.array(names)
    return inputs, out, env


def main():
    blob = {}
    for tag, (case, m, r, seed, steps, c0) in CASES.items():
        inputs, out, env = run_reference(case, m, r, seed, steps, c0)
        for k, v in inputs.items():
            blob[f"{tag}__in__{k}"] = v.numpy()
        for k, v in out.items():
            blob[f"{tag}__{k}"] = v
        blob[f"{tag}__meta"] = np.array([m, r, seed, steps, c0], dtype=np.int64)
    path = os.path.join(HERE, "rollout_step.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
