"""Generate tests/golden/rollout_step.npz FROM THE UNMODIFIED REFERENCE (container only): the main / rollout variant of
the per-step path -- RobotBatchRollout.post_physics_step with its own _post_physics_step_callback, check_termination,
reset_idx, _reset_root_states, _push_robots, compute_reward / compute_observations and reward mixin
(envs/batch_rollout/robot_batch_rollout.py:718-1014, :1306-1413, robot_batch_rollout_rew_mixin.py) -- on synthetic state.

The reference class cannot run its __init__ (it creates the simulator), so the env is built like the other fixtures
(oracle/ref_harness.make_reference_env runs the base class's own _parse_cfg / _init_buffers on the synthetic tensors), the
object's class is then switched to RobotBatchRollout, and the rollout class's own _parse_cfg / _init_env_indices /
_prepare_reward_function run on it.  Every per-step method that runs afterwards is the rollout class's, unmodified.

    python tests/golden/make_rollout_step_golden.py            # rollout_step.npz (tags a, b: RobotBatchRollout)
    python tests/golden/make_rollout_step_golden.py --robot    # rollout_step_anymal.npz (tags c, d: AnymalCBatchRollout, ElSpiderAirBatchRollout)
    python tests/golden/make_rollout_step_golden.py --rollout-mode   # rollout_mode_step.npz: post_physics_step_rollout (:763-817) of all four

Tag c switches the object's class to the robot-specific ``AnymalCBatchRollout`` (envs/anymal_c/batch_rollout/
anymal_c_batch_rollout.py:49-225: upside-down MAIN rows terminate :192-199, the gait scheduler follows the env clock :143-150)
on the anymal_c_rough state, with a few main and rollout robots turned upside down.
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


ROBOT_CASES = {   # tag: (case, mains, rollouts, seed, steps, step counter before the first step)
    "c": ("anymal_c_rough", 8, 7, 3, 3, 0),         # AnymalCBatchRollout: upside-down MAIN rows reset, gait scheduler on the env clock
    "d": ("elspider_air_rough", 6, 4, 5, 3, 0),     # ElSpiderAirBatchRollout (elspider_air_batch_rollout.py:46-230): EVERY upside-down row resets
}
UPSIDE_DOWN_ROWS = {"c": (0, 3, 8, 21, 40, 42),      # mains 0, 8, 40 (rows k * 8) and rollout rows 3, 21, 42
                    "d": (0, 2, 10, 13, 27)}          # mains 0, 10 (rows k * 5) and rollout rows 2, 13, 27


def reference_rollout_env(case, num_main, rollouts, spec, st, hf, robot=False):
    rh.install()
    if robot == "d":
        from legged_gym.envs.elspider_air.batch_rollout.elspider_air_batch_rollout import ElSpiderAirBatchRollout as Ref
    elif robot:
        from legged_gym.envs.anymal_c.batch_rollout.anymal_c_batch_rollout import AnymalCBatchRollout as Ref
    else:
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
    if robot == "d":      # what ElSpiderAirBatchRollout._init_buffers adds (:143-150): the actuator-network state reset_idx clears
        n, A = env.total_num_envs, env.num_actions
        env.sea_hidden_state = torch.zeros(2, n * A, 8)
        env.sea_cell_state = torch.zeros(2, n * A, 8)
        env.sea_hidden_state_per_env = env.sea_hidden_state.view(2, n, A, 8)
        env.sea_cell_state_per_env = env.sea_cell_state.view(2, n, A, 8)
        env.noise_scale_vec = env._get_noise_scale_vec(env.cfg)      # the class's own 18-DOF slices (:94-117; the harness ran the base class's)
    if robot == "c":      # what AnymalCBatchRollout.__init__ adds (:58-98): the reference's own scheduler object with the class's config
        from legged_gym.utils import GaitScheduler
        from legged_gym.envs.anymal_c.batch_rollout.anymal_c_batch_rollout_config import AnymalCBatchRolloutCfg
        env.gait_scheduler = GaitScheduler(None, env.base_quat, env.base_lin_vel, env.base_ang_vel, env.projected_gravity, env.dof_pos,
                                           env.dof_vel, env.foot_positions, env.foot_velocities, env.total_num_envs, env.device,
                                           gait_cfg=AnymalCBatchRolloutCfg.gait_scheduler)
        env.t_main = 0.0
    return env


def snapshot(env):
    snap = common.snapshot(env)
    for k in EXTRA:
        if hasattr(env, k):
            snap[k] = getattr(env, k).clone()
    names = list(env.episode_sums.keys())
    snap["episode_sums"] = torch.stack([env.episode_sums[k] for k in names])
    return snap, names


def run_reference(case, num_main, rollouts, seed, steps, counter0, robot=False):
    n = num_main * (1 + rollouts)
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=True)
    if robot:
        for row in UPSIDE_DOWN_ROWS[robot]:      # half a turn about the body x axis (xyzw): projected_gravity.z = +1
            st["root_states"][row, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    inputs = {k: v.clone() for k, v in st.items()}
    hf = mg.height_field()
    env = reference_rollout_env(case, num_main, rollouts, spec, st, hf, robot=robot)
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
        if robot == "c":
            snap["gait_idx"] = env.gait_scheduler.gait_idx.clone()
            env.t_main += env.dt          # (RobotBatchRollout.step :597, after post_physics_step)
        for k, v in snap.items():
            out[f"s{s}__{k}"] = v.numpy()
        out[f"s{s}__noise_u"] = noise_u.numpy()
        for k, v in env.extras.get("episode", {}).items():
            out[f"s{s}__extras__{k}"] = np.asarray(float(v))
    out["meta__sum_names"] = np.array(names)
    return inputs, out, env


def run_reference_rollout_mode(tag, case, num_main, rollouts, seed, counter0, robot):
    """one main step (as in run_reference), fresh actions, torques, then the class's own ``post_physics_step_rollout``
    (robot_batch_rollout.py:763-817 + the robot classes' scheduler call): the state of the ROLLOUT rows afterwards"""
    n = num_main * (1 + rollouts)
    cfg, spec, st = common.make_case_state(case, n, seed=seed, adversarial=True)
    if robot:
        for row in UPSIDE_DOWN_ROWS[robot]:
            st["root_states"][row, 3:7] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    inputs = {k: v.clone() for k, v in st.items()}
    env = reference_rollout_env(case, num_main, rollouts, spec, st, mg.height_field(), robot=robot)
    env.common_step_counter = counter0
    if robot == "d":      # ElSpiderAirBatchRollout.__init__ (:64-78): the reference scheduler object with the class's config
        from legged_gym.utils import GaitScheduler
        from legged_gym.envs.elspider_air.batch_rollout.elspider_air_batch_rollout_config import ElSpiderAirBatchRolloutCfg
        env.gait_scheduler = GaitScheduler(None, env.base_quat, env.base_lin_vel, env.base_ang_vel, env.projected_gravity, env.dof_pos,
                                           env.dof_vel, env.foot_positions, env.foot_velocities, env.total_num_envs, env.device,
                                           gait_cfg=ElSpiderAirBatchRolloutCfg.gait_scheduler)
    g = torch.Generator().manual_seed(2000 + seed)
    u_main, u_roll = torch.rand(n, env.num_obs, generator=g), torch.rand(n, env.num_obs, generator=g)
    new_actions = torch.randn(n, env.num_actions, generator=g)
    orig = torch.rand_like
    out = {"noise_u_main": u_main.numpy(), "noise_u_rollout": u_roll.numpy(), "actions_rollout": new_actions.numpy()}
    try:
        torch.manual_seed(5000 + seed)
        torch.rand_like = lambda t, *a, **k: u_main.clone() if t.shape == u_main.shape else orig(t, *a, **k)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step()
        if robot == "c":
            env.t_main += env.dt
        env.t_rollout = getattr(env, "t_main", 0.0)
        env.actions[:] = new_actions
        torch.rand_like = lambda t, *a, **k: u_roll.clone() if t.shape == u_roll.shape else orig(t, *a, **k)
        env.torques = env._compute_torques(env.actions).view(env.torques.shape)
        env.post_physics_step_rollout()
    finally:
        torch.rand_like = orig
    snap, names = snapshot(env)
    if robot:
        snap["gait_idx"] = env.gait_scheduler.gait_idx.clone()
    for k, v in snap.items():
        out[k] = v.numpy()
    return inputs, out


def main():
    if "--rollout-mode" in sys.argv:
        blob = {}
        for tag, (case, m, r, seed, steps, c0) in {**CASES, **ROBOT_CASES}.items():
            inputs, out = run_reference_rollout_mode(tag, case, m, r, seed, c0, tag if tag in ROBOT_CASES else False)
            for k, v in inputs.items():
                blob[f"{tag}__in__{k}"] = v.numpy()
            for k, v in out.items():
                blob[f"{tag}__out__{k}"] = v
            blob[f"{tag}__meta"] = np.array([m, r, seed, steps, c0], dtype=np.int64)
        path = os.path.join(HERE, "rollout_mode_step.npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
        return
    blob = {}
    robot = "--robot" in sys.argv
    for tag, (case, m, r, seed, steps, c0) in (ROBOT_CASES if robot else CASES).items():
        inputs, out, env = run_reference(case, m, r, seed, steps, c0, robot=tag if robot else False)
        for k, v in inputs.items():
            blob[f"{tag}__in__{k}"] = v.numpy()
        for k, v in out.items():
            blob[f"{tag}__{k}"] = v
        blob[f"{tag}__meta"] = np.array([m, r, seed, steps, c0], dtype=np.int64)
    path = os.path.join(HERE, "rollout_step_anymal.npz" if robot else "rollout_step.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
