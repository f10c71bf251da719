"""Config of the main/rollout env layout (mirrors envs/batch_rollout/robot_batch_rollout_config.py:35-71 of the
reference: ``env.rollout_envs``, ``domain_rand.rollout_envs_sync_pos_drift``, ``viewer.render_rollouts``)."""
from ..base.legged_robot_config import LeggedRobotCfg, LeggedRobotCfgPPO


class RobotBatchRolloutCfg(LeggedRobotCfg):
    class env(LeggedRobotCfg.env):
        num_envs = 64        # main environments
        rollout_envs = 32    # rollout environments per main environment
        env_spacing = 4.0
        episode_length_s = 20

    class viewer(LeggedRobotCfg.viewer):
        render_rollouts = False

    class domain_rand(LeggedRobotCfg.domain_rand):
        rollout_envs_sync_pos_drift = 0.0


class RobotBatchRolloutPerceptCfg(RobotBatchRolloutCfg):
    """envs/batch_rollout/robot_batch_rollout_percept_config.py:35-85: the ray caster block of the base config + the SDF block"""
    class sdf:
        enable_sdf = False
        mesh_paths = []
        max_distance = 10.0
        enable_caching = True
        update_freq = 5
        query_bodies = []                # e.g. ["base", "LF_FOOT", ...]
        collision_sphere_radius = []
        collision_sphere_pos = []        # [x, y, z] per query body, body frame
        compute_gradients = True
        compute_nearest_points = True
        include_in_obs = True


class RobotBatchRolloutCfgPPO(LeggedRobotCfgPPO):
    class runner(LeggedRobotCfgPPO.runner):
        num_steps_per_env = 24
        max_iterations = 1500
