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


class RobotBatchRolloutCfgPPO(LeggedRobotCfgPPO):
    class runner(LeggedRobotCfgPPO.runner):
        num_steps_per_env = 24
        max_iterations = 1500
