"""Config of the navigation task (mirrors envs/batch_rollout/robot_batch_rollout_nav_config.py:7-48 of the reference)."""
from .robot_batch_rollout_config import RobotBatchRolloutCfg, RobotBatchRolloutCfgPPO


class RobotBatchRolloutNavCfg(RobotBatchRolloutCfg):
    class env(RobotBatchRolloutCfg.env):
        episode_length_s = 30

    class navi_opt:
        start_pos = [0.0, 0.0, 0.5]          # one pose [x, y, z] or a list of poses, one per main env
        start_quat = [0.0, 0.0, 0.0, 1.0]
        goal_pos = [5.0, 5.0, 0.5]
        tolerance_rad = 0.5
        max_linear_vel = 1.0
        max_angular_vel = 1.0
        kp_linear = 1.0
        kp_angular = 2.0
        cmd_smooth_factor = 0.1
        use_2d_nav = True

    class commands(RobotBatchRolloutCfg.commands):
        class ranges:
            lin_vel_x = [0.0, 0.0]
            lin_vel_y = [0.0, 0.0]
            ang_vel_yaw = [0.0, 0.0]
            heading = [0.0, 0.0]


class RobotBatchRolloutNavCfgPPO(RobotBatchRolloutCfgPPO):
    class runner(RobotBatchRolloutCfgPPO.runner):
        num_steps_per_env = 32
        max_iterations = 2000
