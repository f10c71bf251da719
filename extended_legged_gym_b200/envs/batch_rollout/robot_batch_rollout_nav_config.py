"""Config of the navigation task (same attribute names and values as envs/batch_rollout/robot_batch_rollout_nav_config.py:7-48
of the reference; ``tests/test_nav_commands.py`` exercises every field)."""
from .robot_batch_rollout_config import RobotBatchRolloutCfg, RobotBatchRolloutCfgPPO



def _zero_range():
    """the navigation controller writes the commands; nothing is sampled"""
    return [0.0, 0.0]


class RobotBatchRolloutNavCfg(RobotBatchRolloutCfg):
    class navi_opt:
        # controller: v = clip(kp_linear * (goal - pos)), yaw rate = clip(kp_angular * heading error), exponentially smoothed
        kp_linear, kp_angular = 1.0, 2.0
        max_linear_vel, max_angular_vel = 1.0, 1.0          # m/s, rad/s
        cmd_smooth_factor = 0.1                              # weight of the previous command
        use_2d_nav = True                                    # planar distance / commands, z ignored
        tolerance_rad = 0.5                                  # goal radius [m]
        # per main env: one entry [x, y, z] / [x, y, z, w] for all, or a list with one entry per main env
        goal_pos = [5.0, 5.0, 0.5]
        start_pos = [0.0, 0.0, 0.5]
        start_quat = [0.0, 0.0, 0.0, 1.0]

    class commands(RobotBatchRolloutCfg.commands):
        class ranges:
            heading = _zero_range()
            ang_vel_yaw = _zero_range()
            lin_vel_y = _zero_range()
            lin_vel_x = _zero_range()

    class env(RobotBatchRolloutCfg.env):
        episode_length_s = 30


class RobotBatchRolloutNavCfgPPO(RobotBatchRolloutCfgPPO):
    class runner(RobotBatchRolloutCfgPPO.runner):
        max_iterations = 2000
        num_steps_per_env = 32
