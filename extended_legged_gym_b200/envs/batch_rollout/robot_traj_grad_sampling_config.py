"""Config of the trajectory-optimising rollout env (mirrors envs/batch_rollout/robot_traj_grad_sampling_config.py:36-71 of the
reference: ``trajectory_opt.*`` -- only the fields the in-tree MPPI update uses are read here; the spline interpolation and
the WBFO / AVWBFO update rules belong to the external ``traj_sampling`` package)."""
from .robot_batch_rollout_config import RobotBatchRolloutCfg, RobotBatchRolloutCfgPPO


class RobotTrajGradSamplingCfg(RobotBatchRolloutCfg):
    class env(RobotBatchRolloutCfg.env):
        num_envs = 1         # main environments
        rollout_envs = 128   # rollout environments per main environment

    class trajectory_opt:
        enable_traj_opt = True
        num_diffuse_steps = 2
        num_diffuse_steps_init = 10
        num_samples = 127
        temp_sample = 0.05
        horizon_samples = 16
        horizon_nodes = 4
        horizon_diffuse_factor = 0.9
        traj_diffuse_factor = 0.5
        noise_scaling = 1.0
        update_method = "mppi"
        gamma = 0.99
        interp_method = "linear"
        compute_predictions = False


class RobotTrajGradSamplingCfgPPO(RobotBatchRolloutCfgPPO):
    pass
