"""Config of the kinematic planning variant (mirrors envs/batch_rollout/robot_plan_grad_sampling_config.py:30-62 of the
reference; the trajectory-optimiser / warm-start sections configure the external ``traj_sampling`` package and are not here)."""
from .robot_batch_rollout_config import RobotBatchRolloutCfg, RobotBatchRolloutCfgPPO


class RobotPlanGradSamplingCfg(RobotBatchRolloutCfg):
    class planning:
        integration_method = "euler"      # "euler" | "rk4"
        max_base_lin_vel = 3.0            # m/s
        max_base_ang_vel = 2.0            # rad/s
        max_joint_vel = 10.0              # rad/s
        max_integration_step = 0.01       # s; dt is split into ceil(dt / this) equal sub-steps
        enforce_joint_limits = False
        state_vel_noise_scale = 1.0


class RobotPlanGradSamplingCfgPPO(RobotBatchRolloutCfgPPO):
    pass
