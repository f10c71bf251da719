"""Gait schedulers of the reference (utils/gait_scheduler.py in /root/reference/legged_gym/legged_gym).

``GaitScheduler`` (:28-81: phase clock + swing-height tracking) runs INSIDE the step kernels here (``ElgStepParams.gait_*``,
``LeggedRobot.gait_cfg``); this module keeps its config class for API parity and holds the hexapod's ``AsyncGaitScheduler``
(:104-173): three posture terms -- joint alignment inside the two tripods, weighted distance to a nominal joint pose, foot
height alignment inside the tripods -- combined by ``ElSpider._reward_async_gait_scheduler`` (envs/elspider_air/elspider.py:351-363,
enabled e.g. by flat/pose_elspider_air_flat_config.py:66).  They are small per-env reductions over 3-element sets and run as
torch ops on the device, as one Python-side reward term next to the kernel's built-ins.
"""
import torch


class GaitSchedulerCfg(object):
    period = 1.0
    duty = 0.5
    foot_phases = [0.0, 0.5, 0.0, 0.5, 0.0, 0.5]
    dt = 0.02
    swing_height = 0.04
    track_sigma = 0.25


class AsyncGaitSchedulerCfg(object):
    # same tag should keep same motion (:85-102)
    dof_names = ['LB_HAA', 'LB_HFE', 'LB_KFE', 'LF_HAA', 'LF_HFE', 'LF_KFE', 'LM_HAA', 'LM_HFE', 'LM_KFE',
                 'RB_HAA', 'RB_HFE', 'RB_KFE', 'RF_HAA', 'RF_HFE', 'RF_KFE', 'RM_HAA', 'RM_HFE', 'RM_KFE']
    dof_align_sets = [['RF_HFE', 'RB_HFE', 'LM_HFE'], ['LF_HFE', 'LB_HFE', 'RM_HFE'],
                      ['RF_KFE', 'RB_KFE', 'LM_KFE'], ['LF_KFE', 'LB_KFE', 'RM_KFE']]
    dof_nominal_pos = [0.0, 1.0, 1.0] * 6          # HAA, HFE, KFE
    dof_nominal_pos_weight = [1.0, 1.0, 3.0] * 6
    foot_names = ['LB_FOOT', 'LF_FOOT', 'LM_FOOT', 'RB_FOOT', 'RF_FOOT', 'RM_FOOT']
    foot_z_align_sets = [['RF_FOOT', 'RB_FOOT', 'LM_FOOT'], ['LF_FOOT', 'LB_FOOT', 'RM_FOOT']]

    def __init__(self) -> None:
        self.dof_align_sets_idx = [[self.dof_names.index(d) for d in s] for s in self.dof_align_sets]
        self.foot_z_align_sets_idx = [[self.foot_names.index(f) for f in s] for s in self.foot_z_align_sets]


class AsyncGaitScheduler(object):
    """Holds REFERENCES to the robot state tensors, like the reference class (:104-131).  Note what that means there:
    ``dof_pos`` is a view of the simulator's dof_state and stays current, but ``foot_pos`` is the ``foot_positions`` tensor of
    construction time -- the reference env rebinds ``self.foot_positions`` to a new tensor every step (legged_robot.py:136), so
    its scheduler keeps reading the initial one.  The host class here passes a snapshot to reproduce exactly that."""

    def __init__(self, height_samples, base_quat, base_lin_vel, base_ang_vel, projected_gravity, dof_pos, dof_vel, foot_pos, foot_vel,
                 num_envs, device, gait_cfg: AsyncGaitSchedulerCfg = None) -> None:
        self.height_samples = height_samples
        self.base_quat, self.base_lin_vel, self.base_ang_vel, self.projected_gravity = base_quat, base_lin_vel, base_ang_vel, projected_gravity
        self.dof_pos, self.dof_vel, self.foot_pos, self.foot_vel = dof_pos, dof_vel, foot_pos, foot_vel
        self.num_envs, self.device = num_envs, device
        self.gait_cfg = gait_cfg if gait_cfg is not None else AsyncGaitSchedulerCfg()

    def reward_dof_align(self):
        reward = torch.zeros(self.num_envs, device=self.device, dtype=torch.float)
        for idx in self.gait_cfg.dof_align_sets_idx:
            reward += torch.std(self.dof_pos[:, idx], dim=1)
        return reward

    def reward_dof_nominal_pos(self):
        nominal = torch.tensor(self.gait_cfg.dof_nominal_pos, device=self.device, dtype=torch.float).repeat(self.num_envs, 1)
        err = torch.square(self.dof_pos - nominal)
        weight = torch.tensor(self.gait_cfg.dof_nominal_pos_weight, device=self.device, dtype=torch.float).repeat(self.num_envs, 1)
        return torch.sum(err * weight, dim=1)

    def reward_foot_z_align(self):
        """z align is only for flat env"""
        reward = torch.zeros(self.num_envs, device=self.device, dtype=torch.float)
        for idx in self.gait_cfg.foot_z_align_sets_idx:
            reward += torch.std(self.foot_pos[:, idx, 2], dim=1)
        return reward
