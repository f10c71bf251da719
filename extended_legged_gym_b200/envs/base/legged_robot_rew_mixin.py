"""The ``_reward_*`` registry as Python-visible methods.

The fused kernel (csrc/elg_step.cu) evaluates every *stock* term of the reference registry
(envs/base/legged_robot_rew_mixin.py:41-234) itself; these methods exist because the registry
is part of the drop-in API: ``getattr(env, '_reward_' + name)`` must resolve
(legged_robot.py:664-670), subclasses may override a term or add new ones, and user code calls
terms directly.  A term whose bound method is still the stock one below is routed to the
kernel; an overridden or new term is evaluated here with torch on the device and handed to the
kernel as ``extra_reward`` (added before the only-positive clip).

``_STOCK`` marks the stock implementations so overrides can be detected.
"""
import torch

from ...utils.helpers import class_to_dict


def _stock(fn):
    fn._elg_stock = True
    return fn


class LeggedRobotRewMixin:
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.speed_min = 0.1

    # ---- scale tables (legged_robot_rew_mixin.py:15-38) -----------------------------------------
    def _get_reward_scales(self, stage=0):
        self.reward_scales_dict = class_to_dict(self.cfg.rewards.scales)
        if not self.cfg.rewards.multi_stage_rewards:
            return self.reward_scales_dict
        picked = {}
        for name, val in self.reward_scales_dict.items():
            picked[name] = val if not isinstance(val, list) else val[min(stage, len(val) - 1)]
        return picked

    def update_reward_scales(self, mean_reward):
        rw = self.cfg.rewards
        if mean_reward > rw.reward_stage_threshold and self.reward_scales_stage < rw.reward_max_stage:
            self.reward_scales_stage += 1
            self.reward_scales = self._get_reward_scales(self.reward_scales_stage)
            self._prepare_reward_function()
            return True
        return False

    # ---- shared sub-expressions --------------------------------------------------------------------
    def _feet_forces(self):
        return self.contact_forces[:, self.feet_indices, :]

    def _feet_contact(self):
        touching = self._feet_forces()[..., 2] > 1.0
        return touching, touching | self.last_contacts

    def _feet_stumbling(self):
        f = self._feet_forces()
        return f[..., :2].norm(dim=-1) > 5 * f[..., 2].abs()

    def _cmd_speed(self):
        return self.commands[:, :2].norm(dim=1)

    # ---- base ---------------------------------------------------------------------------------------
    @_stock
    def _reward_lin_vel_z(self):
        return self.base_lin_vel[:, 2] ** 2

    @_stock
    def _reward_ang_vel_xy(self):
        return (self.base_ang_vel[:, :2] ** 2).sum(dim=1)

    @_stock
    def _reward_orientation(self):
        return (self.projected_gravity[:, :2] ** 2).sum(dim=1)

    @_stock
    def _reward_base_height(self):
        clearance = (self.root_states[:, 2:3] - self.measured_heights).mean(dim=1)
        return (clearance - self.cfg.rewards.base_height_target) ** 2

    @_stock
    def _reward_base_foot_height(self):
        target = self.cfg.rewards.base_height_target
        on_ground = self.feet_contact_time > 1e-3
        z = self.foot_positions[:, :, 2]
        cnt = on_ground.sum(dim=1)
        mean_z = (z * on_ground).sum(dim=1) / cnt.clamp(min=1)
        ground = torch.where(cnt > 0, mean_z, self.root_states[:, 2] - target)
        return (self.root_states[:, 2] - ground - target) ** 2

    @_stock
    def _reward_tracking_lin_vel(self):
        err = ((self.commands[:, :2] - self.base_lin_vel[:, :2]) ** 2).sum(dim=1)
        return torch.exp(-err / self.cfg.rewards.tracking_sigma)

    @_stock
    def _reward_tracking_ang_vel(self):
        err = (self.commands[:, 2] - self.base_ang_vel[:, 2]) ** 2
        return torch.exp(-err / self.cfg.rewards.tracking_sigma)

    # ---- joints -------------------------------------------------------------------------------------
    @_stock
    def _reward_torques(self):
        return (self.torques ** 2).sum(dim=1)

    @_stock
    def _reward_dof_vel(self):
        return (self.dof_vel ** 2).sum(dim=1)

    @_stock
    def _reward_dof_acc(self):
        return (((self.last_dof_vel - self.dof_vel) / self.dt) ** 2).sum(dim=1)

    @_stock
    def _reward_action_rate(self):
        return ((self.last_actions - self.actions) ** 2).sum(dim=1)

    @_stock
    def _reward_dof_pos_limits(self):
        below = (self.dof_pos - self.dof_pos_limits[:, 0]).clamp(max=0.0)
        above = (self.dof_pos - self.dof_pos_limits[:, 1]).clamp(min=0.0)
        return (above - below).sum(dim=1)

    @_stock
    def _reward_dof_vel_limits(self):
        over = self.dof_vel.abs() - self.dof_vel_limits * self.cfg.rewards.soft_dof_vel_limit
        return over.clamp(min=0.0, max=1.0).sum(dim=1)

    @_stock
    def _reward_torque_limits(self):
        over = self.torques.abs() - self.torque_limits * self.cfg.rewards.soft_torque_limit
        return over.clamp(min=0.0).sum(dim=1)

    @_stock
    def _reward_stand_still(self):
        return (self.dof_pos - self.default_dof_pos).abs().sum(dim=1) * (self._cmd_speed() < self.speed_min)

    # ---- contacts -----------------------------------------------------------------------------------
    @_stock
    def _reward_collision(self):
        f = self.contact_forces[:, self.penalised_contact_indices, :]
        return (f.norm(dim=-1) > 0.1).float().sum(dim=1)

    @_stock
    def _reward_feet_stumble(self):
        return self._feet_stumbling().any(dim=1)

    @_stock
    def _reward_feet_stumble_liftup(self):
        return (self._feet_stumbling() * self.foot_velocities[:, :, 2]).sum(dim=1)

    @_stock
    def _reward_feet_slip(self):
        _, filt = self._feet_contact()
        speed_sq = self.foot_velocities[:, :, 0:2].norm(dim=2) ** 2
        return (filt * speed_sq).sum(dim=1)

    @_stock
    def _reward_jump_air(self):
        _, filt = self._feet_contact()
        return ((~filt) * (self.feet_air_time - 0.5)).sum(dim=1).sub(len(self.feet_indices) / 2).clamp(min=0.0)

    @_stock
    def _reward_feet_air_time(self):
        # side effects are part of the contract: last_contacts is rebound, both timers advance
        touching, filt = self._feet_contact()
        self.last_contacts = touching
        landed = (self.feet_air_time > 0.0) * filt
        self.feet_air_time += self.dt
        self.feet_contact_time += self.dt
        rew = ((self.feet_air_time - 0.5) * landed).sum(dim=1)
        rew *= self._cmd_speed() > 0.1
        self.feet_air_time *= ~filt
        self.feet_contact_time *= filt
        return rew

    @_stock
    def _reward_feet_contact_forces(self):
        return (self._feet_forces().norm(dim=-1) - self.cfg.rewards.max_contact_force).clamp(min=0.0).sum(dim=1)

    @_stock
    def _reward_four_footup(self):
        self.all_feet_up = (self._feet_forces()[..., 2] < 1).all(dim=1)
        return 0.1 * self.all_feet_up.float()

    # ---- trot gait ----------------------------------------------------------------------------------
    def _sync_reward_func(self, foot_0: int, foot_1: int, max_err=2):
        air, con = self.feet_air_time, self.feet_contact_time
        return ((air[:, foot_0] - air[:, foot_1]) ** 2).clamp(max=max_err ** 2) + \
            ((con[:, foot_0] - con[:, foot_1]) ** 2).clamp(max=max_err ** 2)

    def _async_reward_func(self, foot_0: int, foot_1: int, max_err=2):
        air, con = self.feet_air_time, self.feet_contact_time
        return ((air[:, foot_0] - con[:, foot_1]) ** 2).clamp(max=max_err ** 2) + \
            ((con[:, foot_0] - air[:, foot_1]) ** 2).clamp(max=max_err ** 2)

    @_stock
    def _reward_gait_2_step(self):
        in_phase = (self._sync_reward_func(0, 3) + self._sync_reward_func(1, 2)) / 2
        anti = (self._async_reward_func(0, 1) + self._async_reward_func(0, 2) +
                self._async_reward_func(3, 2) + self._async_reward_func(3, 1)) / 4
        turn = self.commands[:, 3] if self.cfg.commands.heading_command else self.commands[:, 2]
        moving = (self._cmd_speed() > self.speed_min) | (turn.abs() >= self.speed_min / 2)
        return (in_phase + anti) * moving

    # ---- misc ---------------------------------------------------------------------------------------
    @_stock
    def _reward_termination(self):
        return self.reset_buf * ~self.time_out_buf
