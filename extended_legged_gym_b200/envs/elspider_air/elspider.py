"""``ElSpider`` -- hexapod task class of the reference (envs/elspider_air/elspider.py:225-408 in
/root/reference/legged_gym/legged_gym): ``LeggedRobot`` with 18 DOF / 6 feet plus

  _compute_torques()        :295-310  actuator-network torque path (same LSTMsea as ANYmal; ``Anymal``'s implementation, one thread
                                      per (env, dof), works for any DOF count) or the PD controller
  check_termination()       :336-345  additionally resets envs that are upside down (projected_gravity.z > 0)
  _reward_gait_2_step()     :365-408  tripod form: feet (LB, LF, RM) in phase, (LM, RB, RF) in phase, the groups in anti-phase
  _get_noise_scale_vec()    :312-332  18-DOF slices (the base class here is already generic in the DOF count)
  gait scheduler            :237-252  period 1.4 s, swing height 0.07 m, six phases alternating 0 / 0.5

  async gait scheduler      :254-266, :351-363  ``_reward_async_gait_scheduler``: three posture terms of ``AsyncGaitScheduler``
                                      weighted per reward stage by ``cfg.rewards.async_gait_scheduler`` (enabled by the reference's
                                      pose / foot-track / trajectory-sampling hexapod configs)

gait_2_step and the upside-down check run inside the generic step kernel (``ElgStepParams.gait_2_step_hexapod`` /
``terminate_upside_down``); the async-scheduler term is a Python-side reward term (torch ops on the device, ``_python_terms``).
"""
from types import SimpleNamespace

from ...utils.gait_scheduler import AsyncGaitScheduler, AsyncGaitSchedulerCfg
from ...utils.helpers import class_to_dict
from ..anymal_c.anymal import Anymal
from ..base.legged_robot_rew_mixin import _stock


class ElSpider(Anymal):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        if len(self.feet_indices) != 6:
            raise ValueError(f"ElSpider expects six feet, the asset has {len(self.feet_indices)}")
        # GaitSchedulerCfg defaults for six feet (utils/gait_scheduler.py:18-25) with the overrides of elspider.py:237-240
        self.gait_cfg = SimpleNamespace(dt=self.dt, period=1.4, foot_phases=[0.0, 0.5, 0.0, 0.5, 0.0, 0.5], swing_height=0.07)
        self._params_dirty = True
        # (:254-266) the scheduler keeps the foot_positions tensor of construction time (see AsyncGaitScheduler's docstring)
        self.async_gait_scheduler = AsyncGaitScheduler(self.height_samples, self.base_quat, self.base_lin_vel, self.base_ang_vel,
                                                       self.projected_gravity, self.dof_pos, self.dof_vel, self.foot_positions.clone(),
                                                       self.foot_velocities.clone(), self.num_envs, self.device, AsyncGaitSchedulerCfg())

    def _reward_async_gait_scheduler(self):
        """(:351-363) stage-dependent weights from cfg.rewards.async_gait_scheduler"""
        scales = class_to_dict(self.cfg.rewards.async_gait_scheduler)

        def weight(key, stage):
            v = scales[key]
            return v[min(stage, len(v) - 1)] if isinstance(v, list) else v
        g, st = self.async_gait_scheduler, self.reward_scales_stage
        return g.reward_dof_align() * weight("dof_align", st) + g.reward_dof_nominal_pos() * weight("dof_nominal_pos", st) + \
            g.reward_foot_z_align() * weight("reward_foot_z_align", st)

    def _native_params(self):
        p = super()._native_params()
        p.gait_2_step_hexapod = 1
        p.terminate_upside_down = 1
        return p

    @_stock
    def _reward_gait_2_step(self):
        sync, anti = self._sync_reward_func, self._async_reward_func
        g1 = (sync(0, 1) + sync(0, 5) + sync(1, 5)) / 3
        g2 = (sync(2, 3) + sync(2, 4) + sync(3, 4)) / 3
        across = (anti(0, 2) + anti(0, 3) + anti(0, 4) + anti(1, 2) + anti(1, 3) + anti(1, 4) + anti(5, 2) + anti(5, 3) + anti(5, 4)) / 9
        turn = self.commands[:, 3] if self.cfg.commands.heading_command else self.commands[:, 2]
        moving = (self._cmd_speed() > self.speed_min) | (turn.abs() >= self.speed_min / 2)
        return ((g1 + g2) / 2 + across) * moving
