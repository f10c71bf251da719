"""``ElSpider`` -- hexapod task class of the reference (envs/elspider_air/elspider.py:225-408 in
/root/reference/legged_gym/legged_gym): ``LeggedRobot`` with 18 DOF / 6 feet plus

  _compute_torques()        :295-310  actuator-network torque path (same LSTMsea as ANYmal; ``Anymal``'s implementation, one thread
                                      per (env, dof), works for any DOF count) or the PD controller
  check_termination()       :336-345  additionally resets envs that are upside down (projected_gravity.z > 0)
  _reward_gait_2_step()     :365-408  tripod form: feet (LB, LF, RM) in phase, (LM, RB, RF) in phase, the groups in anti-phase
  _get_noise_scale_vec()    :312-332  18-DOF slices (the base class here is already generic in the DOF count)
  gait scheduler            :237-252  period 1.4 s, swing height 0.07 m, six phases alternating 0 / 0.5

Both variants run inside the generic step kernel (``ElgStepParams.gait_2_step_hexapod`` / ``terminate_upside_down``); the
``AsyncGaitScheduler`` reward is commented out in the reference's configs and not implemented.
"""
from types import SimpleNamespace

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
