"""``AnymalCBatchRollout`` -- the ANYmal-C main / rollout task class of the reference
(envs/anymal_c/batch_rollout/anymal_c_batch_rollout.py:49-225 in /root/reference/legged_gym/legged_gym): ``RobotBatchRolloutPercept``
with the robot's own hooks.  Every hook is an existing launch here:

  _compute_torques(actions, env_ids=None)  :175-190  actuator-network torque path over ALL rows (``Anymal``'s launch, one thread per
                                                    (env, dof) row; ``env_ids`` selects rows of the result) or the PD controller
  _init_buffers / reset_idx                :152-173  LSTM state per (row, dof); cleared for the rows that reset (both reset paths)
  check_termination                        :192-199  upside-down MAIN robots (projected_gravity.z > 0) are reset -- inside the step
                                                    kernel (``ElgStepParams.terminate_upside_down = 2``); the rollout-mode step has
                                                    no termination and keeps the lean kernel
  gait scheduler                           :66-81, :143-150  ``cfg.gait_scheduler`` (period 1 s, phases 0 / 0.5 / 0 / 0.5, 0.04 m);
                                                    stepped after every env step WITH THE ENV CLOCK: gait_idx = remainder(t / period, 1)
                                                    on every row (t_main after a main step, t_rollout after a rollout step), feet
                                                    heights kept by the kernel
  _reward_gait_scheduler / _reward_async_gait_scheduler  :208-225  the kernel's foot-height tracking term / the three posture terms
                                                    of ``AsyncGaitScheduler`` as a Python-side term (as in ``ElSpider``)
  _get_noise_scale_vec, _reward_orientation :101-124, :202-206  the base class's expressions

``Go2BatchRollout`` (envs/go2/batch_rollout/go2_batch_rollout.py:49-230) overrides exactly the same hooks in the same way.
"""
from types import SimpleNamespace

import numpy as np
import torch

from ....utils.gait_scheduler import AsyncGaitScheduler, AsyncGaitSchedulerCfg
from ....utils.helpers import class_to_dict
from ...batch_rollout.robot_batch_rollout_percept import RobotBatchRolloutPercept
from ..anymal import Anymal


class GaitClockMixin:
    """The gait scheduler of the robot-specific main / rollout classes: ``cfg.gait_scheduler`` for the phases and the swing height,
    stepped after every env step WITH THE ENV CLOCK -- ``GaitScheduler.step(foot_positions, foot_velocities, commands, t)``
    (utils/gait_scheduler.py:63-72): gait_idx = remainder(t / period, 1) on every row, t_main after a main step, t_rollout after a
    rollout step (anymal_c_batch_rollout.py:143-150, anymal_c_traj_grad_sampling.py:324-331).  The step kernels keep the feet heights
    and evaluate the foot-height tracking term; the phase is one fill behind the launch."""

    def _init_gait_clock(self):
        g = getattr(self.cfg, "gait_scheduler", None)
        if g is not None:
            self.gait_cfg = SimpleNamespace(dt=g.dt, period=g.period, foot_phases=list(g.foot_phases), swing_height=g.swing_height)
        elif getattr(self, "gait_cfg", None) is None:
            return
        if self.gait_idx is None:
            self.gait_idx = torch.zeros(self.num_envs, device=self.device)
            self.gait_prev_foot_z = torch.zeros(self.num_envs, len(self.feet_indices), device=self.device)
        self._params_dirty = True

    def _gait_follow_clock(self, t):
        if self.gait_idx is None:
            return
        if torch.cuda.is_current_stream_capturing():
            # a captured horizon (rollout_batch) would freeze the value of its capture; the graph is only taken while the
            # scheduler's reward is off (rollout_batch below), and the first eager step afterwards sets the phase from the clock again
            return
        self.gait_idx.fill_(self.clock_phase(t, self.gait_cfg.period))

    @staticmethod
    def clock_phase(t, period):
        """``torch.remainder(t / period * ones(float32), 1.0)`` (utils/gait_scheduler.py:66-67) for Python floats t, period: the
        quotient is formed in double, rounded to float32 by the multiplication with the float32 tensor, then reduced"""
        return float(np.remainder(np.float32(t / period) * np.float32(1.0), np.float32(1.0)))

    def post_physics_step(self):
        super().post_physics_step()
        self._gait_follow_clock(self.t_main)

    def post_physics_step_rollout(self, noise_step=0):
        super().post_physics_step_rollout(noise_step=noise_step)
        self._gait_follow_clock(self.t_rollout)

    def rollout_batch(self, all_us, use_graph=None):
        if use_graph is None and "gait_scheduler" in self.reward_scales:
            use_graph = False      # the phase of every horizon step comes from the host clock
        return super().rollout_batch(all_us, use_graph=use_graph)

    _reward_gait_scheduler = Anymal._reward_gait_scheduler      # (stock term: the kernel evaluates it inside step())


class AnymalCBatchRollout(GaitClockMixin, Anymal, RobotBatchRolloutPercept):
    UPSIDE_DOWN_ROWS = 2        # ElgStepParams.terminate_upside_down: main rows of the main / rollout layout only

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self._init_gait_clock()        # (:66-81) the class's own scheduler config instead of Anymal's training values
        a = getattr(self.cfg, "async_gait_scheduler", None)
        if a is not None and not hasattr(a, "dof_align_sets_idx"):
            a = a() if isinstance(a, type) else AsyncGaitSchedulerCfg()
        # (:84-98) like the reference object, the scheduler keeps the foot tensors of construction time (see AsyncGaitScheduler)
        self.async_gait_scheduler = AsyncGaitScheduler(self.height_samples, self.base_quat, self.base_lin_vel, self.base_ang_vel,
                                                       self.projected_gravity, self.dof_pos, self.dof_vel, self.foot_positions.clone(),
                                                       self.foot_velocities.clone(), self.total_num_envs, self.device, a)

    def _native_params(self):
        p = super()._native_params()
        p.terminate_upside_down = self.UPSIDE_DOWN_ROWS
        return p

    def _compute_torques(self, actions, env_ids=None):
        torques = super()._compute_torques(actions)
        return torques if env_ids is None else torques[env_ids]

    def _reward_async_gait_scheduler(self):
        """(:208-221) stage-dependent weights from cfg.rewards.async_gait_scheduler"""
        scales = class_to_dict(self.cfg.rewards.async_gait_scheduler)

        def weight(key, stage):
            v = scales[key]
            return v[min(stage, len(v) - 1)] if isinstance(v, list) else v
        g, st = self.async_gait_scheduler, self.reward_scales_stage
        return g.reward_dof_align() * weight("dof_align", st) + g.reward_dof_nominal_pos() * weight("dof_nominal_pos", st) + \
            g.reward_foot_z_align() * weight("reward_foot_z_align", st)
