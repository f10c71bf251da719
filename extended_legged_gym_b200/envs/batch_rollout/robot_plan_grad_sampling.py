"""``RobotPlanGradSampling`` -- kinematic planning variant of the main/rollout layout, host side.

Mirrors envs/batch_rollout/robot_plan_grad_sampling.py of the reference (:19-560): actions are *state velocities*
``[base_lin_vel(3), base_ang_vel(3), joint_vel(D)]`` which are integrated kinematically instead of being simulated, once per
horizon step for every rollout env.  The reference's ``_integrate_state_velocities`` + ``_sync_integration_to_sim`` pair is
~60 index_put / gather launches per call; here it is ONE launch of ``elg_integrate_state_velocities`` which integrates and
writes the result through to ``root_states`` / ``dof_state`` / ``base_lin_vel`` / ``base_ang_vel``.

``KinematicStateIntegration`` is the reusable part (method names and argument meaning of the reference); the trajectory
optimiser the reference pulls from the external ``traj_sampling`` package is not part of this repository.
"""
import ctypes as C
import math

import numpy as np
import torch

from ... import _lib
from .robot_batch_rollout import RobotBatchRollout

_METHODS = {"euler": 0, "rk4": 1}
INTEGRATION_FIELDS = ("integration_base_pos", "integration_base_quat", "integration_dof_pos", "integration_base_lin_vel",
                      "integration_base_ang_vel", "integration_dof_vel")


class KinematicStateIntegration:
    """needs ``total_num_envs, num_dof, device, root_states [N,13], dof_state [N*D,2], base_lin_vel, base_ang_vel`` on self"""

    def _init_planning_settings(self, planning):
        """:52-61"""
        self.state_vel_dim = 6 + self.num_dof
        self.max_base_lin_vel = float(planning.max_base_lin_vel)
        self.max_base_ang_vel = float(planning.max_base_ang_vel)
        self.max_joint_vel = float(planning.max_joint_vel)
        self.integration_method = planning.integration_method
        self.max_integration_step = float(planning.max_integration_step)
        self.enforce_joint_limits = bool(planning.enforce_joint_limits)
        if self.integration_method not in _METHODS:
            raise ValueError(f"unknown integration_method {self.integration_method!r} (euler | rk4)")

    def _init_planning_buffers(self):
        """:82-101"""
        n, d, dev = self.total_num_envs, self.num_dof, self.device
        self.integration_base_pos = torch.zeros(n, 3, device=dev)
        self.integration_base_quat = torch.zeros(n, 4, device=dev)
        self.integration_dof_pos = torch.zeros(n, d, device=dev)
        self.integration_base_lin_vel = torch.zeros(n, 3, device=dev)
        self.integration_base_ang_vel = torch.zeros(n, 3, device=dev)
        self.integration_dof_vel = torch.zeros(n, d, device=dev)
        if self.enforce_joint_limits and getattr(self, "dof_pos_limits", None) is None:
            asset = getattr(getattr(self, "cfg", None), "asset", None)
            lo = getattr(asset, "dof_pos_limit_lower", [-np.pi] * d)
            hi = getattr(asset, "dof_pos_limit_upper", [np.pi] * d)
            self.dof_pos_limits = torch.tensor(list(zip(lo, hi)), dtype=torch.float, device=dev)

    def _plan_native(self, dt):
        n_steps = int(math.ceil(dt / min(dt, self.max_integration_step)))
        prm = _lib.ElgPlanParams(self.num_dof, _METHODS[self.integration_method], n_steps, int(self.enforce_joint_limits), dt / n_steps,
                                 self.max_base_lin_vel, self.max_base_ang_vel, self.max_joint_vel)
        lim = self.dof_pos_limits if self.enforce_joint_limits else None
        if lim is not None and not (lim.is_contiguous() and lim.dtype == torch.float32):
            raise ValueError("dof_pos_limits must be a contiguous float32 [D, 2] tensor")
        tensors = (self.integration_base_pos, self.integration_base_quat, self.integration_dof_pos, self.integration_base_lin_vel,
                   self.integration_base_ang_vel, self.integration_dof_vel, lim, self.root_states, self.dof_state, self.base_lin_vel,
                   self.base_ang_vel)
        for t in tensors:
            if t is not None and not (t.is_contiguous() and t.dtype == torch.float32 and t.is_cuda):
                raise ValueError("state integration needs contiguous float32 CUDA tensors")
        buf = _lib.ElgPlanBuffers(*[None if t is None else t.data_ptr() for t in tensors])
        return prm, buf

    def _launch_integration(self, state_vels, dt, env_indices):
        prm, buf = self._plan_native(dt)
        if env_indices is None:
            ids, rows = None, self.total_num_envs
        else:
            ids = env_indices.to(device=self.device, dtype=torch.int64).contiguous()
            rows = ids.numel()
        if state_vels is not None:
            state_vels = state_vels.to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(state_vels.shape) != (rows, self.state_vel_dim):
                raise ValueError(f"Expected state_vels shape ({rows}, {self.state_vel_dim}), got {tuple(state_vels.shape)}")
        _lib.check(_lib.load().elg_integrate_state_velocities(C.byref(prm), C.byref(buf), None if state_vels is None else state_vels.data_ptr(),
                                                             None if ids is None else ids.data_ptr(), rows,
                                                             torch.cuda.current_stream().cuda_stream), "elg_integrate_state_velocities")

    def _integrate_state_velocities(self, state_vels, dt, env_indices=None):
        """:103-195, fused with the write-through of :197-225 (the reference always calls the pair back to back)"""
        self._launch_integration(state_vels, dt, env_indices)

    def _sync_integration_to_sim(self, env_indices=None):
        """:197-225 on its own (e.g. after a reset wrote the integration_* state)"""
        self._launch_integration(None, 1.0, env_indices)

    def _sync_sim_to_integration(self, env_indices=None):
        """:227-243 -- reset-time copy, plain tensor indexing"""
        idx = slice(None) if env_indices is None else env_indices
        d = self.dof_state.view(self.total_num_envs, self.num_dof, 2)
        self.integration_base_pos[idx] = self.root_states[idx, :3]
        self.integration_base_quat[idx] = self.root_states[idx, 3:7]
        self.integration_base_lin_vel[idx] = self.root_states[idx, 7:10]
        self.integration_base_ang_vel[idx] = self.root_states[idx, 10:13]
        self.integration_dof_pos[idx] = d[idx, :, 0]
        self.integration_dof_vel[idx] = d[idx, :, 1]


class RobotPlanGradSampling(KinematicStateIntegration, RobotBatchRollout):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.fused_reset = False          # resets also re-seed the integration state (reset_idx below)
        self._init_planning_settings(cfg.planning)
        self._init_planning_buffers()
        self._sync_sim_to_integration()

    def _sync_main_to_rollout(self):
        """:478-498 -- the simulator fields, then the integration state of every main row into its rollout rows (one more launch
        of the clone kernel instead of 6 copies per main env in a Python loop; no position drift on the integration state, as
        in the reference).  The reference also caches / restores the main rows' integration state around a rollout
        (:500-537); rollouts never write those rows here, so there is nothing to put back."""
        super()._sync_main_to_rollout()
        if self.num_rollout_per_main and getattr(self, "integration_base_pos", None) is not None:
            self._clone(_lib.CLONE_SYNC, INTEGRATION_FIELDS)

    def step(self, actions):
        """:245-319 -- main envs: integrate, copy to the rollouts, score with the regular post-physics step"""
        self._integrate_state_velocities(actions, self.dt, self.main_env_indices)
        self._sync_main_to_rollout()
        self.post_physics_step()
        out = self._rows(self.main_env_indices, self.cfg.normalization.clip_observations)
        self._cache_main_env_states()
        self._sync_main_to_rollout()
        self.t_main += self.dt
        self.t_rollout = self.t_main
        return out

    def step_rollout(self, rollout_state_vels, noise_scales=None):
        """:321-394 -- rollout envs: integrate, rollout-mode post-physics step, main rows restored"""
        n_roll = len(self.rollout_env_indices)
        if rollout_state_vels.shape[0] == self.num_main_envs and self.num_main_envs != n_roll:
            mean = rollout_state_vels.to(self.device).repeat_interleave(self.num_rollout_per_main, dim=0)
            state_vels = mean + torch.randn_like(mean) * noise_scales.to(self.device) if noise_scales is not None else mean
        else:
            state_vels = rollout_state_vels
        self._integrate_state_velocities(state_vels, self.dt, self.rollout_env_indices)
        self.post_physics_step_rollout()
        self._restore_main_env_states()
        out = self._rows(self.rollout_env_indices, self.cfg.normalization.clip_observations)
        self.t_rollout += self.dt
        return out

    def reset_idx(self, env_ids):
        """:539-560"""
        super().reset_idx(env_ids)
        if len(env_ids):
            self._sync_sim_to_integration(env_ids)
