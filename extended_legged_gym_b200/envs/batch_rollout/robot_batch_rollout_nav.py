"""``RobotBatchRolloutNav`` -- navigation task on the main/rollout layout, host side.

Mirrors envs/batch_rollout/robot_batch_rollout_nav.py of the reference (:12-290): fixed start poses and goal positions per
main env (``cfg.navi_opt``), velocity commands computed from the goal every step for every env (mains and rollouts), goal
detection.  The reference evaluates this in ``_post_physics_step_callback`` / ``_post_physics_step_callback_rollout`` with two
Python loops over ``total_num_envs`` (:144-147, :235-238) and ~40 ATen launches; here it is ONE launch of
``elg_nav_commands`` issued from the pre-step hook, i.e. after the command resampling and before the fused step kernel, so
that this step's rewards and observations see the navigation commands exactly as in the reference.

The reference class sits on ``RobotBatchRolloutPercept`` (ray-cast / SDF observations); those sensors are separate classes
here (``utils/ray_caster.py``, ``utils/mesh_sdf.py``) and are not part of this class.

Differences a user should know: the kernel's heading update is switched off for this task (``heading_command`` only matters
to the reference's callback, whose result the navigation update overwrites in the same call), and resets go through the
host-driven path (``fused_reset = False``) because they also restore the start pose.
"""
import ctypes as C

import torch

from ... import _lib
from .robot_batch_rollout import RobotBatchRollout


class RobotBatchRolloutNav(RobotBatchRollout):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        self.goal_reached = None
        self.prev_commands = None
        self.cfg = cfg
        self.num_main_envs = cfg.env.num_envs
        self.device = sim_device if isinstance(sim_device, str) else str(sim_device)
        self._process_start_goal_config()
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.fused_reset = False
        n = self.total_num_envs
        self._goal_reached_u8 = torch.zeros(n, dtype=torch.bool, device=self.device)
        self._prev_commands = torch.zeros(n, 3, device=self.device)
        self._distance = torch.zeros(n, device=self.device)
        # env origins follow the start position of the main env (:72-84)
        self.env_origins[:] = self.start_positions.repeat_interleave(1 + self.num_rollout_per_main, dim=0)

    def _process_start_goal_config(self):
        """:41-70 -- a single pose / goal is repeated, a list is truncated or padded with its last entry"""
        def expand(v, m):
            rows = [v] * m if isinstance(v[0], (int, float)) else (v[:m] if m <= len(v) else v + [v[-1]] * (m - len(v)))
            return torch.tensor(rows, device=self.device, dtype=torch.float)
        opt, m = self.cfg.navi_opt, self.num_main_envs
        self.start_positions = expand(opt.start_pos, m)
        self.start_orientations = expand(opt.start_quat, m)
        self.goal_positions = expand(opt.goal_pos, m).contiguous()

    def _native_params(self):
        p = super()._native_params()
        p.heading_command = 0
        return p

    # ------------------------------------------------------------------------------------------
    # reset (:86-112)
    # ------------------------------------------------------------------------------------------
    def reset_idx(self, env_ids):
        super().reset_idx(env_ids)
        if len(env_ids) == 0:
            return
        main = torch.div(env_ids, 1 + self.num_rollout_per_main, rounding_mode="floor")
        self.root_states[env_ids, 0:3] = self.start_positions[main]
        self.root_states[env_ids, 3:7] = self.start_orientations[main]
        self.root_states[env_ids, 7:13] = 0.0
        if self.goal_reached is None:
            self.goal_reached = self._goal_reached_u8
        self.goal_reached[env_ids] = False
        if self.prev_commands is None:
            self.prev_commands = self._prev_commands
        self.prev_commands[env_ids] = 0.0

    # ------------------------------------------------------------------------------------------
    # navigation commands + goal detection (:114-247)
    # ------------------------------------------------------------------------------------------
    def _nav_params(self):
        opt, p = self.cfg.navi_opt, _lib.ElgNavParams()
        p.use_2d_nav = int(bool(opt.use_2d_nav))
        p.num_commands = self.cfg.commands.num_commands
        p.kp_linear, p.kp_angular = opt.kp_linear, opt.kp_angular
        p.max_linear_vel, p.max_angular_vel = opt.max_linear_vel, opt.max_angular_vel
        p.smooth, p.smooth_c = opt.cmd_smooth_factor, 1 - opt.cmd_smooth_factor
        p.tolerance_rad = opt.tolerance_rad
        return p

    def _update_navigation_commands(self):
        """_update_navigation_commands followed by _check_goal_reached (the reference always calls them as a pair)"""
        p = self._nav_params()
        p.use_prev = int(self.prev_commands is not None)
        p.zero_reached = int(self.goal_reached is not None)
        if not self.commands.is_contiguous():
            raise _lib.ElgError("navigation commands: 'commands' must be contiguous")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_nav_commands(self.num_main_envs, self.num_rollout_per_main, C.byref(p), self.root_states.data_ptr(),
                                        self.goal_positions.data_ptr(), self.commands.data_ptr(), self._prev_commands.data_ptr(),
                                        self._goal_reached_u8.data_ptr(), self._distance.data_ptr(), stream)
        _lib.check(rc, "elg_nav_commands")
        self.prev_commands = self._prev_commands
        self.goal_reached = self._goal_reached_u8

    def _check_goal_reached(self):
        """part of the same launch (see _update_navigation_commands); kept for API parity"""

    def _pre_step_hook(self):
        self._update_navigation_commands()
        self._check_goal_reached()

    def _pre_step_hook_rollout(self):
        self._update_navigation_commands()
        self._check_goal_reached()

    def get_goal_reached_status(self, main_env_only=True):
        if self.goal_reached is None:
            return torch.zeros(self.total_num_envs, dtype=torch.bool, device=self.device)
        return self.goal_reached[self.main_env_indices] if main_env_only else self.goal_reached.clone()

    def get_distance_to_goal(self, main_env_only=True):
        goal = self.goal_positions.repeat_interleave(1 + self.num_rollout_per_main, dim=0)
        pos = self.root_states[:, 0:3]
        d = torch.norm(goal[:, 0:2] - pos[:, 0:2], dim=1) if self.cfg.navi_opt.use_2d_nav else torch.norm(goal - pos, dim=1)
        return d[self.main_env_indices] if main_env_only else d.clone()
