"""``RobotBatchRollout`` -- main/rollout env layout for sampling-based MPC, host side.

Same layout and API as the reference class (envs/batch_rollout/robot_batch_rollout.py in
/root/reference/legged_gym/legged_gym): ``num_main_envs = cfg.env.num_envs`` main envs, each followed by
``cfg.env.rollout_envs`` rollout envs (main k at row k (1 + R), :119-164).  What the reference does with a Python
loop over the mains plus 14 gather/scatter pairs per call is ONE launch of ``elg_clone_rows`` here:

  _sync_main_to_rollout()     :1447-1535  -> elg_clone_rows(ELG_CLONE_SYNC)   (+ optional position drift)
  _cache_main_env_states()    :1537-1583  -> elg_clone_rows(ELG_CLONE_CACHE)
  _restore_main_env_states()  :1585-1640  -> elg_clone_rows(ELG_CLONE_RESTORE)
  step(actions)               :535-600    main-env actions in, main-env rows out
  step_rollout(actions)       :602-716    rollout-env actions in, rollout rows out, mains restored
"""
import ctypes as C
import warnings

import numpy as np
import torch

from ... import _lib
from ..base.legged_robot import LeggedRobot

# fields copied main -> rollout (robot_batch_rollout.py:1474-1502).  base_pos / base_quat are views of root_states and
# dof_pos / dof_vel views of dof_state, so the underlying tensors are passed once.
SYNC_FIELDS = ("root_states", "dof_state", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_lin_vel",
               "base_ang_vel", "projected_gravity", "feet_air_time", "feet_contact_time", "last_contacts")
# fields cached / restored for the main rows (:1546-1583): the sync set + the acceleration EMAs
CACHE_FIELDS = SYNC_FIELDS + ("base_lin_acc", "base_ang_acc")


class RobotBatchRollout(LeggedRobot):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        self.num_main_envs = cfg.env.num_envs
        self.num_rollout_per_main = cfg.env.rollout_envs
        self.total_num_envs = self.num_main_envs * (1 + self.num_rollout_per_main)
        self.original_num_envs = cfg.env.num_envs
        # main env k is row k (1 + R): the kernels take the group size (time-outs, command resampling, curriculum and episode
        # statistics belong to main rows -- :819-838, :857-866, :876-940)
        self._rows_per_main = 1 + self.num_rollout_per_main
        # _reset_root_states (:1366-1404) puts a reset robot on the terrain surface below its new xy position
        self._reset_z_from_terrain = cfg.terrain.mesh_type in ("heightfield", "trimesh")
        cfg.env.num_envs = self.total_num_envs          # every per-env tensor covers mains + rollouts (:77-80)
        try:
            super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        finally:
            cfg.env.num_envs = self.original_num_envs
        self._init_env_indices()
        self.t_main = 0.0
        self.t_rollout = 0.0
        self.main_env_cache = None
        self.drift_u = None          # [num_rollout, 3] uniform samples for parity with torch.rand_like, else in-kernel Philox
        self._clone_calls = 0
        self._clone_tables = {}

    def _parse_cfg(self, cfg):
        super()._parse_cfg(cfg)
        # robot_batch_rollout.py:1646-1651: episode length in whole steps, integer push interval
        self.max_episode_length_s = self.max_episode_length * self.dt
        self.cfg.domain_rand.push_interval = int(self.cfg.domain_rand.push_interval_s / self.dt)
        # (:1657-1659) the rollout class starts from the STAGE-0 scales of a multi-stage config (``_get_reward_scales()`` with its
        # default argument) while reward_scales_stage is reward_min_stage -- kept as it is
        self.reward_scales = self._get_reward_scales()

    # ------------------------------------------------------------------------------------------
    # env layout (robot_batch_rollout.py:119-164) -- closed form instead of the reference's loops
    # ------------------------------------------------------------------------------------------
    def _init_env_indices(self):
        dev, R = self.device, self.num_rollout_per_main
        total = self.total_num_envs
        ar = torch.arange(total, device=dev)
        self.main_env_indices = torch.arange(0, total, 1 + R, device=dev)
        self.rollout_to_main_map = (ar // (1 + R)) * (1 + R)
        self.is_main_env = (ar % (1 + R)) == 0
        self.is_rollout_env = ~self.is_main_env
        self.rollout_env_indices = torch.nonzero(self.is_rollout_env).flatten()
        self.main_to_rollout_indices = [torch.arange(int(m) + 1, int(m) + 1 + R, device=dev) for m in self.main_env_indices.tolist()]

    # ------------------------------------------------------------------------------------------
    # clone / cache / restore
    # ------------------------------------------------------------------------------------------
    def _clone_table(self, fields, with_cache):
        tb = _lib.ElgCloneTable()
        tb.num_fields, tb.num_main, tb.rollouts_per_main = len(fields), self.num_main_envs, self.num_rollout_per_main
        tb.drift_field = -1
        keep = []
        for i, name in enumerate(fields):
            t = getattr(self, name)
            if not (t.is_cuda and t.is_contiguous()):
                raise _lib.ElgError(f"clone field '{name}' must be a contiguous CUDA tensor")
            rows = self.total_num_envs
            row_bytes = t.numel() * t.element_size() // rows
            tb.fields[i].base = t.data_ptr()
            tb.fields[i].row_bytes = row_bytes
            if with_cache:
                c = self.main_env_cache[name]
                tb.fields[i].cache = c.data_ptr()
                keep.append(c)
            if name == "root_states":
                tb.drift_field = i
            keep.append(t)
        return tb, keep

    def _clone(self, mode, fields, drift=0.0):
        # the table only holds pointers and sizes: rebuilt when a tensor was rebound, otherwise reused (the ctypes fill
        # costs more host time than the kernel takes)
        key = tuple(getattr(self, f).data_ptr() for f in fields)
        cached = self._clone_tables.get((mode, fields))
        if cached is None or cached[0] != key:
            cached = (key,) + self._clone_table(fields, with_cache=mode != _lib.CLONE_SYNC)
            self._clone_tables[(mode, fields)] = cached
        tb, keep = cached[1], cached[2]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        du = self.drift_u
        if du is not None and not (du.is_cuda and du.is_contiguous() and du.dtype == torch.float):
            du = du.to(self.device, torch.float).contiguous()
        rc = self._lib.elg_clone_rows(C.byref(tb), mode, float(drift), _lib.ptr(du), self.noise_seed, self._clone_calls, stream)
        _lib.check(rc, "elg_clone_rows")
        self._clone_calls += 1

    def _sync_main_to_rollout(self):
        if self.num_rollout_per_main == 0:
            return
        self._clone(_lib.CLONE_SYNC, SYNC_FIELDS, drift=getattr(self.cfg.domain_rand, "rollout_envs_sync_pos_drift", 0.0))
        self.sim.set_dof_state()
        self.sim.set_root_state()
        self.t_rollout = self.t_main

    def _cache_main_env_states(self):
        if self.main_env_cache is None:
            self.main_env_cache = {}
            for name in CACHE_FIELDS:
                t = getattr(self, name)
                per_row = t.numel() // self.total_num_envs
                self.main_env_cache[name] = torch.zeros(self.num_main_envs, per_row, dtype=t.dtype, device=self.device)
            # views with the reference's cache keys (:1546-1564)
            c = self.main_env_cache
            c["dof_pos"] = c["dof_state"].view(self.num_main_envs, self.num_dof, 2)[..., 0]
            c["dof_vel"] = c["dof_state"].view(self.num_main_envs, self.num_dof, 2)[..., 1]
            c["base_pos"] = c["root_states"][:, :3]
            c["base_quat"] = c["root_states"][:, 3:7]
        self._clone(_lib.CLONE_CACHE, CACHE_FIELDS)

    def _restore_main_env_states(self):
        if self.main_env_cache is None:
            warnings.warn("Attempted to restore main environment states without cache.")
            return
        self._clone(_lib.CLONE_RESTORE, CACHE_FIELDS)
        self.sim.set_dof_state()
        self.sim.set_root_state()

    # ------------------------------------------------------------------------------------------
    # stepping (robot_batch_rollout.py:535-716)
    # ------------------------------------------------------------------------------------------
    def _rows(self, idx, clip_obs):
        """Rows of the step outputs for the main or the rollout envs.  Main env k is row k (1 + R) and its rollouts sit right
        behind it (_init_env_indices, robot_batch_rollout.py:119-164), so both selections are strided views: the clip
        reads them in place and writes the result -- no index gather of the [rows, num_obs] block."""
        M, R1 = self.num_main_envs, 1 + self.num_rollout_per_main
        if idx is self.main_env_indices:
            pick = lambda t: t.view(M, R1, *t.shape[1:])[:, 0].clone()          # (copies, like the reference's index expressions)
        elif idx is self.rollout_env_indices and self.num_rollout_per_main > 0:
            pick = lambda t: t.view(M, R1, *t.shape[1:])[:, 1:].clone().view(M * (R1 - 1), *t.shape[1:])
        else:
            pick = lambda t: t[idx]
        if idx is self.rollout_env_indices and self.num_rollout_per_main > 0:
            obs = torch.clip(self.obs_buf.view(M, R1, -1)[:, 1:], -clip_obs, clip_obs).view(M * (R1 - 1), -1)
        else:
            obs = torch.clip(pick(self.obs_buf), -clip_obs, clip_obs)
        priv = None
        if self.privileged_obs_buf is not None:
            priv = torch.clip(pick(self.privileged_obs_buf), -clip_obs, clip_obs)
        extras = {k: (pick(v) if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == self.total_num_envs else v)
                  for k, v in self.extras.items()}
        return obs, priv, pick(self.rew_buf), pick(self.reset_buf), extras

    def step(self, actions):
        """actions: [num_main_envs, num_actions]; returns the main-env rows."""
        full = torch.zeros((self.total_num_envs, self.num_actions), device=self.device)
        full[self.main_env_indices] = actions.to(self.device)
        clip_actions = self.cfg.normalization.clip_actions
        torch.clamp(full, -clip_actions, clip_actions, out=self.actions)
        self._sync_main_to_rollout()
        for _ in range(self.cfg.control.decimation):
            self.torques = self._compute_torques(self.actions).view(self.torques.shape)
            self.sim.set_dof_actuation_force(self.torques)
            self.sim.simulate()
            self.sim.refresh()
        self.post_physics_step()
        out = self._rows(self.main_env_indices, self.cfg.normalization.clip_observations)
        self._cache_main_env_states()
        self._sync_main_to_rollout()
        self.t_main += self.dt
        self.t_rollout = self.t_main
        return out

    def step_rollout(self, rollout_actions, noise_scales=None):
        """rollout_actions: [num_rollout_envs, A] (or the legacy [num_main_envs, A] mean actions with optional
        Gaussian ``noise_scales``, :623-641); returns the rollout-env rows and leaves the main rows restored."""
        n_roll = len(self.rollout_env_indices)
        if rollout_actions.shape[0] == self.num_main_envs and self.num_main_envs != n_roll:
            mean = rollout_actions.to(self.device).repeat_interleave(self.num_rollout_per_main, dim=0)
            actions = mean + torch.randn_like(mean) * noise_scales.to(self.device) if noise_scales is not None else mean
        else:
            actions = rollout_actions
            if actions.shape[0] != n_roll:
                raise ValueError(f"Expected actions shape ({n_roll}, {self.num_actions}), got {actions.shape}")
        clip_actions = self.cfg.normalization.clip_actions
        self.actions[self.rollout_env_indices] = torch.clip(actions, -clip_actions, clip_actions).to(self.device)
        for _ in range(self.cfg.control.decimation):
            self.torques = self._compute_torques(self.actions).view(self.torques.shape)
            self.sim.set_dof_actuation_force(self.torques)
            self.sim.simulate()
            self.sim.refresh()
        self.post_physics_step_rollout()
        self._restore_main_env_states()
        out = self._rows(self.rollout_env_indices, self.cfg.normalization.clip_observations)
        self.t_rollout += self.dt
        return out

    # ------------------------------------------------------------------------------------------
    # rollout_batch (robot_traj_grad_sampling.py:249-280): the horizon loop of one MPPI iteration
    # ------------------------------------------------------------------------------------------
    def _rollout_actions(self, rollout_actions):
        """step_rollout's action hand-over (:643-656) as ONE launch: clip + scatter into the rollout rows of ``actions``
        (subclasses add the joint-target denormalisation through ``_action_denorm``)."""
        a = rollout_actions
        # rows may be strided (step i of an [M * R, horizon, A] plan is all_us[:, i]): the kernel takes the row stride, no copy
        if not (a.is_cuda and a.dtype == torch.float and a.dim() == 2 and a.stride(1) == 1 and a.stride(0) >= a.shape[1]):
            a = a.to(self.device, torch.float).contiguous()
        lower, rng = getattr(self, "_action_denorm", (None, None))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.elg_rollout_actions(a.data_ptr(), self.num_main_envs, self.num_rollout_per_main, self.num_actions, int(a.stride(0)),
                                                 float(self.cfg.normalization.clip_actions), _lib.ptr(lower), _lib.ptr(rng),
                                                 self.actions.data_ptr(), stream), "elg_rollout_actions")

    def _rollout_horizon(self, all_us, rewards):
        """sync -> horizon x [actions, decimation x torques, rollout-mode step (its reward column written in place), restore]
        -> sync.  No host synchronisation anywhere: this is what gets captured into a CUDA graph."""
        horizon = all_us.shape[1]
        self._sync_main_to_rollout()
        self._sync_native()
        for i in range(horizon):
            self._rollout_actions(all_us[:, i])
            for _ in range(self.cfg.control.decimation):
                self.torques = self._compute_torques(self.actions).view(self.torques.shape)
                self.sim.set_dof_actuation_force(self.torques)
                self.sim.simulate()
                self.sim.refresh()
            self._bufs.rollout_rew_out = rewards.data_ptr() + 4 * i
            self._params.rollout_rew_stride = horizon
            try:
                self.post_physics_step_rollout(noise_step=i)
            finally:
                self._bufs.rollout_rew_out = None
            self._restore_main_env_states()
            self.t_rollout += self.dt
        self._sync_main_to_rollout()

    def rollout_batch(self, all_us, use_graph=None):
        """all_us [num_rollout_envs, horizon, A] -> rewards [num_rollout_envs, horizon]; rollouts are re-synchronised with
        their mains before and after (RobotTrajGradSampling.rollout_batch).  With a capturable simulator backend (the
        synthetic one; PhysX is not) the whole loop is ONE CUDA graph, captured at the first call for a given shape and
        replayed afterwards: ``all_us`` is copied into the graph's input buffer, the returned table is the graph's output
        buffer (valid until the next call).  In-kernel noise stays fresh across replays through a device-side step counter."""
        if all_us.dim() != 3 or all_us.shape[0] != len(self.rollout_env_indices) or all_us.shape[2] != self.num_actions:
            raise ValueError(f"Expected all_us of shape ({len(self.rollout_env_indices)}, horizon, {self.num_actions}), got {tuple(all_us.shape)}")
        if self.main_env_cache is None:
            self._cache_main_env_states()
        if use_graph is None:
            use_graph = bool(getattr(self.sim, "capturable", False)) and not self._python_terms and \
                getattr(self.cfg.domain_rand, "rollout_envs_sync_pos_drift", 0.0) <= 0.0
        horizon = all_us.shape[1]
        if not use_graph:
            rewards = torch.zeros((all_us.shape[0], horizon), device=self.device)
            self._rollout_horizon(all_us.to(self.device, torch.float).contiguous(), rewards)
            return rewards
        key = (tuple(all_us.shape), self.root_states.data_ptr(), self.actions.data_ptr(), self.obs_buf.data_ptr())
        g = getattr(self, "_rollout_graph", None)
        if g is None or g["key"] != key:
            g = {"key": key, "us": torch.zeros(all_us.shape, device=self.device), "rew": torch.zeros((all_us.shape[0], horizon), device=self.device)}
            if getattr(self, "_step_counter", None) is None:
                self._step_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
                object.__setattr__(self, "_ptrs_dirty", True)       # the native buffer struct picks the counter up
            g["us"].copy_(all_us)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._rollout_horizon(g["us"], g["rew"])        # eager once: lazily built tables exist before the capture
                side.synchronize()
                g["graph"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g["graph"], stream=side):
                    self._step_counter.add_(horizon)
                    self._rollout_horizon(g["us"], g["rew"])
            cur.wait_stream(side)
            self._rollout_graph = g
        g["us"].copy_(all_us)
        g["graph"].replay()
        return g["rew"]

    def post_physics_step_rollout(self, noise_step=0):
        """robot_batch_rollout.py:763-817: derive + rewards + observations + histories, no episode counter, no
        termination / reset, no command / height refresh (``_post_physics_step_callback_rollout`` is empty).
        The kernel runs over every row; main rows are put back by ``_restore_main_env_states`` right after."""
        self.sim.refresh()
        P = _lib
        self._pre_step_hook_rollout()
        if self._python_terms:
            # subclass terms written in Python read the derived state: derive first, evaluate them with torch (no episode sums in
            # the rollout step: compute_reward_rollout :969-985), then the registry / observations / histories
            self._launch(P.PHASE_DERIVE, rollout=True, noise_step=noise_step)
            rew_out = self._bufs.rollout_rew_out          # (rollout_batch points this at its reward column; a struct rebuild drops it)
            if self.extra_reward is None or self.extra_reward.shape[0] != self.num_envs:
                self.extra_reward = torch.zeros(self.num_envs, device=self.device)
            extra = self.extra_reward                     # accumulated in place: the native struct keeps its pointer
            extra.zero_()
            for name in self._python_terms:
                extra += getattr(self, "_reward_" + name)() * self.reward_scales[name]
            self._sync_native()
            self._bufs.rollout_rew_out = rew_out
            self._launch(P.PHASE_REWARD | P.PHASE_OBS | P.PHASE_HISTORY, rollout=True, noise_step=noise_step)
            return
        self._launch(P.PHASE_DERIVE | P.PHASE_REWARD | P.PHASE_OBS | P.PHASE_HISTORY, rollout=True, noise_step=noise_step)

    def _pre_step_hook_rollout(self):
        """``_post_physics_step_callback_rollout`` (robot_batch_rollout.py:868: empty in the base class)"""

    def check_termination(self):
        """:857-866 -- contact termination everywhere, time-outs only OR-ed into the main rows (the kernel section knows
        the layout through ``ElgStepParams.rows_per_main``)."""
        super().check_termination()

    # ------------------------------------------------------------------------------------------
    # sparse RNG-driven paths of the main / rollout layout, host-driven form (the fused kernels follow the same rules through
    # ElgResetParams.rows_per_main / root_z_from_terrain)
    # ------------------------------------------------------------------------------------------
    def _group_view(self, t):
        return t.view(self.num_main_envs, 1 + self.num_rollout_per_main, *t.shape[1:])

    def _propagate_commands(self, main_env_ids):
        """:827-838 / :903-915 -- every rollout row takes the command row of its main env."""
        if len(main_env_ids) == 0 or self.num_rollout_per_main == 0:
            return
        k = torch.div(main_env_ids, 1 + self.num_rollout_per_main, rounding_mode="floor")
        g = self._group_view(self.commands)
        g[k, 1:] = g[k, :1]

    def _post_physics_step_callback(self):
        """:819-850 -- commands are resampled for MAIN envs only and copied to their rollouts; pushes hit main rows only."""
        interval = int(self.cfg.commands.resampling_time / self.dt)
        env_ids = ((self.episode_length_buf + 1) % interval == 0).nonzero(as_tuple=False).flatten()
        main_env_ids = env_ids[self.is_main_env[env_ids]]
        if len(main_env_ids) > 0:
            self._resample_commands(main_env_ids)
            self._propagate_commands(main_env_ids)
        dr = self.cfg.domain_rand
        return bool(dr.push_robots and (self.common_step_counter % dr.push_interval == 0))

    def _push_robots(self):
        """:1406-1413 -- only main envs are pushed."""
        mv = self.cfg.domain_rand.max_push_vel_xy
        self.root_states[self.main_env_indices, 7:9] = self._rand(-mv, mv, (self.num_main_envs, 2))
        self.sim.set_root_state()

    def _reset_root_states(self, env_ids):
        """:1366-1404 -- like the base class, plus: with custom origins on a heightfield / trimesh terrain the base height is the
        terrain height under the new xy position (one cell, no min-of-3) + the nominal height."""
        self.root_states[env_ids] = self.base_init_state
        self.root_states[env_ids, :3] += self.env_origins[env_ids]
        if self.custom_origins:
            self.root_states[env_ids, :2] += self._rand(-0.5, 0.5, (len(env_ids), 2))
            if self._reset_z_from_terrain:
                points = self.root_states[env_ids, :2].clone().unsqueeze(1)
                points += self.cfg.terrain.border_size
                points = (points / self.cfg.terrain.horizontal_scale).long()
                px = torch.clip(points[:, :, 0].view(-1), 0, self.height_samples.shape[0] - 2)
                py = torch.clip(points[:, :, 1].view(-1), 0, self.height_samples.shape[1] - 2)
                heights = self.height_samples[px, py] * self.cfg.terrain.vertical_scale
                self.root_states[env_ids, 2] = heights + self.base_init_state[2]
        self.root_states[env_ids, 7:13] = self._rand(-0.5, 0.5, (len(env_ids), 6))
        self.sim.set_root_state_indexed(env_ids.to(dtype=torch.int32))

    def _stats_rows(self, env_ids):
        return env_ids[self.is_main_env[env_ids]]

    def reset_idx(self, env_ids):
        """:876-940 -- joints, root and histories of every listed row; terrain curriculum, command resampling (+ copy to the
        rollouts) and the extras["episode"] means for the MAIN envs among them; the episode sums of the listed rows are
        cleared only when a main env is among them (the reference's statement order)."""
        if len(env_ids) == 0:
            return
        main_env_ids = env_ids[self.is_main_env[env_ids]]
        if self.cfg.terrain.curriculum and len(main_env_ids) > 0:
            self._update_terrain_curriculum(main_env_ids)
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self.update_command_curriculum(main_env_ids)
        self._reset_dofs(env_ids)
        self._reset_root_states(env_ids)
        if len(main_env_ids) > 0:
            self._resample_commands(main_env_ids)
            self._propagate_commands(main_env_ids)
        self.last_actions[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.feet_contact_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        if len(main_env_ids) > 0:
            if getattr(self, "episode_stats", None) is not None:
                self.episode_stats.accumulate(self._episode_sums_all, main_env_ids)
            self.extras["episode"] = {}
            for key in self.episode_sums.keys():
                self.extras["episode"]["rew_" + key] = torch.mean(self.episode_sums[key][main_env_ids]) / self.max_episode_length_s
                self.episode_sums[key][env_ids] = 0.0
            if self.cfg.terrain.curriculum:
                self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
            if self.cfg.commands.curriculum:
                self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
            if self.cfg.rewards.multi_stage_rewards:
                self.extras["episode"]["reward_stage"] = float(self.reward_scales_stage)
            if self.cfg.env.send_timeouts:
                self.extras["time_outs"] = self.time_out_buf

    def set_commands(self, main_env_idx, commands):
        lo = int(main_env_idx) * (1 + self.num_rollout_per_main)
        self.commands[lo:lo + 1 + self.num_rollout_per_main] = commands.to(self.device)

    def set_all_commands(self, commands):
        self.commands[:] = commands.to(self.device).repeat_interleave(1 + self.num_rollout_per_main, dim=0)

    def get_observations(self):
        return self.obs_buf[self.main_env_indices]

    def get_observations_rollout(self):
        return self.obs_buf[self.rollout_env_indices]

    def get_observations_all(self):
        return self.obs_buf
