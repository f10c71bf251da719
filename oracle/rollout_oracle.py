"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatement of the main/rollout state clone of the reference
(/root/reference/legged_gym/legged_gym/envs/batch_rollout/robot_batch_rollout.py):

  init_env_indices        :119-164   main env k at row k (1 + R), its rollouts behind it
  sync_main_to_rollout    :1447-1535 14 gather -> scatter pairs (+ optional position drift :1493-1497)
  cache_main_env_states   :1537-1583
  restore_main_env_states :1585-1640

``tests/test_rollout_oracle.py`` pins it against the unmodified reference methods bound to a synthetic ``self``
(container only) and against ``tests/golden/rollout_clone.npz`` (generated from the reference).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
from types import SimpleNamespace

import torch

SYNC_ORDER = ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_pos",
              "base_quat", "base_lin_vel", "base_ang_vel", "projected_gravity")          # then drift, then the feet state
SYNC_FEET = ("feet_air_time", "feet_contact_time", "last_contacts")
CACHE_KEYS = ("root_states", "dof_pos", "dof_vel", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_pos",
              "base_quat", "base_lin_vel", "base_ang_vel", "base_lin_acc", "base_ang_acc", "projected_gravity", "feet_air_time",
              "feet_contact_time", "last_contacts")


def make_rollout_state(num_main, rollouts, num_dof=12, num_feet=4, seed=0):
    """Synthetic per-env tensors with the aliasing of the reference (_init_buffers, legged_robot.py:575-580)."""
    g = torch.Generator().manual_seed(seed)
    n = num_main * (1 + rollouts)
    o = SimpleNamespace()
    o.num_main_envs, o.num_rollout_per_main, o.total_num_envs, o.device = num_main, rollouts, n, "cpu"
    o.root_states = torch.randn(n, 13, generator=g)
    o.dof_state = torch.randn(n * num_dof, 2, generator=g)
    o.dof_pos = o.dof_state.view(n, num_dof, 2)[..., 0]
    o.dof_vel = o.dof_state.view(n, num_dof, 2)[..., 1]
    o.base_pos = o.root_states[:, :3]
    o.base_quat = o.root_states[:, 3:7]
    for name, w in (("actions", num_dof), ("last_actions", num_dof), ("last_dof_vel", num_dof), ("last_root_vel", 6),
                    ("base_lin_vel", 3), ("base_ang_vel", 3), ("base_lin_acc", 3), ("base_ang_acc", 3), ("projected_gravity", 3),
                    ("feet_air_time", num_feet), ("feet_contact_time", num_feet)):
        setattr(o, name, torch.randn(n, w, generator=g))
    o.last_contacts = torch.rand(n, num_feet, generator=g) > 0.5
    return o


def init_env_indices(o):
    R, total = o.num_rollout_per_main, o.total_num_envs
    o.main_env_indices = torch.arange(0, total, 1 + R)
    o.rollout_to_main_map = torch.zeros(total, dtype=torch.long)
    o.is_main_env = torch.zeros(total, dtype=torch.bool)
    o.is_main_env[o.main_env_indices] = True
    o.is_rollout_env = ~o.is_main_env
    for i in range(o.num_main_envs):
        m = int(o.main_env_indices[i])
        o.rollout_to_main_map[m:min(m + 1 + R, total)] = m
    o.rollout_env_indices = torch.nonzero(o.is_rollout_env).flatten()
    o.main_to_rollout_indices = []
    for i in range(o.num_main_envs):
        m = o.main_env_indices[i]
        idx = torch.nonzero(o.rollout_to_main_map == m).flatten()
        o.main_to_rollout_indices.append(idx[idx != m])


def sync_main_to_rollout(o, drift=0.0, drift_u=None):
    """drift_u: the [num_rollout, 3] sample torch.rand_like(base_pos[rollout]) would return (None: draw it)."""
    ri = o.rollout_env_indices
    if len(ri) == 0:
        return
    src = o.rollout_to_main_map[ri]
    for name in SYNC_ORDER:
        t = getattr(o, name)
        t[ri] = t[src]
    if drift > 0.0:
        u = torch.rand_like(o.base_pos[ri]) if drift_u is None else drift_u
        o.base_pos[ri] += (u - 0.5) * drift
    for name in SYNC_FEET:
        t = getattr(o, name)
        t[ri] = t[src]


def cache_main_env_states(o):
    mi = o.main_env_indices
    o.main_env_cache = {k: getattr(o, k)[mi].clone() for k in CACHE_KEYS}


def restore_main_env_states(o):
    mi = o.main_env_indices
    for k in CACHE_KEYS:
        getattr(o, k)[mi] = o.main_env_cache[k].clone()


STATE_KEYS = ("root_states", "dof_state", "actions", "last_actions", "last_dof_vel", "last_root_vel", "base_lin_vel", "base_ang_vel",
              "base_lin_acc", "base_ang_acc", "projected_gravity", "feet_air_time", "feet_contact_time", "last_contacts")


def snapshot(o):
    return {k: getattr(o, k).clone() for k in STATE_KEYS}


# =====================================================================================================================
# the main / rollout variant of the per-step path: RobotBatchRollout.post_physics_step and what it calls
# (envs/batch_rollout/robot_batch_rollout.py:718-761, :819-850, :857-866, :876-940, :1366-1413, :1644-1651)
# =====================================================================================================================
from .legged_oracle import LeggedOracle  # noqa: E402


class BatchRolloutOracle(LeggedOracle):
    """``LeggedOracle`` with the deltas of the reference's ``RobotBatchRollout``: rows are ``num_main`` groups of
    ``1 + rollouts`` envs (main first).  Commands are resampled for main envs and copied to their rollouts; time-outs reset
    main rows only; pushes hit main rows only; ``reset_idx`` runs the terrain curriculum / command resampling / extras for
    the main envs among the reset rows, drops a reset robot onto the terrain surface, and clears episode sums only when a
    main env resets.  Multi-stage reward scales: the rollout class's ``_parse_cfg`` takes the STAGE-0 scales whatever
    ``reward_min_stage`` says (robot_batch_rollout.py:1657-1659: ``_get_reward_scales()`` with its default argument).  ``tests/test_rollout_step.py`` pins it to the unmodified reference methods (container) and to
    ``tests/golden/rollout_step.npz``."""

    INITIAL_SCALES_STAGE = 0

    def __init__(self, cfg, spec, state, height_samples, num_main, rollouts, **kw):
        super().__init__(cfg, spec, state, height_samples, **kw)
        import numpy as np
        self.num_main_envs, self.num_rollout_per_main, self.total_num_envs = num_main, rollouts, self.num_envs
        assert self.num_envs == num_main * (1 + rollouts)
        init_env_indices(self)
        # _parse_cfg (:1644-1651): episode length in whole steps, integer push interval
        self.max_episode_length_s = self.max_episode_length * self.dt
        self.push_interval = int(cfg.domain_rand.push_interval_s / self.dt)
        self.stand_still_threshold = 0.1      # robot_batch_rollout_rew_mixin.py:152 (literal instead of speed_min)
        self.reset_z_from_terrain = cfg.terrain.mesh_type in ("heightfield", "trimesh")

    def _propagate(self, main_ids):
        for m in main_ids.tolist():           # (:827-838)
            self.commands[m + 1:m + 1 + self.num_rollout_per_main] = self.commands[m].clone()

    def callback(self):
        ids = (self.episode_length_buf % int(self.cfg.commands.resampling_time / self.dt) == 0).nonzero(as_tuple=False).flatten()
        main_ids = ids[torch.isin(ids, self.main_env_indices)]
        if len(main_ids) > 0:
            self.resample_commands(main_ids)
            self._propagate(main_ids)
        if self.cfg.commands.heading_command:
            self.heading_command()
        if self.measure_heights:
            self.measured_heights = self.get_heights()
        if self.cfg.domain_rand.push_robots and (self.common_step_counter % self.push_interval == 0):
            mv = self.cfg.domain_rand.max_push_vel_xy      # (:1406-1413) main envs only
            self.root_states[self.main_env_indices, 7:9] = self.rand(-mv, mv, (self.num_main_envs, 2))

    def check_termination(self):
        f = self.contact_forces[:, self.termination_contact_indices, :]
        self.reset_buf = torch.any(torch.norm(f, dim=-1) > 1.0, dim=1)
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf[self.main_env_indices] |= self.time_out_buf[self.main_env_indices]

    def reset_idx(self, env_ids):
        if len(env_ids) == 0:
            return
        main_ids = env_ids[torch.isin(env_ids, self.main_env_indices)]
        if self.curriculum and len(main_ids) > 0:
            self._terrain_curriculum(main_ids)
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self._command_curriculum(main_ids)
        n = len(env_ids)
        self.dof_pos[env_ids] = self.default_dof_pos * self.rand(0.5, 1.5, (n, self.num_dof))
        self.dof_vel[env_ids] = 0.0
        self.root_states[env_ids] = self.base_init_state
        self.root_states[env_ids, :3] += self.env_origins[env_ids]
        if self.custom_origins:               # (:1369-1390)
            self.root_states[env_ids, :2] += self.rand(-0.5, 0.5, (n, 2))
            if self.reset_z_from_terrain:
                points = self.root_states[env_ids, :2].clone().unsqueeze(1)
                points += self.cfg.terrain.border_size
                points = (points / self.cfg.terrain.horizontal_scale).long()
                px = torch.clip(points[:, :, 0].view(-1), 0, self.height_samples.shape[0] - 2)
                py = torch.clip(points[:, :, 1].view(-1), 0, self.height_samples.shape[1] - 2)
                self.root_states[env_ids, 2] = self.height_samples[px, py] * self.cfg.terrain.vertical_scale + self.base_init_state[2]
        self.root_states[env_ids, 7:13] = self.rand(-0.5, 0.5, (n, 6))
        if len(main_ids) > 0:
            self.resample_commands(main_ids)
            self._propagate(main_ids)
        self.last_actions[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.feet_contact_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        if len(main_ids) > 0:
            self.extras["episode"] = {}
            for key in self.episode_sums.keys():
                self.extras["episode"]["rew_" + key] = torch.mean(self.episode_sums[key][main_ids]) / self.max_episode_length_s
                self.episode_sums[key][env_ids] = 0.0
            if self.curriculum:
                self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
            if self.cfg.commands.curriculum:
                self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
            if self.cfg.rewards.multi_stage_rewards:
                self.extras["episode"]["reward_stage"] = float(self.cfg.rewards.reward_min_stage)
            if self.cfg.env.send_timeouts:
                self.extras["time_outs"] = self.time_out_buf


class RobotBatchRolloutOracle(BatchRolloutOracle):
    """``BatchRolloutOracle`` with the deltas of the reference's robot-specific main / rollout classes
    (envs/anymal_c/batch_rollout/anymal_c_batch_rollout.py:49-225, envs/go2/batch_rollout/go2_batch_rollout.py:49-230; the hexapod
    variant envs/elspider_air/batch_rollout/elspider_air_batch_rollout.py:176 flags every row):

      check_termination  (:192-199)  upside-down robots (projected_gravity.z > 0) are reset -- ``upside_down_rows`` = "main" (ANYmal,
                                     Go2: main rows only) or "all" (hexapod)
      gait scheduler     (:143-150)  stepped AFTER the env step with the env clock: gait_idx = remainder(t / period, 1) for every
                                     row (utils/gait_scheduler.py:63-72 with ``t`` given), t = t_main before the caller advances it

    Pinned to the unmodified ``AnymalCBatchRollout`` (tag c) and ``ElSpiderAirBatchRollout`` (tag d) methods by
    tests/golden/rollout_step_anymal.npz (tests/golden/make_rollout_step_golden.py --robot)."""

    def __init__(self, cfg, spec, state, height_samples, num_main, rollouts, upside_down_rows="main", gait_period=1.0, **kw):
        """``gait_period`` None: the hexapod class, whose main step does not touch its scheduler (it is advanced by
        post_physics_step_rollout only, elspider_air_batch_rollout.py:132-135)"""
        super().__init__(cfg, spec, state, height_samples, num_main, rollouts, **kw)
        assert upside_down_rows in ("main", "all")
        self.upside_down_rows = upside_down_rows
        self.gait_period = gait_period
        self.t_main = 0.0
        self.gait_idx = torch.zeros(self.num_envs)

    def check_termination(self):
        super().check_termination()
        if self.upside_down_rows == "main":
            self.reset_buf[self.main_env_indices] |= self.projected_gravity[self.main_env_indices, 2] > 0
        else:
            self.reset_buf |= self.projected_gravity[:, 2] > 0

    def post_physics_step(self, noise_u=None, do_reset=True):
        super().post_physics_step(noise_u, do_reset)
        if self.gait_period is not None:
            self.gait_idx = torch.remainder(self.t_main / self.gait_period * torch.ones(self.num_envs, dtype=torch.float), 1.0)

    def post_physics_step_rollout(self, noise_u=None, t_rollout=None, gait_increment=(0.005, 1.4)):
        """the robot classes' post_physics_step_rollout: the base step, then the scheduler -- on the rollout clock (ANYmal / Go2,
        anymal_c_batch_rollout.py:143-146) or by one increment dt / period of ``cfg.gait_scheduler`` (hexapod,
        elspider_air_batch_rollout.py:132-135 -> utils/gait_scheduler.py:68-69)"""
        super().post_physics_step_rollout(noise_u)
        if self.gait_period is not None:
            self.gait_idx = torch.remainder(t_rollout / self.gait_period * torch.ones(self.num_envs, dtype=torch.float), 1.0)
        else:
            self.gait_idx = torch.remainder(self.gait_idx + gait_increment[0] / gait_increment[1], 1.0)
