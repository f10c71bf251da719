"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatement of the reference's per-step hot path, i.e. what
``LeggedRobot.step`` does around PhysX and everything ``post_physics_step`` does to the
state tensors PhysX hands back.  Each function cites the reference lines it restates
(paths relative to /root/reference/legged_gym/legged_gym/).  The restatement uses the same
torch ops in the same order as the reference so that on the same CPU it is bit-identical;
``tests/test_oracle_pinned.py`` pins it against the *unmodified* reference imported through
``oracle/ref_harness.py`` (container only) and against the committed fixtures under
``tests/golden/`` (which were generated from the reference, see ``tests/golden/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module.
"""
import math
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import torch_utils as tu

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# small helpers (utils/math_utils.py:40-58, utils/helpers.py:43-58)
# ----------------------------------------------------------------------------------------------
def quat_apply_yaw(quat: Tensor, vec: Tensor) -> Tensor:
    q = quat.clone().view(-1, 4)
    q[:, :2] = 0.0
    return tu.quat_apply(tu.normalize(q), vec)


def wrap_to_pi(angles: Tensor) -> Tensor:
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles


def sorted_public_dict(obj):
    """class_to_dict: keys come out of dir() and are therefore alphabetical."""
    if not hasattr(obj, "__dict__"):
        return obj
    res = {}
    for k in dir(obj):
        if k.startswith("_"):
            continue
        v = getattr(obj, k)
        res[k] = [sorted_public_dict(i) for i in v] if isinstance(v, list) else sorted_public_dict(v)
    return res


def default_rand(lower, upper, shape) -> Tensor:
    """torch_rand_float on the global CPU generator (isaacgym.torch_utils)."""
    return (upper - lower) * torch.rand(*shape) + lower


# ----------------------------------------------------------------------------------------------
# reward registry (envs/base/legged_robot_rew_mixin.py:41-234 + anymal.py:112-114)
# every entry: f(o) -> Tensor[N]
# ----------------------------------------------------------------------------------------------
def _feet_force(o):
    return o.contact_forces[:, o.feet_indices, :]


def _r_lin_vel_z(o):
    return torch.square(o.base_lin_vel[:, 2])


def _r_ang_vel_xy(o):
    return torch.sum(torch.square(o.base_ang_vel[:, :2]), dim=1)


def _r_orientation(o):
    return torch.sum(torch.square(o.projected_gravity[:, :2]), dim=1)


def _r_base_height(o):
    h = torch.mean(o.root_states[:, 2].unsqueeze(1) - o.measured_heights, dim=1)
    return torch.square(h - o.cfg.rewards.base_height_target)


def _r_base_foot_height(o):
    touching = o.feet_contact_time > 1e-3
    z = o.foot_positions[:, :, 2]
    ground = torch.nanmean(torch.where(touching, z, torch.nan), dim=1)
    none = torch.isnan(ground)
    ground = torch.where(none, o.root_states[:, 2] - o.cfg.rewards.base_height_target, ground)
    rel = o.root_states[:, 2] - ground
    return torch.square(rel - o.cfg.rewards.base_height_target)


def _r_torques(o):
    return torch.sum(torch.square(o.torques), dim=1)


def _r_dof_vel(o):
    return torch.sum(torch.square(o.dof_vel), dim=1)


def _r_dof_acc(o):
    return torch.sum(torch.square((o.last_dof_vel - o.dof_vel) / o.dt), dim=1)


def _r_action_rate(o):
    return torch.sum(torch.square(o.last_actions - o.actions), dim=1)


def _r_dof_pos_limits(o):
    out = -(o.dof_pos - o.dof_pos_limits[:, 0]).clip(max=0.0)
    out += (o.dof_pos - o.dof_pos_limits[:, 1]).clip(min=0.0)
    return torch.sum(out, dim=1)


def _r_dof_vel_limits(o):
    over = torch.abs(o.dof_vel) - o.dof_vel_limits * o.cfg.rewards.soft_dof_vel_limit
    return torch.sum(over.clip(min=0.0, max=1.0), dim=1)


def _r_torque_limits(o):
    over = torch.abs(o.torques) - o.torque_limits * o.cfg.rewards.soft_torque_limit
    return torch.sum(over.clip(min=0.0), dim=1)


def _r_collision(o):
    f = o.contact_forces[:, o.penalised_contact_indices, :]
    return torch.sum(1.0 * (torch.norm(f, dim=-1) > 0.1), dim=1)


def _stumbling(o):
    f = _feet_force(o)
    return torch.norm(f[:, :, :2], dim=2) > 5 * torch.abs(f[:, :, 2])


def _r_feet_stumble(o):
    return torch.any(_stumbling(o), dim=1)


def _r_feet_stumble_liftup(o):
    return torch.sum(_stumbling(o) * o.foot_velocities[:, :, 2], dim=1)


def _contact_filt(o):
    contact = _feet_force(o)[:, :, 2] > 1.0
    return contact, torch.logical_or(contact, o.last_contacts)


def _r_feet_slip(o):
    _, filt = _contact_filt(o)
    v2 = torch.square(torch.norm(o.foot_velocities[:, :, 0:2], dim=2).view(o.num_envs, -1))
    return torch.sum(filt * v2, dim=1)


def _r_jump_air(o):
    _, filt = _contact_filt(o)
    return torch.clip(torch.sum((~filt) * (o.feet_air_time - 0.5), dim=1) - len(o.feet_indices) / 2, 0.0)


def _r_feet_air_time(o):
    # NB side effects: rebinds last_contacts, advances both timers
    contact, filt = _contact_filt(o)
    o.last_contacts = contact
    first = (o.feet_air_time > 0.0) * filt
    o.feet_air_time += o.dt
    o.feet_contact_time += o.dt
    rew = torch.sum((o.feet_air_time - 0.5) * first, dim=1)
    rew *= torch.norm(o.commands[:, :2], dim=1) > 0.1
    o.feet_air_time *= ~filt
    o.feet_contact_time *= filt
    return rew


def _r_feet_contact_forces(o):
    return torch.sum((torch.norm(_feet_force(o), dim=-1) - o.cfg.rewards.max_contact_force).clip(min=0.0), dim=1)


def _pair_sq(a, b, cap=4):
    return torch.clip(torch.square(a - b), max=cap)


def _r_gait_2_step(o):
    air, con = o.feet_air_time, o.feet_contact_time

    def sync(i, j):
        return _pair_sq(air[:, i], air[:, j]) + _pair_sq(con[:, i], con[:, j])

    def anti(i, j):
        return _pair_sq(air[:, i], con[:, j]) + _pair_sq(con[:, i], air[:, j])

    if o.elspider:
        # ElSpider._reward_gait_2_step (envs/elspider_air/elspider.py:365-408): feet LB LF LM RB RF RM, tripods (0,1,5) / (2,3,4)
        g1 = (sync(0, 1) + sync(0, 5) + sync(1, 5)) / 3
        g2 = (sync(2, 3) + sync(2, 4) + sync(3, 4)) / 3
        a = (anti(0, 2) + anti(0, 3) + anti(0, 4) + anti(1, 2) + anti(1, 3) + anti(1, 4) + anti(5, 2) + anti(5, 3) + anti(5, 4)) / 9
        s = (g1 + g2) / 2
    else:
        s = (sync(0, 3) + sync(1, 2)) / 2
        a = (anti(0, 1) + anti(0, 2) + anti(3, 2) + anti(3, 1)) / 4
    re = s + a
    k = 3 if o.cfg.commands.heading_command else 2
    moving = torch.logical_or(torch.norm(o.commands[:, :2], dim=1) > o.speed_min,
                              torch.abs(o.commands[:, k]) >= o.speed_min / 2)
    return re * moving


def _r_four_footup(o):
    up = torch.ones(o.num_envs, dtype=torch.float) * 0.1
    o.all_feet_up = torch.all(_feet_force(o)[:, :, 2] < 1, dim=1)
    return up * o.all_feet_up


def _r_termination(o):
    return o.reset_buf * ~o.time_out_buf


def _r_stand_still(o):
    return torch.sum(torch.abs(o.dof_pos - o.default_dof_pos), dim=1) * (torch.norm(o.commands[:, :2], dim=1) < o.speed_min)


def _r_tracking_lin_vel(o):
    err = torch.sum(torch.square(o.commands[:, :2] - o.base_lin_vel[:, :2]), dim=1)
    return torch.exp(-err / o.cfg.rewards.tracking_sigma)


def _r_tracking_ang_vel(o):
    err = torch.square(o.commands[:, 2] - o.base_ang_vel[:, 2])
    return torch.exp(-err / o.cfg.rewards.tracking_sigma)


def _r_gait_scheduler(o):
    # utils/gait_scheduler.py:74-81 via anymal.py:112-114
    rew = torch.zeros(o.num_envs, dtype=torch.float)
    for i, ph in enumerate(o.gait_phases):
        target = torch.where(ph < 0.5, o.gait_cfg.swing_height * torch.sin(2 * torch.pi * ph), torch.zeros_like(ph))
        rew += torch.square(target - o.gait_foot_pos[:, i][:, 2])
    return rew


REWARD_TERMS: Dict[str, Callable] = {
    "lin_vel_z": _r_lin_vel_z, "ang_vel_xy": _r_ang_vel_xy, "orientation": _r_orientation,
    "base_height": _r_base_height, "base_foot_height": _r_base_foot_height, "torques": _r_torques,
    "dof_vel": _r_dof_vel, "dof_acc": _r_dof_acc, "action_rate": _r_action_rate,
    "dof_pos_limits": _r_dof_pos_limits, "dof_vel_limits": _r_dof_vel_limits, "torque_limits": _r_torque_limits,
    "collision": _r_collision, "feet_stumble": _r_feet_stumble, "feet_stumble_liftup": _r_feet_stumble_liftup,
    "feet_slip": _r_feet_slip, "jump_air": _r_jump_air, "feet_air_time": _r_feet_air_time,
    "feet_contact_forces": _r_feet_contact_forces, "gait_2_step": _r_gait_2_step, "four_footup": _r_four_footup,
    "termination": _r_termination, "stand_still": _r_stand_still, "tracking_lin_vel": _r_tracking_lin_vel,
    "tracking_ang_vel": _r_tracking_ang_vel, "gait_scheduler": _r_gait_scheduler,
}


# ----------------------------------------------------------------------------------------------
class LeggedOracle:
    INITIAL_SCALES_STAGE = None      # multi-stage reward scales start at cfg.rewards.reward_min_stage (legged_robot.py:_parse_cfg)

    """State + per-step methods of the reference ``LeggedRobot`` restated on torch-CPU.

    ``state`` uses the PhysX layouts of ``extended_legged_gym_b200.synthetic.make_state``; the
    tensors are used in place (views, like ``_init_buffers`` legged_robot.py:575-584).
    """

    def __init__(self, cfg, spec, state: Dict[str, Tensor], height_samples: Optional[Tensor] = None,
                 use_gait_scheduler: bool = False, rand: Callable = default_rand):
        self.cfg, self.spec, self.rand = cfg, spec, rand
        # the task registry binds the elspider_air configs to the ElSpider class (envs/elspider_air/elspider.py:225-408): hexapod
        # gait_2_step, upside-down termination, 18-DOF noise slices
        self.elspider = getattr(cfg.asset, "name", "") == "elspider_air"
        N = state["root_states"].shape[0]
        D, B = spec.num_dof, spec.num_bodies
        self.num_envs, self.num_dof, self.num_bodies = N, D, B
        self.num_obs, self.num_actions = cfg.env.num_observations, cfg.env.num_actions
        self.sim_dt = cfg.sim.dt
        # _parse_cfg (legged_robot.py:847-860)
        self.dt = cfg.control.decimation * self.sim_dt
        self.obs_scales = cfg.normalization.obs_scales
        self.reward_scales = self._stage_scales(cfg.rewards.reward_min_stage if self.INITIAL_SCALES_STAGE is None else self.INITIAL_SCALES_STAGE)
        self.command_ranges = sorted_public_dict(cfg.commands.ranges)
        self.curriculum = cfg.terrain.curriculum and cfg.terrain.mesh_type in ("heightfield", "trimesh", "confined_trimesh")
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.push_interval = np.ceil(cfg.domain_rand.push_interval_s / self.dt)
        self.speed_min = 0.1
        self.acc_ema = 0.9
        self.common_step_counter = 0
        self.extras = {}
        # indices (legged_robot.py:764-815)
        L = lambda idx: torch.tensor(idx, dtype=torch.long)
        self.feet_indices = L(spec.indices_matching(cfg.asset.foot_name))
        self.penalised_contact_indices = L(spec.indices_matching(cfg.asset.penalize_contacts_on))
        self.termination_contact_indices = L(spec.indices_matching(cfg.asset.terminate_after_contacts_on))
        F = len(self.feet_indices)
        # _process_dof_props (legged_robot.py:357-372)
        self.dof_pos_limits = torch.zeros(D, 2)
        self.dof_vel_limits = torch.zeros(D)
        self.torque_limits = torch.zeros(D)
        for i in range(D):
            self.dof_pos_limits[i, 0] = spec.dof_lower[i]
            self.dof_pos_limits[i, 1] = spec.dof_upper[i]
            self.dof_vel_limits[i] = spec.dof_velocity[i]
            self.torque_limits[i] = spec.dof_effort[i]
            m = (self.dof_pos_limits[i, 0] + self.dof_pos_limits[i, 1]) / 2
            r = self.dof_pos_limits[i, 1] - self.dof_pos_limits[i, 0]
            self.dof_pos_limits[i, 0] = m - 0.5 * r * cfg.rewards.soft_dof_pos_limit
            self.dof_pos_limits[i, 1] = m + 0.5 * r * cfg.rewards.soft_dof_pos_limit
        # PhysX tensors and views (legged_robot.py:575-584)
        self.root_states = state["root_states"]
        self.dof_state = state["dof_state"]
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        self.base_pos = self.root_states[:, :3]
        self.base_quat = self.root_states[:, 3:7]
        self.contact_forces = state["contact_forces"].view(N, -1, 3)
        self.rigid_body_state = state["rigid_body_state"]
        # env-owned buffers (base_task.py:71-79, legged_robot.py:586-647)
        self.obs_buf = torch.zeros(N, self.num_obs)
        self.rew_buf = torch.zeros(N)
        self.reset_buf = torch.zeros(N, dtype=torch.bool)
        self.time_out_buf = torch.zeros(N, dtype=torch.bool)
        self.episode_length_buf = state["episode_length_buf"].clone()
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat((N, 1))
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat((N, 1))
        self.torques = torch.zeros(N, self.num_actions)
        self.actions = state["actions"].clone()
        self.last_actions = state["last_actions"].clone()
        self.last_dof_vel = state["last_dof_vel"].clone()
        self.last_root_vel = state["last_root_vel"].clone()
        self.commands = state["commands"].clone()
        self.commands_scale = torch.tensor([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel])
        self.feet_air_time = state["feet_air_time"].clone()
        self.feet_contact_time = state["feet_contact_time"].clone()
        self.last_contacts = state["last_contacts"].clone()
        self.base_lin_vel = tu.quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = tu.quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = tu.quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.base_lin_acc = state["base_lin_acc"].clone()
        self.base_ang_acc = state["base_ang_acc"].clone()
        self._gather_feet()
        self.measure_heights = cfg.terrain.measure_heights
        self.height_samples = height_samples
        if self.measure_heights:
            self.height_points = self._init_height_points()
        self.measured_heights = 0
        # gains (legged_robot.py:630-647)
        self.p_gains = torch.zeros(self.num_actions)
        self.d_gains = torch.zeros(self.num_actions)
        self.default_dof_pos = torch.zeros(D)
        for i, name in enumerate(spec.dof_names):
            self.default_dof_pos[i] = cfg.init_state.default_joint_angles[name]
            for key in cfg.control.stiffness.keys():
                if key in name:
                    self.p_gains[i] = cfg.control.stiffness[key]
                    self.d_gains[i] = cfg.control.damping[key]
        self.default_dof_pos = self.default_dof_pos.unsqueeze(0)
        self.noise_scale_vec = self._noise_scale_vec()
        self.add_noise = cfg.noise.add_noise
        # terrain bookkeeping used by reset_idx (legged_robot.py:817-844)
        self.base_init_state = torch.tensor(cfg.init_state.pos + cfg.init_state.rot + cfg.init_state.lin_vel + cfg.init_state.ang_vel,
                                            dtype=torch.float)
        self._get_env_origins()
        # gait scheduler (anymal.py:60-79, gait_scheduler.py)
        self.use_gait_scheduler = use_gait_scheduler
        if use_gait_scheduler:
            self.gait_cfg = SimpleNamespace(period=0.6, dt=self.dt, foot_phases=[0.0, 0.5, 0.5, 0.0], swing_height=0.15)
            self.gait_idx = state["gait_idx"].clone()
            self.gait_phases = [torch.remainder(self.gait_idx + p, 1.0) for p in self.gait_cfg.foot_phases]
            self.gait_foot_pos = self.foot_positions
        self.prepare_reward_function()

    # ---- env origins (legged_robot.py:817-844); terrain tile centres follow the synthetic contract of
    # oracle/ref_harness.synthetic_terrain_origins; levels are drawn from a forked, seeded global RNG
    def _get_env_origins(self):
        cfg, N = self.cfg, self.num_envs
        self.env_origins = torch.zeros(N, 3)
        if cfg.terrain.mesh_type in ("heightfield", "trimesh", "confined_trimesh"):
            self.custom_origins = True
            nrow, ncol = cfg.terrain.num_rows, cfg.terrain.num_cols
            to = torch.zeros(nrow, ncol, 3)
            to[..., 0] = (torch.arange(nrow).float().view(-1, 1) + 0.5) * cfg.terrain.terrain_length
            to[..., 1] = (torch.arange(ncol).float().view(1, -1) + 0.5) * cfg.terrain.terrain_width
            max_init = cfg.terrain.max_init_terrain_level if cfg.terrain.curriculum else nrow - 1
            with torch.random.fork_rng():
                torch.manual_seed(1234)
                self.terrain_levels = torch.randint(0, max_init + 1, (N,))
            self.terrain_types = torch.div(torch.arange(N), (N / ncol), rounding_mode="floor").to(torch.long)
            self.max_terrain_level = nrow
            self.terrain_origins = to
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
            self.env_length = cfg.terrain.terrain_length
        else:
            self.custom_origins = False
            num_cols = np.floor(np.sqrt(N))
            num_rows = np.ceil(N / num_cols)
            xx, yy = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
            spacing = cfg.env.env_spacing
            self.env_origins[:, 0] = spacing * xx.flatten()[:N]
            self.env_origins[:, 1] = spacing * yy.flatten()[:N]
            self.env_origins[:, 2] = 0.0

    # ---- reward bookkeeping (legged_robot_rew_mixin.py:15-38, legged_robot.py:649-674) -------
    def _stage_scales(self, stage):
        d = sorted_public_dict(self.cfg.rewards.scales)
        if not self.cfg.rewards.multi_stage_rewards:
            return d
        return {k: (v if not isinstance(v, list) else v[min(stage, len(v) - 1)]) for k, v in d.items()}

    def prepare_reward_function(self):
        for k in list(self.reward_scales.keys()):
            if self.reward_scales[k] == 0:
                self.reward_scales.pop(k)
            else:
                self.reward_scales[k] *= self.dt
        self.reward_names = [n for n in self.reward_scales if n != "termination"]
        self.episode_sums = {n: torch.zeros(self.num_envs) for n in self.reward_scales}

    # ---- init helpers ---------------------------------------------------------------------------
    def _init_height_points(self):
        # legged_robot.py:884-898 ('ij' meshgrid -> index ix*len(y)+iy)
        y = torch.tensor(self.cfg.terrain.measured_points_y)
        x = torch.tensor(self.cfg.terrain.measured_points_x)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = gx.numel()
        pts = torch.zeros(self.num_envs, self.num_height_points, 3)
        pts[:, :, 0] = gx.flatten()
        pts[:, :, 1] = gy.flatten()
        return pts

    def _noise_scale_vec(self):
        # legged_robot.py:533-556
        v = torch.zeros(self.num_obs)
        ns, lvl, os_ = self.cfg.noise.noise_scales, self.cfg.noise.noise_level, self.obs_scales
        v[:3] = ns.lin_vel * lvl * os_.lin_vel
        v[3:6] = ns.ang_vel * lvl * os_.ang_vel
        v[6:9] = ns.gravity * lvl
        v[9:12] = 0.0
        if self.elspider:      # elspider.py:312-332
            v[12:30] = ns.dof_pos * lvl * os_.dof_pos
            v[30:48] = ns.dof_vel * lvl * os_.dof_vel
            if self.cfg.terrain.measure_heights:
                v[66:253] = ns.height_measurements * lvl * os_.height_measurements
            return v
        v[12:24] = ns.dof_pos * lvl * os_.dof_pos
        v[24:36] = ns.dof_vel * lvl * os_.dof_vel
        v[36:48] = 0.0
        if self.cfg.terrain.measure_heights:
            v[48:235] = ns.height_measurements * lvl * os_.height_measurements
        return v

    def _gather_feet(self):
        rb = self.rigid_body_state.view(self.num_envs, self.num_bodies, 13)
        self.foot_positions = rb[:, self.feet_indices, 0:3]
        self.foot_velocities = rb[:, self.feet_indices, 7:10]

    # ---- a1: torques (legged_robot.py:425-448) -----------------------------------------------------
    def compute_torques(self, actions: Tensor) -> Tensor:
        a = actions * self.cfg.control.action_scale
        mode = self.cfg.control.control_type
        if mode == "P":
            tq = self.p_gains * (a + self.default_dof_pos - self.dof_pos) - self.d_gains * self.dof_vel
        elif mode == "V":
            tq = self.p_gains * (a - self.dof_vel) - self.d_gains * (self.dof_vel - self.last_dof_vel) / self.sim_dt
        elif mode == "T":
            tq = a
        else:
            raise NameError(f"Unknown controller type: {mode}")
        return torch.clip(tq, -self.torque_limits, self.torque_limits)

    # ---- a4: heights (legged_robot.py:900-938) ---------------------------------------------------
    def get_heights(self) -> Tensor:
        if self.cfg.terrain.mesh_type == "plane":
            return torch.zeros(self.num_envs, self.num_height_points)
        pts = quat_apply_yaw(self.base_quat.repeat(1, self.num_height_points), self.height_points) \
            + (self.root_states[:, :3]).unsqueeze(1)
        pts += self.cfg.terrain.border_size
        pts = (pts / self.cfg.terrain.horizontal_scale).long()
        px = torch.clip(pts[:, :, 0].view(-1), 0, self.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].view(-1), 0, self.height_samples.shape[1] - 2)
        self.height_cells = (px.view(self.num_envs, -1), py.view(self.num_envs, -1))   # exposed for index parity
        h = torch.min(self.height_samples[px, py], self.height_samples[px + 1, py])
        h = torch.min(h, self.height_samples[px, py + 1])
        return h.view(self.num_envs, -1) * self.cfg.terrain.vertical_scale

    # ---- a3: callback (legged_robot.py:386-423, 491-496) -----------------------------------------
    def resample_commands(self, env_ids: Tensor):
        r, n = self.command_ranges, len(env_ids)
        self.commands[env_ids, 0] = self.rand(r["lin_vel_x"][0], r["lin_vel_x"][1], (n, 1)).squeeze(1)
        self.commands[env_ids, 1] = self.rand(r["lin_vel_y"][0], r["lin_vel_y"][1], (n, 1)).squeeze(1)
        if self.cfg.commands.heading_command:
            self.commands[env_ids, 3] = self.rand(r["heading"][0], r["heading"][1], (n, 1)).squeeze(1)
        else:
            self.commands[env_ids, 2] = self.rand(r["ang_vel_yaw"][0], r["ang_vel_yaw"][1], (n, 1)).squeeze(1)
        self.commands[env_ids, :2] *= (torch.norm(self.commands[env_ids, :2], dim=1) > 0.2).unsqueeze(1)

    def heading_command(self):
        fwd = tu.quat_apply(self.base_quat, self.forward_vec)
        heading = torch.atan2(fwd[:, 1], fwd[:, 0])
        self.commands[:, 2] = torch.clip(0.5 * wrap_to_pi(self.commands[:, 3] - heading), -1.0, 1.0)

    def callback(self):
        ids = (self.episode_length_buf % int(self.cfg.commands.resampling_time / self.dt) == 0).nonzero(as_tuple=False).flatten()
        self.resample_commands(ids)
        if self.cfg.commands.heading_command:
            self.heading_command()
        if self.measure_heights:
            self.measured_heights = self.get_heights()
        if self.cfg.domain_rand.push_robots and (self.common_step_counter % self.push_interval == 0):
            mv = self.cfg.domain_rand.max_push_vel_xy
            self.root_states[:, 7:9] = self.rand(-mv, mv, (self.num_envs, 2))

    # ---- a2: derived state (legged_robot.py:122-137) ----------------------------------------------
    def derive(self):
        self.episode_length_buf += 1
        self.common_step_counter += 1
        q, rs = self.base_quat, self.root_states
        self.base_lin_vel[:] = tu.quat_rotate_inverse(q, rs[:, 7:10])
        self.base_lin_acc[:] = self.base_lin_acc[:] * self.acc_ema + (1 - self.acc_ema) * \
            tu.quat_rotate_inverse(q, rs[:, 7:10] - self.last_root_vel[:, :3]) / self.dt
        self.base_ang_vel[:] = tu.quat_rotate_inverse(q, rs[:, 10:13])
        self.base_ang_acc[:] = self.base_ang_acc[:] * self.acc_ema + (1 - self.acc_ema) * \
            tu.quat_rotate_inverse(q, rs[:, 10:13] - self.last_root_vel[:, 3:]) / self.dt
        self.projected_gravity[:] = tu.quat_rotate_inverse(q, self.gravity_vec)
        self._gather_feet()

    # ---- a5: termination (legged_robot.py:155-160) ------------------------------------------------
    def check_termination(self):
        f = self.contact_forces[:, self.termination_contact_indices, :]
        self.reset_buf = torch.any(torch.norm(f, dim=-1) > 1.0, dim=1)
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf |= self.time_out_buf
        if self.elspider:      # elspider.py:340-345
            self.reset_buf |= self.projected_gravity[:, 2] > 0

    # ---- a6: rewards (legged_robot.py:215-232) ----------------------------------------------------
    def compute_reward(self):
        self.rew_buf[:] = 0.0
        self.reward_terms_raw = {}
        for name in self.reward_names:
            rew = REWARD_TERMS[name](self) * self.reward_scales[name]
            self.rew_buf += rew
            self.episode_sums[name] += rew
        if self.cfg.rewards.only_positive_rewards:
            self.rew_buf[:] = torch.clip(self.rew_buf[:], min=0.0)
        if "termination" in self.reward_scales:
            rew = _r_termination(self) * self.reward_scales["termination"]
            self.rew_buf += rew
            self.episode_sums["termination"] += rew

    # ---- a7: reset (legged_robot.py:162-213, 450-531) ---------------------------------------------
    def reset_idx(self, env_ids: Tensor):
        if len(env_ids) == 0:
            return
        if self.curriculum:
            self._terrain_curriculum(env_ids)
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self._command_curriculum(env_ids)
        n = len(env_ids)
        self.dof_pos[env_ids] = self.default_dof_pos * self.rand(0.5, 1.5, (n, self.num_dof))
        self.dof_vel[env_ids] = 0.0
        self.root_states[env_ids] = self.base_init_state
        self.root_states[env_ids, :3] += self.env_origins[env_ids]
        if self.custom_origins:
            self.root_states[env_ids, :2] += self.rand(-0.5, 0.5, (n, 2))
        self.root_states[env_ids, 7:13] = self.rand(-0.5, 0.5, (n, 6))
        self.resample_commands(env_ids)
        self.last_actions[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.feet_contact_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        self.extras["episode"] = {}
        for key in self.episode_sums.keys():
            self.extras["episode"]["rew_" + key] = torch.mean(self.episode_sums[key][env_ids]) / self.max_episode_length_s
            self.episode_sums[key][env_ids] = 0.0
        if self.curriculum:
            self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
        if self.cfg.commands.curriculum:
            self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.cfg.rewards.multi_stage_rewards:
            self.extras["episode"]["reward_stage"] = float(self.cfg.rewards.reward_min_stage)
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    def _terrain_curriculum(self, env_ids):
        dist = torch.norm(self.root_states[env_ids, :2] - self.env_origins[env_ids, :2], dim=1)
        up = dist > self.env_length / 2
        down = (dist < torch.norm(self.commands[env_ids, :2], dim=1) * self.max_episode_length_s * 0.5) * ~up
        self.terrain_levels[env_ids] += 1 * up - 1 * down
        self.terrain_levels[env_ids] = torch.where(self.terrain_levels[env_ids] >= self.max_terrain_level,
                                                   torch.randint_like(self.terrain_levels[env_ids], self.max_terrain_level),
                                                   torch.clip(self.terrain_levels[env_ids], 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]

    def _command_curriculum(self, env_ids):
        if torch.mean(self.episode_sums["tracking_lin_vel"][env_ids]) / self.max_episode_length > 0.8 * self.reward_scales["tracking_lin_vel"]:
            mc = self.cfg.commands.max_curriculum
            self.command_ranges["lin_vel_x"][0] = np.clip(self.command_ranges["lin_vel_x"][0] - 0.5, -mc, 0.0)
            self.command_ranges["lin_vel_x"][1] = np.clip(self.command_ranges["lin_vel_x"][1] + 0.5, 0.0, mc)

    # ---- a8: observations (legged_robot.py:234-252) -----------------------------------------------
    def compute_observations(self, noise_u: Optional[Tensor] = None):
        os_ = self.obs_scales
        self.obs_buf = torch.cat((self.base_lin_vel * os_.lin_vel, self.base_ang_vel * os_.ang_vel, self.projected_gravity,
                                  self.commands[:, :3] * self.commands_scale, (self.dof_pos - self.default_dof_pos) * os_.dof_pos,
                                  self.dof_vel * os_.dof_vel, self.actions), dim=-1)
        if self.measure_heights:
            h = torch.clip(self.root_states[:, 2].unsqueeze(1) - 0.5 - self.measured_heights, -1, 1.0) * os_.height_measurements
            self.obs_buf = torch.cat((self.obs_buf, h), dim=-1)
        if self.add_noise:
            u = torch.rand_like(self.obs_buf) if noise_u is None else noise_u
            self.obs_buf += (2 * u - 1) * self.noise_scale_vec

    # ---- a10: gait scheduler (gait_scheduler.py:63-72) ---------------------------------------------
    def gait_step(self):
        self.gait_idx = torch.remainder(self.gait_idx + self.gait_cfg.dt / self.gait_cfg.period, 1.0)
        self.gait_phases = [torch.remainder(self.gait_idx + p, 1.0) for p in self.gait_cfg.foot_phases]
        self.gait_foot_pos = self.foot_positions

    # ---- the whole thing (legged_robot.py:113-150; step() :87-111 around it) ----------------------
    def post_physics_step(self, noise_u: Optional[Tensor] = None, do_reset: bool = True):
        self.derive()
        self.callback()
        self.check_termination()
        self.compute_reward()
        if do_reset:
            self.reset_idx(self.reset_buf.nonzero(as_tuple=False).flatten())
        self.compute_observations(noise_u)
        self.last_actions[:] = self.actions[:]
        self.last_dof_vel[:] = self.dof_vel[:]
        self.last_root_vel[:] = self.root_states[:, 7:13]
        if self.use_gait_scheduler:
            self.gait_step()

    def post_physics_step_rollout(self, noise_u: Optional[Tensor] = None):
        """RobotBatchRollout.post_physics_step_rollout (envs/batch_rollout/robot_batch_rollout.py:763-817) evaluated on every
        row (the caller restores the main rows afterwards): base-frame state and feet, compute_reward_rollout (:969-985:
        the registry WITHOUT episode sums), observations, histories -- no episode counter, no callback (commands,
        heights: ``_post_physics_step_callback_rollout`` is empty), no termination check (the termination term reads the
        flags as they are)."""
        ep, counter = self.episode_length_buf.clone(), self.common_step_counter
        self.derive()
        self.episode_length_buf, self.common_step_counter = ep, counter
        sums = {k: v.clone() for k, v in self.episode_sums.items()}
        self.compute_reward()
        self.episode_sums = sums
        self.compute_observations(noise_u)
        self.last_actions[:] = self.actions[:]
        self.last_dof_vel[:] = self.dof_vel[:]
        self.last_root_vel[:] = self.root_states[:, 7:13]

    def step_no_physics(self, actions: Tensor, noise_u: Optional[Tensor] = None, do_reset: bool = True):
        """step() minus PhysX: clip, one torque evaluation, post-physics, obs clip (legged_robot.py:87-111)."""
        ca = self.cfg.normalization.clip_actions
        self.actions = torch.clip(actions, -ca, ca)
        self.torques = self.compute_torques(self.actions).view(self.torques.shape)
        self.post_physics_step(noise_u, do_reset)
        co = self.cfg.normalization.clip_observations
        self.obs_buf = torch.clip(self.obs_buf, -co, co)
        return self.obs_buf, None, self.rew_buf, self.reset_buf, self.extras

    def hot_step(self, noise_u: Optional[Tensor] = None):
        """The benchmarked unit of work: torques + post-physics body without the sparse
        RNG branches (command resampling, pushes, resets)."""
        self.torques = self.compute_torques(self.actions).view(self.torques.shape)
        self.derive()
        if self.cfg.commands.heading_command:
            self.heading_command()
        if self.measure_heights:
            self.measured_heights = self.get_heights()
        self.check_termination()
        self.compute_reward()
        self.compute_observations(noise_u)
        self.last_actions[:] = self.actions[:]
        self.last_dof_vel[:] = self.dof_vel[:]
        self.last_root_vel[:] = self.root_states[:, 7:13]
