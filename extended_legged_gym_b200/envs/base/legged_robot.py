"""``LeggedRobot`` -- host side of the B200 per-step hot path.

Same constructor, methods and public tensor attributes as the reference class
(envs/base/legged_robot.py:55-939 in /root/reference/legged_gym/legged_gym), but every dense
per-step computation is ONE launch of the hand-written sm_100a kernel behind the C ABI in
``include/elg_b200.h`` instead of ~100 ATen launches:

  step()                    :87-111   clip, decimation x (_compute_torques -> simulator), post_physics_step
  post_physics_step()       :113-150  -> elg_post_physics_step (fused), resets stay sparse host torch
  _compute_torques()        :425-448  -> elg_compute_torques
  _get_heights()            :900-938  -> elg_get_heights (standalone) / fused in the step kernel
  check_termination / compute_reward / compute_observations -> kernel sections (ELG_PHASE_*)

PhysX is out of scope: the simulator is a ``SimBackend`` (``sim_backend.py``); pass one as
``physics_engine`` or let the class create a ``SyntheticSim``.  The sparse, RNG-driven branches
(``reset_idx``, ``_resample_commands``, ``_push_robots``, curricula) keep the reference's
host-driven structure and run as torch ops on the device.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ...sim_backend import SimBackend, SyntheticSim
from ...utils.helpers import class_to_dict
from ...utils.math_utils import torch_rand_float
from .base_task import BaseTask
from .legged_robot_config import LeggedRobotCfg
from .legged_robot_rew_mixin import LeggedRobotRewMixin

_TERRAIN_MESHES = ("heightfield", "trimesh", "confined_trimesh")
# attributes whose device pointers are baked into the ElgStepBuffers struct
_TRACKED = frozenset(_lib._BUF_FIELDS) | {"p_gains", "d_gains", "_episode_sums_all", "_reset_bool"}


class LeggedRobot(BaseTask, LeggedRobotRewMixin):
    def __init__(self, cfg: LeggedRobotCfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        self._ptrs_dirty = True
        self.cfg = cfg
        self._backend = physics_engine if isinstance(physics_engine, SimBackend) else None
        if sim_params is None:
            sim_params = type("SimParams", (), {})()
            sim_params.dt = cfg.sim.dt
            sim_params.use_gpu_pipeline = True
        self.sim_params = sim_params
        self.height_samples = None
        self.debug_viz = False
        self.init_done = False
        self._lib = _lib.load()               # raises when the CUDA library cannot be had -- no CPU fallback
        self._parse_cfg(self.cfg)
        BaseTask.__init__(self, self.cfg, sim_params, physics_engine, sim_device, headless)
        LeggedRobotRewMixin.__init__(self)
        self._init_buffers()
        self._prepare_reward_function()
        self.init_done = True
        self.acc_ema = 0.9

    def __setattr__(self, name, value):
        if name in _TRACKED:
            # step() rebinds torques / reset_buf to the SAME storage every step: the native struct only goes stale when the
            # device pointer changes
            old = self.__dict__.get(name)
            same = (old is value) or (torch.is_tensor(old) and torch.is_tensor(value) and old.data_ptr() == value.data_ptr()
                                      and old.dtype == value.dtype and old.shape == value.shape)
            if not same:
                object.__setattr__(self, "_ptrs_dirty", True)
        object.__setattr__(self, name, value)

    # ------------------------------------------------------------------------------------------
    # configuration (legged_robot.py:847-860)
    # ------------------------------------------------------------------------------------------
    def _parse_cfg(self, cfg):
        self.dt = self.cfg.control.decimation * self.sim_params.dt
        self.obs_scales = self.cfg.normalization.obs_scales
        self.reward_scales_stage = self.cfg.rewards.reward_min_stage
        self.reward_scales = self._get_reward_scales(self.reward_scales_stage)
        self.command_ranges = class_to_dict(self.cfg.commands.ranges)
        if self.cfg.terrain.mesh_type not in _TERRAIN_MESHES:
            self.cfg.terrain.curriculum = False
        self.max_episode_length_s = self.cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)
        self.cfg.domain_rand.push_interval = np.ceil(self.cfg.domain_rand.push_interval_s / self.dt)

    # ------------------------------------------------------------------------------------------
    # simulator boundary (stands in for create_sim/_create_envs, legged_robot.py:254-299, 725-844)
    # ------------------------------------------------------------------------------------------
    def create_sim(self):
        self.up_axis_idx = 2
        if self._backend is None:
            self._backend = SyntheticSim(self.cfg, self.num_envs, self.device)
        sim = self.sim = self._backend
        if sim.num_envs != self.num_envs:
            raise ValueError(f"backend has {sim.num_envs} envs, cfg.env.num_envs = {self.num_envs}")
        spec, cfg = sim.spec, self.cfg
        self.num_dof = self.num_dofs = spec.num_dof
        self.num_bodies = spec.num_bodies
        self.dof_names = list(spec.dof_names)
        dev = self.device
        as_idx = lambda ids: torch.tensor(ids, dtype=torch.long, device=dev)
        self.feet_indices = as_idx(spec.indices_matching(cfg.asset.foot_name))
        self.penalised_contact_indices = as_idx(spec.indices_matching(cfg.asset.penalize_contacts_on))
        self.termination_contact_indices = as_idx(spec.indices_matching(cfg.asset.terminate_after_contacts_on))
        self._store_dof_limits(spec)
        self.base_init_state = torch.tensor(cfg.init_state.pos + cfg.init_state.rot + cfg.init_state.lin_vel + cfg.init_state.ang_vel,
                                            dtype=torch.float, device=dev)
        if cfg.terrain.mesh_type in _TERRAIN_MESHES:
            if sim.height_samples is None:
                raise ValueError(f"mesh_type '{cfg.terrain.mesh_type}' needs backend.height_samples")
            self.height_samples = sim.height_samples
        self._get_env_origins()

    def _store_dof_limits(self, spec):
        # _process_dof_props (legged_robot.py:344-372): URDF limits + soft position limits about the mid-point
        lim = torch.tensor([spec.dof_lower, spec.dof_upper], dtype=torch.float).t().contiguous()
        mid = (lim[:, 0] + lim[:, 1]) / 2
        rng = lim[:, 1] - lim[:, 0]
        soft = self.cfg.rewards.soft_dof_pos_limit
        lim = torch.stack([mid - 0.5 * rng * soft, mid + 0.5 * rng * soft], dim=1)
        self.dof_pos_limits = lim.to(self.device).contiguous()
        self.dof_vel_limits = torch.tensor(spec.dof_velocity, dtype=torch.float, device=self.device)
        self.torque_limits = torch.tensor(spec.dof_effort, dtype=torch.float, device=self.device)

    def _get_env_origins(self):
        cfg, dev, N = self.cfg, self.device, self.num_envs
        self.env_origins = torch.zeros(N, 3, device=dev)
        if cfg.terrain.mesh_type in _TERRAIN_MESHES:
            self.custom_origins = True
            max_init = cfg.terrain.max_init_terrain_level if cfg.terrain.curriculum else cfg.terrain.num_rows - 1
            self.terrain_levels = torch.randint(0, max_init + 1, (N,), device=dev)
            self.terrain_types = torch.div(torch.arange(N, device=dev), (N / cfg.terrain.num_cols), rounding_mode="floor").to(torch.long)
            self.max_terrain_level = cfg.terrain.num_rows
            self.terrain_origins = self.sim.terrain_origins.to(dev).to(torch.float)
            self.env_origins[:] = self.terrain_origins[self.terrain_levels, self.terrain_types]
            self.terrain = type("TerrainInfo", (), {})()
            self.terrain.cfg = cfg.terrain
            self.terrain.env_length = cfg.terrain.terrain_length
            self.terrain.env_width = cfg.terrain.terrain_width
        else:
            self.custom_origins = False
            cols = np.floor(np.sqrt(N))
            rows = np.ceil(N / cols)
            xx, yy = torch.meshgrid(torch.arange(rows), torch.arange(cols), indexing="ij")
            self.env_origins[:, 0] = (cfg.env.env_spacing * xx.flatten()[:N]).to(dev)
            self.env_origins[:, 1] = (cfg.env.env_spacing * yy.flatten()[:N]).to(dev)

    # ------------------------------------------------------------------------------------------
    # buffers (legged_robot.py:559-647, base_task.py:71-79)
    # ------------------------------------------------------------------------------------------
    def _init_buffers(self):
        sim, cfg, dev = self.sim, self.cfg, self.device
        N, D, F = self.num_envs, self.num_dof, len(self.feet_indices)
        z = lambda *s, dtype=torch.float: torch.zeros(*s, dtype=dtype, device=dev)
        self.root_states = sim.root_states
        self.dof_state = sim.dof_state
        self.dof_pos = self.dof_state.view(N, D, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, D, 2)[..., 1]
        self.base_pos = self.root_states[:, :3]
        self.base_quat = self.root_states[:, 3:7]
        self._contact_forces_flat = sim.contact_forces
        self.contact_forces = sim.contact_forces.view(N, -1, 3)
        self.rigid_body_state = sim.rigid_body_state

        self.common_step_counter = 0
        self.extras = {}
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat((N, 1))
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0], device=dev).repeat((N, 1))
        self.torques = z(N, self.num_actions)
        self.p_gains = z(self.num_actions)
        self.d_gains = z(self.num_actions)
        self.actions = z(N, self.num_actions)
        self.last_actions = z(N, self.num_actions)
        self.last_dof_vel = z(N, D)
        self.last_root_vel = z(N, 6)
        self.commands = z(N, cfg.commands.num_commands)
        self.commands_scale = torch.tensor([self.obs_scales.lin_vel, self.obs_scales.lin_vel, self.obs_scales.ang_vel], device=dev)
        self.feet_air_time = z(N, F)
        self.feet_contact_time = z(N, F)
        self.last_contacts = z(N, F, dtype=torch.bool)
        self.foot_positions = z(N, F, 3)
        self.foot_velocities = z(N, F, 3)
        self.base_lin_vel = z(N, 3)
        self.base_ang_vel = z(N, 3)
        self.base_lin_acc = z(N, 3)
        self.base_ang_acc = z(N, 3)
        self.projected_gravity = z(N, 3)
        self._reset_bool = z(N, dtype=torch.bool)
        self.gait_idx = None
        self.gait_prev_foot_z = None
        self.extra_reward = None
        # True: resampling / reset_idx run as predicated kernels with in-kernel Philox (no nonzero(), no host sync);
        # False: the reference's host-driven structure with torch RNG (what the parity tests pin)
        self.fused_reset = True
        self.reset_uniforms = None                # optional [N, 48] uniform table for the fused reset path (tests)
        self._reset_stats = z(_lib.NUM_REWARD_TERMS + 2)          # per-term sums, count, (main-reset flag of the rollout layout)
        self._episode_means = z(_lib.NUM_REWARD_TERMS)
        self.episode_stats = None                 # utils.distributed.ShardedEpisodeStats when envs are sharded over GPUs
        self.noise_u = None                       # set to a [N,O] tensor of U[0,1) for torch.rand_like-parity noise
        self.noise_seed = int(getattr(cfg, "seed", 0) or 0) + 0x5EED
        self._noise_step = 0

        self.measure_heights = bool(cfg.terrain.measure_heights)
        if self.measure_heights:
            self.height_points = self._init_height_points()
            self._height_grid = self.height_points[0].contiguous()      # one [H,3] grid serves every env
            self.measured_heights = z(N, self.num_height_points)
        else:
            self.num_height_points = 0
            self.height_points = None
            self._height_grid = None
            self.measured_heights = None
        self._user_height_points = False

        self.default_dof_pos = z(D)
        for i, name in enumerate(self.dof_names):
            self.default_dof_pos[i] = cfg.init_state.default_joint_angles[name]
            for key in cfg.control.stiffness.keys():
                if key in name:
                    self.p_gains[i] = cfg.control.stiffness[key]
                    self.d_gains[i] = cfg.control.damping[key]
        self.default_dof_pos = self.default_dof_pos.unsqueeze(0)
        self.noise_scale_vec = self._get_noise_scale_vec(cfg)
        self._episode_sums_all = z(_lib.NUM_REWARD_TERMS, N)

    def _init_height_points(self):
        # legged_robot.py:884-898: 'ij' meshgrid -> point index ix * len(y) + iy
        cfg = self.cfg.terrain
        x = torch.tensor(cfg.measured_points_x, device=self.device)
        y = torch.tensor(cfg.measured_points_y, device=self.device)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        self.num_height_points = gx.numel()
        pts = torch.zeros(self.num_envs, self.num_height_points, 3, device=self.device)
        pts[:, :, 0] = gx.flatten()
        pts[:, :, 1] = gy.flatten()
        return pts

    def _get_noise_scale_vec(self, cfg):
        # legged_robot.py:533-556, generalised from the hard-coded 12-DOF slices to 12+3D(+H)
        v = torch.zeros(self.num_obs, device=self.device)
        self.add_noise = cfg.noise.add_noise
        ns, lvl, D = cfg.noise.noise_scales, cfg.noise.noise_level, self.num_dof
        v[:3] = ns.lin_vel * lvl * self.obs_scales.lin_vel
        v[3:6] = ns.ang_vel * lvl * self.obs_scales.ang_vel
        v[6:9] = ns.gravity * lvl
        v[12:12 + D] = ns.dof_pos * lvl * self.obs_scales.dof_pos
        v[12 + D:12 + 2 * D] = ns.dof_vel * lvl * self.obs_scales.dof_vel
        if cfg.terrain.measure_heights:
            h0 = 12 + 3 * D
            v[h0:h0 + self.num_height_points] = ns.height_measurements * lvl * self.obs_scales.height_measurements
        return v

    def set_env_state(self, state):
        """Load env-owned history/state tensors (e.g. from ``synthetic.make_state``) -- test/bench helper."""
        for k in ("actions", "last_actions", "last_dof_vel", "last_root_vel", "commands", "feet_air_time",
                  "feet_contact_time", "base_lin_acc", "base_ang_acc"):
            if k in state:
                getattr(self, k).copy_(state[k].to(self.device))
        if "last_contacts" in state:
            self.last_contacts.copy_(state["last_contacts"].to(self.device))
        if "episode_length_buf" in state:
            self.episode_length_buf.copy_(state["episode_length_buf"].to(self.device))
        if "gait_idx" in state and self.gait_idx is not None:
            self.gait_idx.copy_(state["gait_idx"].to(self.device))

    # ------------------------------------------------------------------------------------------
    # reward registry (legged_robot.py:649-674) -> kernel term mask + Python-side extras
    # ------------------------------------------------------------------------------------------
    def _prepare_reward_function(self):
        for key in list(self.reward_scales.keys()):
            if self.reward_scales[key] == 0:
                self.reward_scales.pop(key)
            else:
                self.reward_scales[key] *= self.dt
        self.reward_functions, self.reward_names = [], []
        self._kernel_terms, self._python_terms = [], []
        for name in self.reward_scales:
            if name == "termination":
                continue
            fn = getattr(self, "_reward_" + name)
            self.reward_names.append(name)
            self.reward_functions.append(fn)
            stock = getattr(getattr(fn, "__func__", fn), "_elg_stock", False)
            (self._kernel_terms if (stock and name in _lib.TERM_ID) else self._python_terms).append(name)
        # episode sums: rows of one [terms, N] tensor for kernel terms (SoA), separate tensors otherwise
        self._episode_sums_all.zero_()
        self.episode_sums = {}
        for name in self.reward_scales:
            if name in _lib.TERM_ID and name not in self._python_terms:
                self.episode_sums[name] = self._episode_sums_all[_lib.TERM_ID[name]]
            else:
                self.episode_sums[name] = torch.zeros(self.num_envs, dtype=torch.float, device=self.device)
        self._params_dirty = True

    # ------------------------------------------------------------------------------------------
    # native structs
    # ------------------------------------------------------------------------------------------
    def _native_dims(self):
        d = _lib.ElgDims()
        d.num_envs, d.num_dof, d.num_bodies = self.num_envs, self.num_dof, self.num_bodies
        feet = self.feet_indices.tolist()
        pen = self.penalised_contact_indices.tolist()
        term = self.termination_contact_indices.tolist()
        d.num_feet, d.num_penalised, d.num_termination = len(feet), len(pen), len(term)
        if len(feet) > _lib.MAX_FEET or len(pen) > _lib.MAX_PENALISED or len(term) > _lib.MAX_TERMINATION or self.num_dof > _lib.MAX_DOF:
            raise _lib.ElgError("robot exceeds the ELG_MAX_* limits of include/elg_b200.h")
        d.num_height_points = self.num_height_points if self.measure_heights else 0
        d.num_obs, d.num_commands = self.num_obs, self.cfg.commands.num_commands
        for i, v in enumerate(feet):
            d.feet_idx[i] = v
        for i, v in enumerate(pen):
            d.penalised_idx[i] = v
        for i, v in enumerate(term):
            d.termination_idx[i] = v
        return d

    def _native_params(self):
        cfg, p = self.cfg, _lib.ElgStepParams()
        p.dt, p.sim_dt = self.dt, self.sim_params.dt
        p.acc_ema, p.acc_ema_c = self.acc_ema, 1 - self.acc_ema
        p.max_episode_length = int(np.floor(self.max_episode_length))
        if cfg.control.control_type not in _lib.CONTROL_TYPES:
            raise NameError(f"Unknown controller type: {cfg.control.control_type}")
        p.control_type = _lib.CONTROL_TYPES[cfg.control.control_type]
        p.action_scale = cfg.control.action_scale
        p.heading_command = int(bool(cfg.commands.heading_command))
        p.measure_heights = int(self.measure_heights)
        p.terrain_is_plane = int(cfg.terrain.mesh_type == "plane")
        p.only_positive_rewards = int(bool(cfg.rewards.only_positive_rewards))
        p.noise_mode = _lib.NOISE_OFF
        p.clip_observations = 0.0
        p.gravity_vec[:] = [0.0, 0.0, -1.0]
        os_ = self.obs_scales
        p.obs_scale_lin_vel, p.obs_scale_ang_vel = os_.lin_vel, os_.ang_vel
        p.obs_scale_dof_pos, p.obs_scale_dof_vel = os_.dof_pos, os_.dof_vel
        p.obs_scale_height = os_.height_measurements
        p.commands_scale[:] = [os_.lin_vel, os_.lin_vel, os_.ang_vel]
        p.border_size, p.horizontal_scale, p.vertical_scale = cfg.terrain.border_size, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale
        if self.height_samples is not None:
            p.hf_rows, p.hf_cols = self.height_samples.shape
        p.height_points_env_stride = self.num_height_points * 3 if self._user_height_points else 0
        mask = 0
        for name in self._kernel_terms + (["termination"] if "termination" in self.reward_scales else []):
            t = _lib.TERM_ID[name]
            mask |= 1 << t
            p.reward_scales[t] = self.reward_scales[name]
        p.reward_mask = mask
        rw = cfg.rewards
        p.tracking_sigma, p.base_height_target, p.max_contact_force = rw.tracking_sigma, rw.base_height_target, rw.max_contact_force
        p.soft_dof_vel_limit, p.soft_torque_limit = rw.soft_dof_vel_limit, rw.soft_torque_limit
        p.speed_min = p.stand_still_threshold = self.speed_min
        gait = getattr(self, "gait_cfg", None)
        if gait is not None:
            p.gait_increment = gait.dt / gait.period
            p.gait_swing_height = gait.swing_height
            for i, ph in enumerate(gait.foot_phases[:_lib.MAX_FEET]):
                p.gait_foot_phases[i] = ph
        p.noise_seed = self.noise_seed
        p.rows_per_main = int(getattr(self, "_rows_per_main", 0))     # main / rollout layout (RobotBatchRollout), 0 = flat
        if self.measure_heights and not self._user_height_points:
            h0 = 12 + 3 * self.num_dof
            nmax = float(self.noise_scale_vec[h0:h0 + self.num_height_points].abs().max()) if self.add_noise else 0.0
            p.height_obs_bound = abs(float(os_.height_measurements)) * 1.0 + nmax
        return p

    def _native_buffers(self):
        b = _lib.ElgStepBuffers()
        self._keepalive = []

        def put(field, t, dtype=None):
            if t is None:
                setattr(b, field, None)
                return
            if not t.is_cuda or not t.is_contiguous() or (dtype is not None and t.dtype != dtype):
                raise _lib.ElgError(f"buffer '{field}' must be a contiguous CUDA tensor of {dtype}, got {t.dtype} "
                                    f"{tuple(t.shape)} contiguous={t.is_contiguous()} on {t.device}")
            self._keepalive.append(t)
            setattr(b, field, t.data_ptr())

        f32 = torch.float
        put("root_states", self.root_states, f32)
        put("dof_state", self.dof_state, f32)
        put("contact_forces", self._contact_forces_flat, f32)
        put("rigid_body_state", self.rigid_body_state, f32)
        put("actions", self.actions, f32)
        put("torques", self.torques, f32)
        put("default_dof_pos", self.default_dof_pos.view(-1), f32)
        put("dof_pos_limits", self.dof_pos_limits, f32)
        put("dof_vel_limits", self.dof_vel_limits, f32)
        put("torque_limits", self.torque_limits, f32)
        put("height_samples", self.height_samples, torch.int16)
        put("height_field_min", self._height_field_min(), f32)
        put("height_points", self.height_points if self._user_height_points else self._height_grid, f32)
        put("noise_scale_vec", self.noise_scale_vec, f32)
        put("noise_u", self.noise_u, f32)
        put("extra_reward", self.extra_reward, f32)
        for name in ("last_actions", "last_dof_vel", "last_root_vel", "base_lin_acc", "base_ang_acc", "commands",
                     "feet_air_time", "feet_contact_time", "gait_idx", "gait_prev_foot_z", "base_lin_vel", "base_ang_vel",
                     "projected_gravity", "foot_positions", "foot_velocities", "measured_heights", "rew_buf", "obs_buf"):
            put(name, getattr(self, name), f32)
        put("last_contacts", self.last_contacts, torch.bool)
        put("episode_length_buf", self.episode_length_buf, torch.int64)
        put("episode_sums", self._episode_sums_all, f32)
        put("reset_buf", self._reset_bool, torch.bool)
        put("time_out_buf", self.time_out_buf, torch.bool)
        # the per-joint constants once more as one block (the lean kernel fetches it with a single bulk copy)
        key = tuple(t.data_ptr() for t in (self.default_dof_pos, self.dof_pos_limits, self.dof_vel_limits, self.torque_limits))
        if getattr(self, "_dof_consts_key", None) != key:
            object.__setattr__(self, "_dof_consts", torch.cat([self.default_dof_pos.view(-1), self.dof_pos_limits.reshape(-1),
                                                                self.dof_vel_limits.view(-1), self.torque_limits.view(-1)]).contiguous())
            object.__setattr__(self, "_dof_consts_key", key)
        put("dof_consts", self._dof_consts, f32)
        put("step_counter", getattr(self, "_step_counter", None), torch.int64)
        return b

    def _height_field_min(self):
        """fp32 min-of-3-cells table of the (static) terrain, built once by elg_prepare_height_field and rebuilt
        when ``height_samples`` is rebound; ``measure_heights`` off or plane terrain: None."""
        hs = self.height_samples
        if hs is None or not self.measure_heights or self.cfg.terrain.mesh_type == "plane":
            return None
        key = (hs.data_ptr(), tuple(hs.shape), float(self.cfg.terrain.vertical_scale))
        if getattr(self, "_hmin_key", None) != key:
            out = torch.empty(hs.shape, dtype=torch.float, device=hs.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self._lib.elg_prepare_height_field(hs.data_ptr(), hs.shape[0], hs.shape[1],
                                                          float(self.cfg.terrain.vertical_scale), out.data_ptr(), stream),
                       "elg_prepare_height_field")
            object.__setattr__(self, "_hmin", out)
            object.__setattr__(self, "_hmin_key", key)
        return self._hmin

    def _sync_native(self):
        # the library launches on the CURRENT device (its per-device caches and the kernels' pointers must agree with the stream)
        idx = self.__dict__.get("_device_index")
        if idx is None:
            idx = torch.device(self.device).index
            idx = torch.cuda.current_device() if idx is None else idx
            object.__setattr__(self, "_device_index", idx)
        if torch.cuda.current_device() != idx:
            raise _lib.ElgError(f"this env lives on cuda:{idx} but cuda:{torch.cuda.current_device()} is the current device: "
                                f"wrap the call in `with torch.cuda.device({idx!r}):` (one env object per GPU, one process per GPU is the intended use)")
        if getattr(self, "_dims", None) is None:
            self._dims = self._native_dims()
        if getattr(self, "_params_dirty", True):
            self._params = self._native_params()
            self._params_dirty = False
        if self._ptrs_dirty:
            self._bufs = self._native_buffers()
            object.__setattr__(self, "_ptrs_dirty", False)

    def _launch(self, phase: int, clip_obs: float = 0.0, rollout: bool = False, noise_step: int = 0):
        self._sync_native()
        p = self._params
        p.clip_observations = clip_obs
        p.rollout_mode = int(rollout)
        if not self.add_noise:
            p.noise_mode = _lib.NOISE_OFF
        else:
            p.noise_mode = _lib.NOISE_TENSOR if self.noise_u is not None else _lib.NOISE_PHILOX
        p.noise_offset = self._noise_step + noise_step
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_post_physics_step(C.byref(self._dims), C.byref(p), C.byref(self._bufs), phase, stream)
        _lib.check(rc, "elg_post_physics_step")

    # ------------------------------------------------------------------------------------------
    # step (legged_robot.py:87-111)
    # ------------------------------------------------------------------------------------------
    def step(self, actions):
        clip_actions = self.cfg.normalization.clip_actions
        torch.clamp(actions.to(self.device), -clip_actions, clip_actions, out=self.actions)
        self.render()
        for _ in range(self.cfg.control.decimation):
            self.torques = self._compute_torques(self.actions).view(self.torques.shape)
            self.sim.set_dof_actuation_force(self.torques)
            self.sim.simulate()
            self.sim.refresh()
        self._obs_clip_for_step = self.cfg.normalization.clip_observations
        try:
            self.post_physics_step()
        finally:
            self._obs_clip_for_step = 0.0
        if self.privileged_obs_buf is not None:
            c = self.cfg.normalization.clip_observations
            self.privileged_obs_buf = torch.clip(self.privileged_obs_buf, -c, c)
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _compute_torques(self, actions):
        self._sync_native()
        if not (actions.is_cuda and actions.is_contiguous() and actions.dtype == torch.float):
            actions = actions.to(self.device, torch.float).contiguous()
        out = self.torques
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_compute_torques(C.byref(self._dims), C.byref(self._params), actions.data_ptr(), self.dof_state.data_ptr(),
                                           self.last_dof_vel.data_ptr(), self.p_gains.data_ptr(), self.d_gains.data_ptr(),
                                           self.torque_limits.data_ptr(), self.default_dof_pos.data_ptr(), out.data_ptr(),
                                           None, 0, stream)
        _lib.check(rc, "elg_compute_torques")
        return out

    # ------------------------------------------------------------------------------------------
    # post-physics (legged_robot.py:113-150)
    # ------------------------------------------------------------------------------------------
    def post_physics_step(self):
        self.sim.refresh()
        self.common_step_counter += 1
        clip = getattr(self, "_obs_clip_for_step", 0.0)
        P = _lib
        dr = self.cfg.domain_rand
        push_step = bool(dr.push_robots and (self.common_step_counter % dr.push_interval == 0))
        cmd_curr = bool(self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0))
        if self.fused_reset and not self._python_terms and not push_step and not cmd_curr:
            # the whole step without a host synchronisation: resample -> fused step -> reset (+ observation repair)
            self._launch_resample()
            self._pre_step_hook()
            self._launch(P.PHASE_FUSED, clip)
            self.reset_buf = self._reset_bool
            self._launch_reset()
            self._noise_step += 1
            return
        push_now = self._post_physics_step_callback()
        self._pre_step_hook()
        if not self._python_terms and not push_now:
            # common case: the whole step is ONE kernel; the reset path below re-runs the cheap POST section
            self._launch(P.PHASE_FUSED, clip)
            self.reset_buf = self._reset_bool
            env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
            if len(env_ids):
                self.reset_idx(env_ids)
                self._launch(P.PHASE_POST, clip)
        else:
            if self._python_terms:
                self._launch(P.PHASE_DERIVE | P.PHASE_TERMINATION)
                self.reset_buf = self._reset_bool
                self._python_reward_terms()
                self._launch(P.PHASE_REWARD)
            else:
                self._launch(P.PHASE_PRE)
                self.reset_buf = self._reset_bool
            if push_now:
                self._push_robots()
            env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
            self.reset_idx(env_ids)
            self._launch(P.PHASE_POST, clip)
        self._noise_step += 1

    # ------------------------------------------------------------------------------------------
    # fused reset path (csrc/elg_reset.cu): reset_idx / _resample_commands as predicated kernels
    # ------------------------------------------------------------------------------------------
    def _native_reset(self):
        cfg, rp = self.cfg, _lib.ElgResetParams()
        r = self.command_ranges
        rp.lin_vel_x[:], rp.lin_vel_y[:] = r["lin_vel_x"], r["lin_vel_y"]
        rp.ang_vel_yaw[:], rp.heading[:] = r["ang_vel_yaw"], r["heading"]
        rp.heading_command = int(bool(cfg.commands.heading_command))
        rp.resample_interval = int(cfg.commands.resampling_time / self.dt)
        rp.base_init_state[:] = self.base_init_state.tolist()
        rp.custom_origins = int(bool(self.custom_origins))
        rp.curriculum = int(bool(cfg.terrain.curriculum and self.init_done))
        if rp.curriculum:
            rp.env_length_half = self.terrain.env_length / 2
            rp.max_terrain_level = int(self.max_terrain_level)
            rp.terrain_cols = int(self.terrain_origins.shape[1])
        rp.max_episode_length_s = self.max_episode_length_s
        rp.rows_per_main = int(getattr(self, "_rows_per_main", 0))
        rp.root_z_from_terrain = int(bool(getattr(self, "_reset_z_from_terrain", False) and self.custom_origins))
        rp.seed = self.noise_seed
        b = _lib.ElgResetBuffers()
        t = lambda x: None if x is None else x.data_ptr()
        b.reset_buf, b.root_states, b.dof_state = t(self._reset_bool), t(self.root_states), t(self.dof_state)
        b.commands, b.env_origins = t(self.commands), t(self.env_origins)
        if rp.curriculum:
            b.terrain_levels, b.terrain_types, b.terrain_origins = t(self.terrain_levels), t(self.terrain_types), t(self.terrain_origins)
        b.default_dof_pos = t(self.default_dof_pos)
        b.last_dof_vel, b.last_root_vel = t(self.last_dof_vel), t(self.last_root_vel)
        b.feet_air_time, b.feet_contact_time = t(self.feet_air_time), t(self.feet_contact_time)
        b.episode_length_buf, b.episode_sums, b.stats = t(self.episode_length_buf), t(self._episode_sums_all), t(self._reset_stats)
        b.obs_buf, b.noise_scale_vec = t(self.obs_buf), t(self.noise_scale_vec)
        b.measured_heights = t(self.measured_heights)
        b.height_samples = t(self.height_samples) if rp.root_z_from_terrain else None
        # sharded envs: the same (sum, count) also goes into the running totals that are all-reduced once per K steps
        b.stats_accum = t(self.episode_stats.buf) if self.episode_stats is not None else None
        for name in ("env_origins", "terrain_levels", "terrain_origins", "commands"):
            x = getattr(self, name, None)
            if x is not None and not x.is_contiguous():
                raise _lib.ElgError(f"fused reset: '{name}' must be contiguous")
        if rp.curriculum and (self.terrain_levels.dtype != torch.int64 or self.terrain_types.dtype != torch.int64):
            raise _lib.ElgError("fused reset: terrain_levels / terrain_types must be int64")
        return rp, b

    def _reset_native_synced(self):
        key = (self._reset_bool.data_ptr(), self.obs_buf.data_ptr(), self.commands.data_ptr(), self.env_origins.data_ptr(),
               self.root_states.data_ptr(), tuple(self.command_ranges["lin_vel_x"]), self.init_done,
               None if self.episode_stats is None else self.episode_stats.buf.data_ptr())
        if getattr(self, "_reset_key", None) != key:
            # a few recent bindings are kept: a caller alternating between output blocks (double-buffered copy-out) must not rebuild
            # the structs -- reading base_init_state back is a host synchronisation -- inside a CUDA-graph capture
            cache = self.__dict__.setdefault("_reset_cache", {})
            if key not in cache:
                if len(cache) >= 8:
                    cache.clear()
                cache[key] = self._native_reset()
            self._reset_rp, self._reset_bufs = cache[key]
            self._reset_key = key
        self._reset_rp.offset = self._noise_step
        self._reset_bufs.noise_u = _lib.ptr(self.noise_u)
        self._reset_bufs.uniforms = _lib.ptr(self.reset_uniforms)
        return self._reset_rp, self._reset_bufs

    def _launch_resample(self):
        self._sync_native()
        rp, b = self._reset_native_synced()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        # (the kernel also zeroes this step's extras["episode"] accumulators, which elg_reset_envs adds to after the step)
        _lib.check(self._lib.elg_resample_commands(C.byref(self._dims), C.byref(rp), self.episode_length_buf.data_ptr(), self.commands.data_ptr(),
                                                   _lib.ptr(self.reset_uniforms), self._reset_stats.data_ptr(), stream), "elg_resample_commands")

    def _launch_reset(self):
        rp, b = self._reset_native_synced()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.elg_reset_envs(C.byref(self._dims), C.byref(rp), C.byref(self._params), C.byref(b), stream), "elg_reset_envs")
        if self.episode_stats is not None:
            # sharded envs: the kernel has added this step's (sum, count) to the running totals; extras["episode"] comes from
            # episode_stats.reduce() -- one all-reduce per K steps -- instead of per-step means of the local shard
            if self.cfg.env.send_timeouts:
                self.extras["time_outs"] = self.time_out_buf
            self.sim.set_dof_state()
            self.sim.set_root_state()
            return
        # extras["episode"] (legged_robot.py:200-213) as device tensors: means over the envs that reset this step; steps
        # without a reset keep the previous values (the reference leaves the dict untouched then)
        cnt = self._reset_stats[_lib.NUM_REWARD_TERMS]
        means = self._reset_stats[:_lib.NUM_REWARD_TERMS] / (cnt * self.max_episode_length_s)
        self._episode_means = torch.where(cnt > 0, means, self._episode_means)
        ep = {"rew_" + k: self._episode_means[_lib.TERM_ID[k]] for k in self.episode_sums if k in _lib.TERM_ID}
        if self.cfg.terrain.curriculum:
            ep["terrain_level"] = torch.mean(self.terrain_levels.float())
        if self.cfg.commands.curriculum:
            ep["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.cfg.rewards.multi_stage_rewards:
            ep["reward_stage"] = float(self.reward_scales_stage)
        self.extras["episode"] = ep
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
        self.sim.set_dof_state()
        self.sim.set_root_state()

    def _stats_rows(self, env_ids):
        """Rows whose episode returns enter extras["episode"] (all reset envs; the rollout layout narrows it to main envs)."""
        return env_ids

    def _pre_step_hook(self):
        """Called after the command resampling and before the fused step kernel: the place where a subclass's
        ``_post_physics_step_callback`` additions that feed this step's rewards / observations belong (e.g. the navigation
        command update of RobotBatchRolloutNav).  The in-kernel heading update and height scan come after it."""

    def _post_physics_step_callback(self):
        """Host part of the callback (legged_robot.py:386-403): sparse command resampling before the
        kernel; returns whether this is a push step.  Heading command and heights are in-kernel."""
        interval = int(self.cfg.commands.resampling_time / self.dt)
        env_ids = ((self.episode_length_buf + 1) % interval == 0).nonzero(as_tuple=False).flatten()
        self._resample_commands(env_ids)
        dr = self.cfg.domain_rand
        return bool(dr.push_robots and (self.common_step_counter % dr.push_interval == 0))

    def _python_reward_terms(self):
        """Overridden / user-defined terms, evaluated with torch in registry order."""
        extra = torch.zeros(self.num_envs, device=self.device)
        for name in self._python_terms:
            rew = getattr(self, "_reward_" + name)() * self.reward_scales[name]
            extra += rew
            self.episode_sums[name] += rew
        self.extra_reward = extra

    # individually callable sections (API parity with legged_robot.py:155-160, :215-252)
    def check_termination(self):
        self._launch(_lib.PHASE_TERMINATION)
        self.reset_buf = self._reset_bool

    def compute_reward(self):
        if self._python_terms:
            self._python_reward_terms()
        self._launch(_lib.PHASE_REWARD)

    def compute_observations(self):
        self._launch(_lib.PHASE_OBS, getattr(self, "_obs_clip_for_step", 0.0))

    def _get_heights(self, env_ids=None):
        if self.cfg.terrain.mesh_type == "plane":
            return torch.zeros(self.num_envs, self.num_height_points, device=self.device, requires_grad=False)
        if self.cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        self._sync_native()
        out = torch.empty(self.num_envs, self.num_height_points, device=self.device)
        pts = self.height_points if self._user_height_points else self._height_grid
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_get_heights(C.byref(self._dims), C.byref(self._params), self.root_states.data_ptr(),
                                       self.height_samples.data_ptr(), pts.data_ptr(), out.data_ptr(), None, stream)
        _lib.check(rc, "elg_get_heights")
        return out if env_ids is None else out[env_ids]

    # ------------------------------------------------------------------------------------------
    # sparse RNG-driven paths (legged_robot.py:162-213, 405-423, 450-531) -- torch ops on the device
    # ------------------------------------------------------------------------------------------
    def _rand(self, lower, upper, shape):
        return torch_rand_float(lower, upper, shape, device=self.device)

    def _randint_like(self, t, high):
        return torch.randint_like(t, high)

    def _resample_commands(self, env_ids):
        if len(env_ids) == 0:
            return
        r, n = self.command_ranges, len(env_ids)
        self.commands[env_ids, 0] = self._rand(r["lin_vel_x"][0], r["lin_vel_x"][1], (n, 1)).squeeze(1)
        self.commands[env_ids, 1] = self._rand(r["lin_vel_y"][0], r["lin_vel_y"][1], (n, 1)).squeeze(1)
        if self.cfg.commands.heading_command:
            self.commands[env_ids, 3] = self._rand(r["heading"][0], r["heading"][1], (n, 1)).squeeze(1)
        else:
            self.commands[env_ids, 2] = self._rand(r["ang_vel_yaw"][0], r["ang_vel_yaw"][1], (n, 1)).squeeze(1)
        self.commands[env_ids, :2] *= (torch.norm(self.commands[env_ids, :2], dim=1) > 0.2).unsqueeze(1)

    def _reset_dofs(self, env_ids):
        self.dof_pos[env_ids] = self.default_dof_pos * self._rand(0.5, 1.5, (len(env_ids), self.num_dof))
        self.dof_vel[env_ids] = 0.0
        self.sim.set_dof_state_indexed(env_ids.to(dtype=torch.int32))

    def _reset_root_states(self, env_ids):
        self.root_states[env_ids] = self.base_init_state
        self.root_states[env_ids, :3] += self.env_origins[env_ids]
        if self.custom_origins:
            self.root_states[env_ids, :2] += self._rand(-0.5, 0.5, (len(env_ids), 2))
        self.root_states[env_ids, 7:13] = self._rand(-0.5, 0.5, (len(env_ids), 6))
        self.sim.set_root_state_indexed(env_ids.to(dtype=torch.int32))

    def _push_robots(self):
        mv = self.cfg.domain_rand.max_push_vel_xy
        self.root_states[:, 7:9] = self._rand(-mv, mv, (self.num_envs, 2))
        self.sim.set_root_state()

    def _update_terrain_curriculum(self, env_ids):
        if not self.init_done:
            return
        dist = torch.norm(self.root_states[env_ids, :2] - self.env_origins[env_ids, :2], dim=1)
        up = dist > self.terrain.env_length / 2
        down = (dist < torch.norm(self.commands[env_ids, :2], dim=1) * self.max_episode_length_s * 0.5) * ~up
        self.terrain_levels[env_ids] += 1 * up - 1 * down
        lv = self.terrain_levels[env_ids]
        self.terrain_levels[env_ids] = torch.where(lv >= self.max_terrain_level, self._randint_like(lv, self.max_terrain_level),
                                                   torch.clip(lv, 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]

    def update_command_curriculum(self, env_ids):
        good = torch.mean(self.episode_sums["tracking_lin_vel"][env_ids]) / self.max_episode_length
        if good > 0.8 * self.reward_scales["tracking_lin_vel"]:
            mc = self.cfg.commands.max_curriculum
            self.command_ranges["lin_vel_x"][0] = np.clip(self.command_ranges["lin_vel_x"][0] - 0.5, -mc, 0.0)
            self.command_ranges["lin_vel_x"][1] = np.clip(self.command_ranges["lin_vel_x"][1] + 0.5, 0.0, mc)

    def reset_idx(self, env_ids):
        if len(env_ids) == 0:
            return
        if self.cfg.terrain.curriculum:
            self._update_terrain_curriculum(env_ids)
        if self.cfg.commands.curriculum and (self.common_step_counter % self.max_episode_length == 0):
            self.update_command_curriculum(env_ids)
        self._reset_dofs(env_ids)
        self._reset_root_states(env_ids)
        self._resample_commands(env_ids)
        self.last_actions[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.feet_contact_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = 1
        if getattr(self, "episode_stats", None) is not None:    # sharded envs: (sum, count) now, all-reduce later (utils/distributed.py)
            self.episode_stats.accumulate(self._episode_sums_all, self._stats_rows(env_ids))
        self.extras["episode"] = {}
        for key in self.episode_sums.keys():
            self.extras["episode"]["rew_" + key] = torch.mean(self.episode_sums[key][env_ids]) / self.max_episode_length_s
            self.episode_sums[key][env_ids] = 0.0
        if self.cfg.terrain.curriculum:
            self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
        if self.cfg.commands.curriculum:
            self.extras["episode"]["max_command_x"] = self.command_ranges["lin_vel_x"][1]
        if self.cfg.rewards.multi_stage_rewards:
            self.extras["episode"]["reward_stage"] = float(self.reward_scales_stage)
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf
