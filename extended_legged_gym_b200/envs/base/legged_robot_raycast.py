"""``LeggedRobotRayCast`` -- LeggedRobot with a ray-cast sensor whose normalised hit distances join the observations.

Mirrors envs/base/legged_robot_raycast.py:76-297 of the reference (``_init_ray_caster`` :101, the callback :219,
``compute_observations`` :232, ``_get_raycast_distances`` :262).  What the reference does per step as
RayCaster.update (GPU -> CPU numpy -> Warp -> CPU -> GPU) + a norm / clamp / mask chain + a ``torch.cat`` into a new
observation tensor is ONE launch here: ``elg_raycast_sensor_obs`` rotates the pattern, walks the BVH and writes
``(1 - clamp(|hit - base| / max, 0, 1)) * found`` straight into the trailing ``num_rays`` columns of ``obs_buf``; the step
kernel fills the columns in front of them (``cfg.env.num_observations`` must count the rays, as in the reference's configs).
"""
import numpy as np
import torch

from ...utils.ray_caster import PatternType, RayCaster, RayCasterCfg, RayCasterPatternCfg
from ... import synthetic
from .legged_robot import LeggedRobot


class RayCastSensorMixin:
    """Ray-caster set-up and the in-place observation binding, shared by LeggedRobotRayCast and RobotBatchRolloutPercept
    (the reference repeats ``_init_ray_caster`` / ``_get_raycast_distances`` in both classes)."""

    def _pattern_cfg(self):
        rc = self.cfg.raycaster
        kind = rc.ray_pattern
        if kind == "single":
            return RayCasterPatternCfg(pattern_type=PatternType.SINGLE_RAY)
        if kind == "grid":
            return RayCasterPatternCfg(pattern_type=PatternType.GRID, grid_dims=(5, 5), grid_width=2.0, grid_height=2.0)
        if kind == "spherical":
            return RayCasterPatternCfg(pattern_type=PatternType.SPHERICAL, spherical_num_azimuth=getattr(rc, "spherical_num_azimuth", 8),
                                       spherical_num_elevation=getattr(rc, "spherical_num_elevation", 4))
        if kind == "spherical2":
            return RayCasterPatternCfg(pattern_type=PatternType.SPHERICAL2, spherical2_num_points=getattr(rc, "spherical2_num_points", 32),
                                       spherical2_polar_axis=getattr(rc, "spherical2_polar_axis", [0.0, 0.0, 1.0]))
        if kind != "cone":
            print(f"Unknown pattern type: {kind}. Using cone pattern.")
        return RayCasterPatternCfg(pattern_type=PatternType.CONE, cone_num_rays=rc.num_rays, cone_angle=rc.ray_angle)

    def _terrain_mesh(self):
        """(vertices [V,3], triangles [M,3]) in world coordinates: the backend's terrain mesh, the triangulated height field,
        or a 200 m ground quad for plane terrain (:168-206)."""
        sim, cfg = self.sim, self.cfg
        v, t = getattr(sim, "terrain_vertices", None), getattr(sim, "terrain_triangles", None)
        if v is not None and t is not None:
            v = torch.as_tensor(v, dtype=torch.float).clone()
            if hasattr(cfg.terrain, "border_size") and not getattr(sim, "terrain_mesh_in_world_frame", False):
                v[:, 0] -= cfg.terrain.border_size          # the Terrain classes keep vertices in map coordinates (:183-186)
                v[:, 1] -= cfg.terrain.border_size
            return v, torch.as_tensor(t, dtype=torch.int32)
        if self.height_samples is not None and cfg.terrain.mesh_type != "plane":
            v, t = synthetic.heightfield_to_trimesh(self.height_samples, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.border_size)
            return torch.from_numpy(v), torch.from_numpy(t)
        if cfg.terrain.mesh_type == "plane":
            size = 100.0
            v = np.array([[-size, -size, 0.0], [size, -size, 0.0], [size, size, 0.0], [-size, size, 0.0]], dtype=np.float32)
            return torch.from_numpy(v), torch.tensor([[0, 1, 2], [0, 2, 3]], dtype=torch.int32)
        raise ValueError("No terrain mesh available for ray casting. Either set use_terrain_obj=True and provide a terrain file, or ensure "
                         "terrain.vertices and terrain.triangles are available.")

    def _init_ray_caster(self):
        rc = self.cfg.raycaster
        rcfg = RayCasterCfg(pattern_cfg=self._pattern_cfg(), max_distance=getattr(rc, "max_distance", 10.0),
                            offset_pos=getattr(rc, "offset_pos", [0.0, 0.0, 0.0]), attach_yaw_only=getattr(rc, "attach_yaw_only", False))
        if getattr(self.cfg.terrain, "use_terrain_obj", False) and rc.terrain_file:
            rcfg.mesh_paths = [rc.terrain_file]
        else:
            rcfg.vertices, rcfg.triangles = self._terrain_mesh()
        self.ray_caster = RayCaster(rcfg, self.num_envs, self.device)
        self.num_ray_observations = self.ray_caster.num_rays
        base = 12 + 3 * self.num_dof + (self.num_height_points if self.measure_heights else 0)
        extra = self._obs_columns_behind_rays()
        if self.num_obs != base + self.num_ray_observations + extra:
            raise ValueError(f"cfg.env.num_observations = {self.num_obs}, but the observation row is {base} entries + "
                             f"{self.num_ray_observations} rays + {extra} further entries")
        self._ray_col0 = base
        self._bind_ray_output()

    def _obs_columns_behind_rays(self):
        return 0

    def _bind_ray_output(self):
        # the ray observations ARE columns of obs_buf: the sensor launch writes them in place
        self.raycast_distances = self.obs_buf[:, self._ray_col0:self._ray_col0 + self.num_ray_observations]
        self.ray_caster.attach_distance_output(self.root_states, 13, self.raycast_distances, self.num_obs, normalize=True)
        self._ray_bound_to = self.obs_buf.data_ptr()

    def _update_ray_sensor(self):
        if getattr(self.cfg.raycaster, "enable_raycast", False) and getattr(self, "ray_caster", None) is not None:
            if self._ray_bound_to != self.obs_buf.data_ptr():
                self._bind_ray_output()
            self.ray_caster.update(dt=self.dt, sensor_pos=self.base_pos, sensor_rot=self.base_quat)

    def _get_raycast_distances(self, env_ids=None, normalize=True):
        """:262-297 on demand (any env subset, raw or normalised) from the sensor's last hits."""
        data = self.ray_caster.data
        if env_ids is not None:
            hits, found, origins = data.ray_hits[env_ids], data.ray_hits_found[env_ids], self.root_states[env_ids, 0:3]
        else:
            hits, found, origins = data.ray_hits, data.ray_hits_found, self.root_states[:, 0:3]
        distances = torch.norm(hits - origins.unsqueeze(1), dim=2)
        if not normalize:
            return distances
        nd = 1.0 - torch.clamp(distances / self.ray_caster.cfg.max_distance, 0.0, 1.0)
        nd = nd * found.float()
        return nd.reshape(nd.shape[0], -1)


class LeggedRobotRayCast(RayCastSensorMixin, LeggedRobot):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.ray_caster = None
        self.raycast_distances = None
        self.num_ray_observations = 0
        if getattr(self.cfg.raycaster, "enable_raycast", False):
            self._init_ray_caster()

    def _pre_step_hook(self):
        """The callback addition of :219-230, placed where this step's observations pick it up."""
        super()._pre_step_hook()
        self._update_ray_sensor()
