"""``RobotBatchRolloutPercept`` -- the main / rollout env with perception: a ray caster on every row and the signed distance
of selected robot bodies to the terrain mesh.

Mirrors envs/batch_rollout/robot_batch_rollout_percept.py:20-470 of the reference (``_init_ray_caster`` :60, ``_init_sdf``
:213, the two callbacks :300 / :325, ``_update_sdf_values`` :385, ``compute_observations`` :443).  Per step the reference
runs, for the rays, a Warp round trip + a norm / clamp / mask chain, and for the SDF a Python loop over the query bodies with
a gather, a ``quat_rotate`` and one Warp round trip per body (two with ``compute_nearest_points``); here each sensor is ONE
launch -- ``elg_raycast_sensor_obs`` and ``elg_sdf_query_bodies`` -- writing its observation columns of ``obs_buf`` in place:
``[ base 12 + 3D | heights H | rays | sdf values ]`` (``cfg.env.num_observations`` counts all of them, as in the reference).
"""
import numpy as np
import torch

from ... import _lib
from ...utils.mesh_sdf import MeshSDF, MeshSDFCfg
from ..base.legged_robot_raycast import RayCastSensorMixin
from .robot_batch_rollout import RobotBatchRollout


class RobotBatchRolloutPercept(RayCastSensorMixin, RobotBatchRollout):
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True):
        super().__init__(cfg, sim_params, physics_engine, sim_device, headless)
        self.ray_caster = None
        self.raycast_distances = None
        self.num_ray_observations = 0
        self.mesh_sdf = None
        self.num_sdf_bodies = 0
        sdf_on = hasattr(self.cfg, "sdf") and getattr(self.cfg.sdf, "enable_sdf", False)
        if sdf_on:
            self._init_sdf()          # (the observation layout needs the body count before the rays are bound)
        if hasattr(self.cfg, "raycaster") and getattr(self.cfg.raycaster, "enable_raycast", False):
            self._init_ray_caster()
        if sdf_on:
            self._bind_sdf_output()

    # ------------------------------------------------------------------------------------------
    def _obs_columns_behind_rays(self):
        return self.num_sdf_bodies if (self.mesh_sdf is not None and getattr(self.cfg.sdf, "include_in_obs", True)) else 0

    def _init_sdf(self):
        """:213-298 -- mesh from files, from the terrain, or the ground quad; query bodies by name (root body when none matches)."""
        c = self.cfg.sdf
        scfg = MeshSDFCfg(max_distance=c.max_distance, enable_caching=False)     # (the byte-string cache forces a D2H copy per query)
        if getattr(c, "mesh_paths", None):
            scfg.mesh_paths = list(c.mesh_paths)
        elif getattr(self.cfg.terrain, "mesh_file", None) and getattr(self.cfg.terrain, "use_terrain_obj", False):
            scfg.mesh_paths = [self.cfg.terrain.mesh_file]
        else:
            scfg.vertices, scfg.triangles = self._terrain_mesh()
        self.mesh_sdf = MeshSDF(scfg, device=self.device)
        names = self.sim.spec.body_names
        self.sdf_body_indices = [names.index(b) for b in c.query_bodies if b in names]
        for b in c.query_bodies:
            if b not in names:
                print(f"Warning: Body '{b}' not found for SDF query")
        if not self.sdf_body_indices:
            self.sdf_body_indices = [0]
            print("No valid SDF query bodies found, defaulting to root body")
        self.num_sdf_bodies = K = len(self.sdf_body_indices)
        N = self.total_num_envs
        self.sdf_gradients = torch.zeros(N, K, 3, device=self.device)
        self.sdf_nearest_points = torch.zeros(N, K, 3, device=self.device)
        self.sdf_query_points = torch.zeros(N, K, 3, device=self.device)
        self.sdf_update_counter = 0
        offs = np.full((K, 3), np.nan, dtype=np.float32)          # NaN: no collision-sphere offset for that body (:401-414)
        pos = list(getattr(c, "collision_sphere_pos", []))
        for i in range(min(K, len(pos))):
            offs[i] = pos[i]
        self._sdf_body_idx = np.asarray(self.sdf_body_indices, dtype=np.int32)
        self._sdf_offsets = offs
        self.sdf_values = torch.zeros(N, K, device=self.device)

    def _bind_sdf_output(self):
        if getattr(self.cfg.sdf, "include_in_obs", True):
            base = 12 + 3 * self.num_dof + (self.num_height_points if self.measure_heights else 0) + self.num_ray_observations
            if self.num_obs != base + self.num_sdf_bodies:
                raise ValueError(f"cfg.env.num_observations = {self.num_obs}, but the observation row is {base} entries + {self.num_sdf_bodies} SDF values")
            self.sdf_values = self.obs_buf[:, base:base + self.num_sdf_bodies]     # written in place by the query launch
        self._sdf_bound_to = self.obs_buf.data_ptr()

    def _update_sdf_values(self, env_ids=None):
        """:385-441 as one launch over (env, query body)."""
        if self._sdf_bound_to != self.obs_buf.data_ptr():
            self._bind_sdf_output()
        ids, n = None, self.total_num_envs
        if env_ids is not None:
            ids = env_ids.to(torch.int64).contiguous()
            n = len(ids)
        c = self.cfg.sdf
        mesh = next(iter(self.mesh_sdf.meshes.values()))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self._lib.elg_sdf_query_bodies(
            mesh.id, self.rigid_body_state.data_ptr(), self.num_bodies, self._sdf_body_idx.ctypes.data, self._sdf_offsets.ctypes.data,
            self.num_sdf_bodies, _lib.ptr(ids), n, float(c.max_distance), MeshSDF.EPSILON, self.sdf_values.data_ptr(), self.sdf_values.stride(0),
            self.sdf_gradients.data_ptr() if getattr(c, "compute_gradients", True) else None,
            self.sdf_nearest_points.data_ptr() if getattr(c, "compute_nearest_points", True) else None,
            self.sdf_query_points.data_ptr(), stream)
        _lib.check(rc, "elg_sdf_query_bodies")

    # ------------------------------------------------------------------------------------------
    def _update_perception(self):
        self._update_ray_sensor()
        if self.mesh_sdf is not None and getattr(self.cfg.sdf, "enable_sdf", False):
            self.sdf_update_counter += 1
            if self.sdf_update_counter >= self.cfg.sdf.update_freq:
                self.sdf_update_counter = 0
                self._update_sdf_values()

    def _pre_step_hook(self):
        """_post_physics_step_callback's additions (:300-323)"""
        super()._pre_step_hook()
        self._update_perception()

    def _pre_step_hook_rollout(self):
        """_post_physics_step_callback_rollout (:325-347)"""
        super()._pre_step_hook_rollout()
        self._update_perception()
