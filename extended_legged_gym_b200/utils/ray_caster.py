"""Ray casting sensor -- host side of the B200 path.

Same public surface as the reference module (utils/ray_caster.py in /root/reference/legged_gym/legged_gym):
``convert_to_warp_mesh``, ``raycast_mesh``, ``PatternType``, ``RayCasterPatternCfg``, ``RayCasterCfg``,
``RayCasterData``, ``RayCaster``.  Where the reference copies every batch of rays to the host, hands it to a Warp
kernel and copies the result back (:139-160), the rays here never leave the GPU:

  convert_to_warp_mesh(vertices, triangles)   :29-42    -> Mesh (elg_mesh_create: host SAH build, device BVH)
  raycast_mesh(origins, directions, max, mesh) :95-167   -> elg_raycast
  RayCaster.update(...)                        :518-594  -> elg_raycast_sensor (pattern transform fused with the cast)

There is no CPU fallback: tensors must live on a CUDA device.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Tuple

import numpy as np
import torch

from .. import _lib


class Mesh:
    """Device-resident BVH of one static triangle mesh (the counterpart of ``wp.Mesh``)."""

    def __init__(self, vertices, triangles, device="cuda:0", leaf_triangles=0):
        """leaf_triangles: triangles per BVH leaf, 0 = the library's default (3, best for ray casts); MeshSDF asks for 4"""
        v = np.ascontiguousarray(np.asarray(vertices.cpu() if torch.is_tensor(vertices) else vertices), dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(np.asarray(triangles.cpu() if torch.is_tensor(triangles) else triangles), dtype=np.int32).reshape(-1, 3)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ElgError("the mesh queries have no CPU path: device must be a CUDA device")
        self._lib = _lib.load()
        self.num_vertices, self.num_triangles = len(v), len(t)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self._lib.elg_mesh_create_ex(v.ctypes.data, len(v), t.ctypes.data, len(t), int(leaf_triangles), C.byref(handle))
        _lib.check(rc, "elg_mesh_create_ex")
        self.id = handle.value
        nt, nn = C.c_int32(), C.c_int32()
        b = (C.c_float * 6)()
        _lib.check(self._lib.elg_mesh_info(self.id, C.byref(nt), C.byref(nn), b))
        self.num_nodes = nn.value
        self.bounds = np.array(list(b), dtype=np.float32).reshape(2, 3)
        # (layers, nx, ny) of the regular-grid accelerator a height-field-derived mesh gets on top of the BVH, else (0, 0, 0)
        gl, gx, gy = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.elg_mesh_grid_info(self.id, C.byref(gl), C.byref(gx), C.byref(gy)))
        self.grid = (gl.value, gx.value, gy.value)
        # kept for callers that read the geometry back (wp.Mesh exposes .points / .indices)
        self.points = v
        self.indices = t.reshape(-1)

    def __del__(self):
        try:
            if getattr(self, "id", None):
                self._lib.elg_mesh_free(self.id)
                self.id = None
        except Exception:
            pass


def convert_to_warp_mesh(vertices, triangles, device="cuda:0") -> Mesh:
    """Name kept from the reference (:29-42); returns the BVH handle the other functions take as ``mesh``."""
    return Mesh(vertices, triangles, "cuda:0" if str(device) == "cuda" else device)


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.ElgError("ray casting has no CPU path: tensors must be CUDA tensors")
    return t if (t.dtype == torch.float and t.is_contiguous()) else t.to(torch.float).contiguous()


def raycast_mesh(ray_origins: torch.Tensor, ray_directions: torch.Tensor, max_dist: float = 100.0, mesh: Mesh = None,
                 return_distance: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closest hit of every ray against ``mesh`` (:95-167).  Shapes (batch, n_rays, 3) or (n_rays, 3); returns
    ``hits`` of the same shape (end point at ``max_dist`` on a miss) and the boolean ``hits_found``."""
    if mesh is None:
        raise ValueError("Mesh cannot be None")
    rank = ray_origins.dim()
    if rank not in (2, 3):
        raise ValueError(f"Expected ray_origins to have rank 2 or 3, got {rank}")
    o = _as_f32(ray_origins).reshape(-1, 3)
    d = _as_f32(ray_directions).reshape(-1, 3)
    n = o.shape[0]
    hits = torch.empty(n, 3, dtype=torch.float, device=o.device)
    found = torch.empty(n, dtype=torch.bool, device=o.device)
    dist = torch.empty(n, dtype=torch.float, device=o.device) if return_distance else None
    stream = torch.cuda.current_stream(o.device).cuda_stream
    rc = _lib.load().elg_raycast(mesh.id, o.data_ptr(), d.data_ptr(), n, float(max_dist), hits.data_ptr(), found.data_ptr(),
                                 _lib.ptr(dist), None, stream)
    _lib.check(rc, "elg_raycast")
    if rank == 3:
        hits = hits.reshape(ray_origins.shape[0], ray_origins.shape[1], 3)
        found = found.reshape(ray_origins.shape[0], ray_origins.shape[1])
        if dist is not None:
            dist = dist.reshape(ray_origins.shape[0], ray_origins.shape[1])
    return (hits, found, dist) if return_distance else (hits, found)


class PatternType(Enum):
    SINGLE_RAY = "single_ray"
    GRID = "grid"
    CONE = "cone"
    SPHERICAL = "spherical"
    SPHERICAL2 = "spherical2"


def _quat_apply_cpu(q, b):
    xyz, w = q[:, :3], q[:, 3:]
    t = torch.cross(xyz, b, dim=-1) * 2
    return b + w * t + torch.cross(xyz, t, dim=-1)


@dataclass
class RayCasterPatternCfg:
    """Ray patterns in the sensor frame (:179-363).  Patterns are built once, on the CPU, with the reference's
    arithmetic (fp32 torch), then moved to the device."""
    pattern_type: PatternType = PatternType.SINGLE_RAY
    single_ray_direction: List[float] = field(default_factory=lambda: [1.0, 0.0, 0.0])
    grid_dims: Tuple[int, int] = (5, 5)
    grid_width: float = 1.0
    grid_height: float = 1.0
    cone_num_rays: int = 16
    cone_angle: float = 30.0
    spherical_num_azimuth: int = 8
    spherical_num_elevation: int = 4
    spherical2_num_points: int = 32
    spherical2_polar_axis: List[float] = field(default_factory=lambda: [0.0, 0.0, 1.0])
    ellipsoid_axes: List[float] = field(default_factory=lambda: [1.0, 1.0, 0.3])

    def create_pattern(self, device: str = "cuda:0") -> Tuple[torch.Tensor, torch.Tensor]:
        pt = self.pattern_type
        if pt == PatternType.SINGLE_RAY:
            dirs = torch.tensor([self.single_ray_direction])
        elif pt == PatternType.GRID:                       # rows of [1, x, y], x inner (:224-243)
            rows, cols = self.grid_dims
            xs = torch.linspace(-self.grid_width / 2, self.grid_width / 2, cols)
            ys = torch.linspace(-self.grid_height / 2, self.grid_height / 2, rows)
            dirs = torch.tensor([[1.0, x, y] for y in ys for x in xs])
            dirs = dirs / torch.norm(dirs, dim=1, keepdim=True)
        elif pt == PatternType.CONE:                       # [cos a, sin a cos th, sin a sin th] (:245-264)
            a = self.cone_angle * (np.pi / 180)
            th = torch.linspace(0, 2 * np.pi * (1.0 - 1.0 / self.cone_num_rays), self.cone_num_rays)
            spread = torch.sin(torch.tensor(a))
            fwd = torch.cos(torch.tensor(a))
            dirs = torch.tensor([[fwd.item(), (torch.cos(t) * spread).item(), (torch.sin(t) * spread).item()] for t in th])
            dirs = dirs / torch.norm(dirs, dim=1, keepdim=True)
        elif pt == PatternType.SPHERICAL:                  # elevation major, azimuth inner (:266-284)
            az = torch.linspace(0, 2 * np.pi * (1.0 - 1.0 / self.spherical_num_azimuth), self.spherical_num_azimuth)
            el = torch.linspace(-np.pi / 2, np.pi / 2, self.spherical_num_elevation)
            dirs = torch.tensor([[(torch.cos(e) * torch.cos(a)).item(), (torch.cos(e) * torch.sin(a)).item(), torch.sin(e).item()]
                                 for e in el for a in az])
        elif pt == PatternType.SPHERICAL2:                 # Fibonacci sphere squashed to an ellipsoid (:286-361)
            n = self.spherical2_num_points
            golden = (1 + 5 ** 0.5) / 2
            d = torch.zeros((n, 3))
            for i in range(n):
                y = 1 - (2 * i) / (n - 1)
                r = (1 - y * y) ** 0.5
                th = 2 * np.pi * i / golden
                d[i, 0], d[i, 1], d[i, 2] = r * np.cos(th), y, r * np.sin(th)
            d = d * torch.tensor(self.ellipsoid_axes)
            dirs = d / torch.norm(d, dim=1, keepdim=True)
            if not np.allclose(self.spherical2_polar_axis, [0.0, 0.0, 1.0]):
                pa = torch.tensor(self.spherical2_polar_axis)
                pa = pa / torch.norm(pa)
                z = torch.tensor([0.0, 0.0, 1.0])
                ax = torch.cross(z, pa, dim=0)
                ang = None
                if torch.norm(ax) < 1e-6:
                    if torch.dot(z, pa) <= 0:
                        ax, ang = torch.tensor([1.0, 0.0, 0.0]), torch.tensor(np.pi)
                else:
                    ax = ax / torch.norm(ax)
                    ang = torch.acos(torch.clamp(torch.dot(z, pa), -1.0, 1.0))
                if ang is not None:
                    s = torch.sin(ang / 2)
                    # the reference packs [w, x, y, z] and applies it with the xyzw quat_apply
                    q = torch.stack([torch.cos(ang / 2), ax[0] * s, ax[1] * s, ax[2] * s]).to(torch.float)
                    dirs = _quat_apply_cpu(q.repeat(n, 1), dirs)
        else:
            raise ValueError(f"Unknown pattern type: {pt}")
        dirs = dirs.to(torch.float)
        return torch.zeros_like(dirs).to(device), dirs.to(device)


@dataclass
class RayCasterCfg:
    pattern_cfg: RayCasterPatternCfg = field(default_factory=RayCasterPatternCfg)
    mesh_paths: List[str] = field(default_factory=list)
    vertices: torch.Tensor = None
    triangles: torch.Tensor = None
    max_distance: float = 100.0
    attach_yaw_only: bool = True
    offset_pos: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])
    offset_rot: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0, 1.0])
    update_period: float = 0.0


@dataclass
class RayCasterData:
    ray_hits: torch.Tensor = None
    ray_hits_found: torch.Tensor = None
    pos: torch.Tensor = None
    rot: torch.Tensor = None


def load_obj(path):
    """Minimal Wavefront OBJ reader (v / f records, polygons fan-triangulated) -- init-time replacement of trimesh.load."""
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            p = line.split()
            if not p:
                continue
            if p[0] == "v":
                vs.append([float(x) for x in p[1:4]])
            elif p[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in p[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, dtype=np.float32), np.asarray(fs, dtype=np.int32)


class RayCaster:
    """Ray casting sensor attached to every env (:405-618)."""

    def __init__(self, cfg: RayCasterCfg, num_envs: int, device: str = "cuda:0"):
        self.cfg, self.num_envs, self.device = cfg, num_envs, device
        self._is_initialized = False
        self._timestamp = torch.zeros(num_envs, device=device)
        self._timestamp_last_update = torch.zeros(num_envs, device=device)
        self._is_outdated = torch.ones(num_envs, dtype=torch.bool, device=device)
        self.meshes = {}
        self.ray_origins = self.ray_directions = None
        self.num_rays = 0
        self._data = RayCasterData()
        self._distance_sink = None
        self._initialize()

    def attach_distance_output(self, origins, origin_stride, out, out_row_stride, normalize=True):
        """Fuse LeggedRobotRayCast._get_raycast_distances (envs/base/legged_robot_raycast.py:262-297) into every update: ``out``
        receives (1 - clamp(|hit - origins[e]| / max_distance, 0, 1)) * found per ray (or the plain distance), rows
        ``out_row_stride`` floats apart; ``origins`` rows ``origin_stride`` floats apart (root_states: 13).  None detaches."""
        self._distance_sink = None if out is None else (origins, int(origin_stride), out, int(out_row_stride), bool(normalize))

    def _initialize(self):
        if self._is_initialized:
            return
        self._initialize_meshes()
        self._initialize_rays()
        self._is_initialized = True

    def _initialize_meshes(self):
        if self.cfg.mesh_paths:
            for path in self.cfg.mesh_paths:
                if not os.path.isfile(path):
                    print(f"Failed to load mesh {path}: file not found")
                    continue
                v, f = load_obj(path)
                self.meshes[path] = Mesh(v, f, self.device)
        elif self.cfg.vertices is not None and self.cfg.triangles is not None:
            self.meshes["custom_mesh"] = Mesh(self.cfg.vertices, self.cfg.triangles, self.device)
        else:
            raise ValueError("No mesh or vertices/triangles provided for ray casting.")
        if not self.meshes:
            raise RuntimeError("No meshes were successfully loaded or created.")

    def _initialize_rays(self):
        o, d = self.cfg.pattern_cfg.create_pattern(self.device)
        self.num_rays = len(d)
        self._pattern_origins = (o + torch.tensor(self.cfg.offset_pos, device=self.device)).contiguous()
        self._pattern_directions = d.contiguous()
        # the reference materialises the pattern per env (:499-500); kept as (stride-0) views for API compatibility
        self.ray_origins = self._pattern_origins.unsqueeze(0).expand(self.num_envs, -1, -1)
        self.ray_directions = self._pattern_directions.unsqueeze(0).expand(self.num_envs, -1, -1)
        self._data.pos = torch.zeros(self.num_envs, 3, device=self.device)
        self._data.rot = torch.zeros(self.num_envs, 4, device=self.device)
        self._data.rot[:, 3] = 1.0
        self._data.ray_hits = torch.zeros(self.num_envs, self.num_rays, 3, device=self.device)
        self._data.ray_hits_found = torch.zeros(self.num_envs, self.num_rays, dtype=torch.bool, device=self.device)

    def update(self, dt: float, sensor_pos: torch.Tensor, sensor_rot: torch.Tensor, env_ids: torch.Tensor = None):
        """sensor_rot is used as xyzw, like the reference does whatever its docstring says."""
        self._timestamp += dt
        all_envs = False
        if env_ids is None:
            if self.cfg.update_period <= 0.0:
                all_envs = True          # every env is outdated every step: no nonzero(), no host sync (:535-536)
            else:
                self._is_outdated |= (self._timestamp - self._timestamp_last_update + 1e-6 >= self.cfg.update_period)
                env_ids = self._is_outdated.nonzero().squeeze(-1)
        else:
            self._is_outdated[env_ids] = True
        if not all_envs and len(env_ids) == 0:
            return
        if all_envs:
            self._data.pos.copy_(sensor_pos)
            self._data.rot.copy_(sensor_rot)
        else:
            self._data.pos[env_ids] = sensor_pos[env_ids].to(self._data.pos.dtype)
            self._data.rot[env_ids] = sensor_rot[env_ids].to(self._data.rot.dtype)
        self._update_ray_casting(None if all_envs else env_ids)
        if all_envs:
            self._timestamp_last_update.copy_(self._timestamp)
            self._is_outdated.zero_()
        else:
            self._timestamp_last_update[env_ids] = self._timestamp[env_ids]
            self._is_outdated[env_ids] = False

    def _update_ray_casting(self, env_ids):
        mesh = next(iter(self.meshes.values()))     # only the first mesh is used (:583-584)
        ids = None
        n = self.num_envs
        if env_ids is not None:
            ids = env_ids.to(torch.int64).contiguous()
            n = len(ids)
        stream = torch.cuda.current_stream(self._data.pos.device).cuda_stream
        if self._distance_sink is not None:
            org, ostride, out, rstride, norm = self._distance_sink
            rc = _lib.load().elg_raycast_sensor_obs(mesh.id, self._pattern_origins.data_ptr(), self._pattern_directions.data_ptr(), self.num_rays,
                                                    self._data.pos.data_ptr(), self._data.rot.data_ptr(), _lib.ptr(ids), n,
                                                    int(bool(self.cfg.attach_yaw_only)), float(self.cfg.max_distance),
                                                    self._data.ray_hits.data_ptr(), self._data.ray_hits_found.data_ptr(),
                                                    org.data_ptr(), ostride, int(norm), out.data_ptr(), rstride, stream)
            _lib.check(rc, "elg_raycast_sensor_obs")
            return
        rc = _lib.load().elg_raycast_sensor(mesh.id, self._pattern_origins.data_ptr(), self._pattern_directions.data_ptr(), self.num_rays,
                                            self._data.pos.data_ptr(), self._data.rot.data_ptr(), _lib.ptr(ids), n,
                                            int(bool(self.cfg.attach_yaw_only)), float(self.cfg.max_distance),
                                            self._data.ray_hits.data_ptr(), self._data.ray_hits_found.data_ptr(), stream)
        _lib.check(rc, "elg_raycast_sensor")

    def reset(self, env_ids=None):
        if env_ids is None:
            env_ids = torch.arange(self.num_envs, device=self.device)
        self._timestamp[env_ids] = 0.0
        self._timestamp_last_update[env_ids] = 0.0
        self._is_outdated[env_ids] = True

    @property
    def data(self) -> RayCasterData:
        return self._data
