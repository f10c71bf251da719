"""Depth cameras -- host side of the B200 path.

Same classes and methods as the reference module (utils/depth_camera.py in /root/reference/legged_gym/legged_gym):
``DepthCameraBase`` (:13-183), ``DepthCameraFake`` (:186-254), ``DepthCameraWarp`` (:256-571).  The simulator-rendered
``DepthCamera`` (:573-728, Isaac Gym camera sensors) is outside the hot path and not provided.

``DepthCameraWarp.update`` is one launch of ``elg_camera_pose``; ``update_depth_buffer`` -- ray rotation, ray cast
(there: a host round trip through Warp), -distance image, noise, clip, bicubic resize, normalisation and the per-env
Python loop over the frame ring buffer -- is ONE launch of ``elg_depth_camera``.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .ray_caster import Mesh


def _resize_taps(n_in, n_out, resize_1d):
    """Taps of a separable linear resize along one axis, read off the resize operator itself: row k of
    ``resize_1d(eye(n_in))`` is the response of every output sample to input sample k."""
    resp = resize_1d(torch.eye(n_in))                       # [n_in, n_out]
    resp = resp.t().contiguous().numpy()                    # [n_out, n_in]
    starts, rows, width = [], [], 1
    for o in range(n_out):
        nz = np.nonzero(resp[o])[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if len(nz) else (0, 1)
        starts.append(lo)
        rows.append(resp[o, lo:hi])
        width = max(width, hi - lo)
    w = np.zeros((n_out, width), dtype=np.float32)
    for o, r in enumerate(rows):
        w[o, :len(r)] = r
    return np.asarray(starts, dtype=np.int32), w


class DepthCameraBase:
    def __init__(self, cfg, device, num_envs):
        self.cfg, self.device, self.num_envs = cfg, device, num_envs
        self.depth_buffer = torch.zeros(num_envs, cfg.buffer_len, cfg.resized[1], cfg.resized[0], device=device)

    def create_camera(self, env_handle, actor_handle, env_id=None):
        return None

    def normalize_depth_image(self, depth_image):
        depth_image = depth_image * -1
        return (depth_image - self.cfg.near_clip) / (self.cfg.far_clip - self.cfg.near_clip) - 0.5

    def get_depth_observation(self):
        if self.cfg.camera_type is None:
            return None
        return self.depth_buffer[:, -2]            # second-to-last frame (:140-149)

    def get_depth_buffer(self):
        return self.depth_buffer

    def is_enabled(self):
        return self.cfg.camera_type is not None


class DepthCameraFake(DepthCameraBase):
    """Constant -0.5 frames (:186-254)."""

    def __init__(self, cfg, device, num_envs):
        super().__init__(cfg, device, num_envs)
        self.depth_buffer = torch.full((num_envs, cfg.buffer_len, cfg.resized[1], cfg.resized[0]), -0.5, device=device)
        self.camera_pos = torch.zeros(num_envs, 3, device=device)
        self.camera_rot = torch.zeros(num_envs, 4, device=device)
        self.camera_rot[:, 3] = 1.0

    def update_depth_buffer(self, envs, episode_length_buf):
        pass

    def update(self, dt, sensor_pos, sensor_rot, env_ids=None):
        pass


class DepthCameraWarp(DepthCameraBase):
    def __init__(self, cfg, device, num_envs, terrain_vertices=None, terrain_triangles=None):
        super().__init__(cfg, device, num_envs)
        self._lib = _lib.load()
        self.camera_pos = torch.zeros(num_envs, 3, device=device)
        self.camera_rot = torch.zeros(num_envs, 4, device=device)
        self.camera_rot[:, 3] = 1.0
        self.actor_handles = [None] * num_envs
        self.terrain_vertices, self.terrain_triangles = terrain_vertices, terrain_triangles
        self.meshes = {}
        self.ray_origins = self.ray_directions = None
        self.noise_u = None          # [num_envs] uniform samples for parity with torch.rand(batch) (:122), else drawn per call
        self.raw_depth = None        # set to a [num_envs, h, w] tensor to also receive the unprocessed image
        if terrain_vertices is not None and terrain_triangles is not None:
            self.meshes["terrain"] = Mesh(terrain_vertices, terrain_triangles, device)
        self._initialize_ray_grid()
        self._offsets = self._mount_offsets()

    # :328-378 -- CPU torch arithmetic of the reference, then moved to the device
    def _initialize_ray_grid(self):
        if self.cfg.camera_type is None:
            return
        width, height = self.cfg.original
        hfov = self.cfg.horizontal_fov
        vfov = 2 * np.arctan(np.tan(np.radians(hfov) / 2) / (width / height))
        vfov_degrees = np.degrees(vfov)
        i, j = torch.meshgrid(torch.linspace(-1, 1, height), torch.linspace(-1, 1, width), indexing="ij")
        i = i * np.tan(np.radians(vfov_degrees / 2))
        j = j * np.tan(np.radians(hfov / 2))
        d = torch.stack([torch.ones_like(i), j, i], dim=-1)
        d = d / torch.norm(d, dim=-1, keepdim=True)
        self._grid_directions = d.reshape(-1, 3).to(self.device).contiguous()
        n = self._grid_directions.shape[0]
        self.ray_origins = torch.zeros(1, n, 3, device=self.device).expand(self.num_envs, -1, -1)
        self.ray_directions = self._grid_directions.unsqueeze(0).expand(self.num_envs, -1, -1)
        self._cam = _lib.ElgCamParams()
        c = self._cam
        c.width, c.height = width, height
        c.out_width, c.out_height = self.cfg.resized
        c.buffer_len = self.cfg.buffer_len
        c.near_clip, c.far_clip = self.cfg.near_clip, self.cfg.far_clip
        c.noise_scale = float(getattr(self.cfg, "dis_noise", 0.0)) * 2
        c.resize = int(tuple(self.cfg.resized) != tuple(self.cfg.original))
        self._taps = None
        if c.resize:
            import torchvision
            bic = torchvision.transforms.InterpolationMode.BICUBIC
            rx = torchvision.transforms.Resize((width, self.cfg.resized[0]), interpolation=bic)      # rows = input columns
            ry = torchvision.transforms.Resize((self.cfg.resized[1], height), interpolation=bic)
            xs, xw = _resize_taps(width, self.cfg.resized[0], lambda eye: rx(eye[None])[0])
            ys, yw = _resize_taps(height, self.cfg.resized[1], lambda eye: ry(eye.t()[None])[0].t())
            taps = max(xw.shape[1], yw.shape[1])
            pad = lambda w: np.pad(w, ((0, 0), (0, taps - w.shape[1])))
            dev = self.device
            self._taps = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (xs, pad(xw), ys, pad(yw)))
            c.max_taps = taps

    def _mount_offsets(self):
        """:519-549 -- mounting position and the 4-vector the reference builds from cfg.rotation / cfg.angle."""
        off = np.asarray(getattr(self.cfg, "position", [0.0, 0.0, 0.0]), dtype=np.float32)
        if hasattr(self.cfg, "rotation"):
            from scipy.spatial.transform import Rotation as R
            q = torch.tensor(R.from_euler("xyz", self.cfg.rotation, degrees=True).as_quat(), dtype=torch.float)
            q = torch.tensor([q[3], q[0], q[1], q[2]], dtype=torch.float)
        elif hasattr(self.cfg, "angle") and len(self.cfg.angle) == 2:
            from scipy.spatial.transform import Rotation as R
            q = torch.tensor(R.from_euler("y", -np.mean(self.cfg.angle), degrees=True).as_quat(), dtype=torch.float)
            q = torch.tensor([q[3], q[0], q[1], q[2]], dtype=torch.float)
        else:
            q = torch.tensor([1.0, 0.0, 0.0, 0.0])
        return np.ascontiguousarray(off), np.ascontiguousarray(q.numpy().astype(np.float32))

    def create_camera(self, env_handle, actor_handle, env_id=None):
        if self.cfg.camera_type is not None and env_id is not None and 0 <= env_id < self.num_envs:
            self.actor_handles[env_id] = actor_handle
        return None

    def update(self, dt, sensor_pos, sensor_rot, env_ids=None):
        if self.cfg.camera_type is None:
            return
        ids, n = None, self.num_envs
        if env_ids is not None:
            if len(env_ids) == 0:
                return
            ids = env_ids.to(torch.int64).contiguous()
            n = len(ids)
        pos = sensor_pos if (sensor_pos.dtype == torch.float and sensor_pos.is_contiguous()) else sensor_pos.to(torch.float).contiguous()
        rot = sensor_rot if (sensor_rot.dtype == torch.float and sensor_rot.is_contiguous()) else sensor_rot.to(torch.float).contiguous()
        off, q = self._offsets
        stream = torch.cuda.current_stream(self.camera_pos.device).cuda_stream
        rc = self._lib.elg_camera_pose(pos.data_ptr(), rot.data_ptr(), _lib.ptr(ids), n, off.ctypes.data, q.ctypes.data,
                                       self.camera_pos.data_ptr(), self.camera_rot.data_ptr(), stream)
        _lib.check(rc, "elg_camera_pose")

    def update_depth_buffer(self, envs, episode_length_buf):
        if self.cfg.camera_type is None:
            return
        if not self.meshes:
            print("Warning: No meshes available for ray casting.")
            return
        c = self._cam
        u = None
        if c.noise_scale != 0.0:
            u = self.noise_u if self.noise_u is not None else torch.rand(self.num_envs, device=self.device)
        taps = self._taps or (None, None, None, None)
        ep = episode_length_buf if episode_length_buf.dtype == torch.int64 else episode_length_buf.to(torch.int64)
        stream = torch.cuda.current_stream(self.camera_pos.device).cuda_stream
        rc = self._lib.elg_depth_camera(self.meshes["terrain"].id, C.byref(c), self._grid_directions.data_ptr(), self.camera_pos.data_ptr(),
                                        self.camera_rot.data_ptr(), ep.data_ptr(), _lib.ptr(u), _lib.ptr(taps[0]), _lib.ptr(taps[1]),
                                        _lib.ptr(taps[2]), _lib.ptr(taps[3]), self.num_envs, self.depth_buffer.data_ptr(),
                                        _lib.ptr(self.raw_depth), stream)
        _lib.check(rc, "elg_depth_camera")
