"""TEST INFRASTRUCTURE ONLY (oracle) -- the checker, never the product path.

torch-CPU restatement of the reference's Warp depth camera (utils/depth_camera.py in
/root/reference/legged_gym/legged_gym): DepthCameraWarp._initialize_ray_grid (:328-378), .update (:501-566),
.update_depth_buffer (:402-499) and DepthCameraBase.process_depth_image / normalize_depth_image (:56-69, :84-138),
same torch ops in the same order.  The ray cast itself (Warp in the reference) is oracle/mesh_oracle.raycast_mesh;
the resize is torchvision's own Resize, exactly the object the reference constructs (:33-36).
``tests/test_depth_camera.py`` pins it against the unmodified reference class (container only, Warp calls patched to
the brute force) and against tests/golden/depth_camera.npz generated the same way.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
import numpy as np
import torch

from . import mesh_oracle as mo
from . import torch_utils as tu


class DepthOracle:
    def __init__(self, cfg, num_envs, vertices, triangles):
        import torchvision
        self.cfg, self.num_envs = cfg, num_envs
        self.v, self.t = np.asarray(vertices, np.float32), np.asarray(triangles, np.int32)
        self.resize_transform = torchvision.transforms.Resize((cfg.resized[1], cfg.resized[0]),
                                                              interpolation=torchvision.transforms.InterpolationMode.BICUBIC)
        self.depth_buffer = torch.zeros(num_envs, cfg.buffer_len, cfg.resized[1], cfg.resized[0])
        self.camera_pos = torch.zeros(num_envs, 3)
        self.camera_rot = torch.zeros(num_envs, 4)
        self.camera_rot[:, 3] = 1.0
        width, height = cfg.original
        hfov = cfg.horizontal_fov
        vfov = 2 * np.arctan(np.tan(np.radians(hfov) / 2) / (width / height))
        vfov_degrees = np.degrees(vfov)
        i, j = torch.meshgrid(torch.linspace(-1, 1, height), torch.linspace(-1, 1, width), indexing="ij")
        i = i * np.tan(np.radians(vfov_degrees / 2))
        j = j * np.tan(np.radians(hfov / 2))
        d = torch.stack([torch.ones_like(i), j, i], dim=-1)
        d = d / torch.norm(d, dim=-1, keepdim=True)
        self.ray_directions = d.reshape(-1, 3).repeat(num_envs, 1, 1)
        self.ray_origins = torch.zeros_like(self.ray_directions)

    def update(self, sensor_pos, sensor_rot):
        from scipy.spatial.transform import Rotation as R
        cfg = self.cfg
        off = torch.tensor(getattr(cfg, "position", [0.0, 0.0, 0.0]), dtype=torch.float)
        if hasattr(cfg, "rotation"):
            q = torch.tensor(R.from_euler("xyz", cfg.rotation, degrees=True).as_quat(), dtype=torch.float)
            q = torch.tensor([q[3], q[0], q[1], q[2]], dtype=torch.float)
        elif hasattr(cfg, "angle") and len(cfg.angle) == 2:
            q = torch.tensor(R.from_euler("y", -np.mean(cfg.angle), degrees=True).as_quat(), dtype=torch.float)
            q = torch.tensor([q[3], q[0], q[1], q[2]], dtype=torch.float)
        else:
            q = torch.tensor([1.0, 0.0, 0.0, 0.0])
        n = sensor_pos.shape[0]
        self.camera_pos[:] = sensor_pos + tu.quat_apply(sensor_rot, off.expand(n, -1))
        self.camera_rot[:] = tu.quat_mul(sensor_rot, q.expand(n, -1))      # wxyz 4-vector into the xyzw quat_mul (App. A-9)

    def process_depth_image(self, depth_image, noise_u=None):
        cfg = self.cfg
        b = depth_image.shape[0]
        if hasattr(cfg, "dis_noise"):
            u = torch.rand(b) if noise_u is None else noise_u
            depth_image = depth_image + (cfg.dis_noise * 2 * (u - 0.5)).view(b, 1, 1)
        depth_image = torch.clip(depth_image, -cfg.far_clip, -cfg.near_clip)
        if cfg.resized[0] != cfg.original[0] or cfg.resized[1] != cfg.original[1]:
            depth_image = self.resize_transform(depth_image.unsqueeze(1)).squeeze(1)
        depth_image = depth_image * -1
        return (depth_image - cfg.near_clip) / (cfg.far_clip - cfg.near_clip) - 0.5

    def update_depth_buffer(self, episode_length_buf, noise_u=None):
        cfg = self.cfg
        b, n_rays, _ = self.ray_origins.shape
        rot = self.camera_rot.unsqueeze(1).expand(-1, n_rays, -1).reshape(-1, 4)
        o = tu.quat_apply(rot, self.ray_origins.reshape(-1, 3)).reshape(b, n_rays, 3) + self.camera_pos.unsqueeze(1)
        d = tu.quat_apply(rot, self.ray_directions.reshape(-1, 3)).reshape(b, n_rays, 3)
        hits, found, _, _ = mo.raycast_mesh(o.reshape(-1, 3).numpy(), d.reshape(-1, 3).numpy(), cfg.far_clip, self.v, self.t)
        hits = torch.from_numpy(hits).reshape(b, n_rays, 3)
        found = torch.from_numpy(found).reshape(b, n_rays)
        dist = torch.norm(hits - self.camera_pos.unsqueeze(1), dim=2)
        depth = torch.where(found, -dist, torch.tensor(-cfg.far_clip)).reshape(b, cfg.original[1], cfg.original[0])
        self.raw_depth = depth.clone()
        proc = self.process_depth_image(depth, noise_u)
        init = episode_length_buf <= 1
        for e in range(b):
            if init[e]:
                self.depth_buffer[e] = torch.stack([proc[e]] * cfg.buffer_len, dim=0)
            else:
                self.depth_buffer[e] = torch.cat([self.depth_buffer[e, 1:], proc[e].unsqueeze(0)], dim=0)
        return proc
