"""Generate tests/golden/depth_camera.npz from the UNMODIFIED reference class DepthCameraWarp
(utils/depth_camera.py:256-571), container only.  The two Warp entry points it reaches (convert_to_warp_mesh,
raycast_mesh -- the wheel is absent) are patched to the float64 brute force of oracle/mesh_oracle.py; everything else
(ray grid, camera pose, depth post-processing, torchvision resize, ring buffer) is the reference's own code.

    python tests/golden/make_depth_golden.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mesh_oracle as mo, ref_harness  # noqa: E402

CASES = {
    "resize_noise": dict(original=(20, 12), resized=(16, 10), far_clip=2.0, near_clip=0.0, dis_noise=0.02, buffer_len=2, n=5, seed=0),
    "plain_far10": dict(original=(16, 8), resized=(16, 8), far_clip=10.0, near_clip=0.1, dis_noise=0.0, buffer_len=3, n=4, seed=1),
}


def make_cfg(c):
    return SimpleNamespace(camera_type="Warp", original=c["original"], resized=c["resized"], horizontal_fov=100, buffer_len=c["buffer_len"],
                           near_clip=c["near_clip"], far_clip=c["far_clip"], dis_noise=c["dis_noise"], position=[0.5, 0, 0.03], angle=[30, 30],
                           update_interval=1, scale=1, invert=True)


def poses(n, seed):
    g = torch.Generator().manual_seed(seed)
    pos = torch.stack([torch.rand(n, generator=g) * 2.5 + 0.7, torch.rand(n, generator=g) * 2.5 + 0.7, torch.rand(n, generator=g) * 0.3 + 0.5], 1)
    q = torch.randn(n, 4, generator=g) * torch.tensor([0.1, 0.1, 1.0, 1.0])
    return pos, q / q.norm(dim=1, keepdim=True)


def reference_run(c):
    ref_harness.install()
    from legged_gym.envs.base.legged_robot import LeggedRobot  # noqa: F401
    import legged_gym.utils.depth_camera as dc
    import legged_gym.utils.ray_caster as rcm
    v, t, _ = mo.heightfield_mesh(40, 40, seed=3)

    def fake_raycast(ray_origins, ray_directions, max_dist=100.0, mesh=None):
        h, f, _, _ = mo.raycast_mesh(ray_origins.reshape(-1, 3).numpy(), ray_directions.reshape(-1, 3).numpy(), max_dist, v, t)
        return torch.from_numpy(h), torch.from_numpy(f)
    dc.convert_to_warp_mesh = lambda vertices, triangles, device=None: "mesh"
    rcm.raycast_mesh = fake_raycast
    cam = dc.DepthCameraWarp(make_cfg(c), "cpu", c["n"], v, t)
    out = {"ray_directions": cam.ray_directions[0].numpy()}
    for step, ep in enumerate(([0, 1, 5, 7, 9][:c["n"]], [1, 2, 6, 8, 10][:c["n"]], [2, 3, 7, 9, 11][:c["n"]])):
        pos, quat = poses(c["n"], 10 * c["seed"] + step)
        cam.update(0.02, pos, quat)
        torch.manual_seed(500 + step)
        u = torch.rand(c["n"])
        torch.manual_seed(500 + step)
        cam.update_depth_buffer(None, torch.tensor(ep))
        out.update({f"s{step}__pos": pos.numpy(), f"s{step}__quat": quat.numpy(), f"s{step}__ep": np.asarray(ep, np.int64), f"s{step}__u": u.numpy(),
                    f"s{step}__camera_pos": cam.camera_pos.numpy().copy(), f"s{step}__camera_rot": cam.camera_rot.numpy().copy(),
                    f"s{step}__depth_buffer": cam.depth_buffer.numpy().copy(), f"s{step}__obs": cam.get_depth_observation().numpy().copy()})
    return out


def main():
    out = {}
    for name, c in CASES.items():
        for k, v in reference_run(c).items():
            out[f"{name}__{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "depth_camera.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
